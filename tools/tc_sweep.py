"""A/B of the tensor-core path (maf_tc.cu, RABI_TC=1) against the FMA kernels (RABI_TC=0) at lotka_volterra shapes:
agreement of the attack gradient / KL / metric, and CUDA-event times of one PGD step and of the S=256 metric pass."""
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
import bench  # noqa: E402
from rabi_b200.models import ZScoreLayer  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "lotka_volterra_l2pgd_rkl"
w = dict(bench.WORKLOADS[wl])
dx = w["dx"]
dev = torch.device("cuda:0")
model = bench.make_weights(w)
model.embedding_net = torch.nn.Sequential(ZScoreLayer(torch.zeros(dx), torch.ones(dx)), model.embedding_net)
model = model.to(dev).eval()
eng = model.engine()
zm, zs = torch.zeros(dx, device=dev), torch.ones(dx, device=dev)
D = w["D"]


def t(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


for S, B in ((5, 100), (5, 300), (5, 1250), (5, 10000), (1, 512)):
    torch.manual_seed(S * 100000 + B)
    x = torch.randn(B, dx, device=dev)
    xq = x + 0.3 * torch.randn(B, dx, device=dev)
    z = torch.randn(S, B, D, device=dev)
    A = eng.ctx_project(x, zm, zs)
    out = {}
    for tc in ("0", "1"):
        os.environ["RABI_TC"] = tc
        gx, kl = eng.rkl_input_grad(xq, zm, zs, A, z, 1.0 / B, want_kl=True)
        gf, kf = eng.rkl_input_grad(xq, zm, zs, A, z, 1.0 / B, want_kl=True, forward_kl=True)
        Aq = eng.ctx_project(xq, zm, zs)
        th, lq = eng.rsample(Aq, z)
        lp = eng.log_prob(A, th)
        delta = torch.zeros_like(x)
        us = t(lambda: eng.pgd_step(x, delta, zm, zs, A, z, 0.5, 0.05, -3.0, 3.0, 1.0 / B))
        out[tc] = (gx, kl, gf, kf, th, lq, lp, us)
    a, b_ = out["1"], out["0"]
    print(f"S={S} B={B:6d} pairs={S*B:6d}: pgd_step fma {b_[7]:7.1f} us  tc {a[7]:7.1f} us  x{b_[7]/a[7]:.2f} | rel.diff grad {rel(a[0], b_[0]):.1e} "
          f"kl {rel(a[1], b_[1]):.1e} fkl-grad {rel(a[2], b_[2]):.1e} fkl {rel(a[3], b_[3]):.1e} theta {rel(a[4], b_[4]):.1e} "
          f"logq {rel(a[5], b_[5]):.1e} logp {rel(a[6], b_[6]):.1e}", flush=True)

B = 10000
x = torch.randn(B, dx, device=dev)
A = eng.ctx_project(x, zm, zs)
A2 = eng.ctx_project(x + 0.3 * torch.randn_like(x), zm, zs)
ze = torch.randn(256, B, D, device=dev)
r = {}
for tc in ("0", "1"):
    os.environ["RABI_TC"] = tc
    r[tc] = (eng.rkl(A2, A, ze), t(lambda: eng.rkl(A2, A, ze), n=5))
print(f"metric S=256 B={B}: fma {r['0'][1]:.0f} us  tc {r['1'][1]:.0f} us  x{r['0'][1]/r['1'][1]:.2f}  rel.diff {rel(r['1'][0], r['0'][0]):.1e}")
