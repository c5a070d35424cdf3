import os, sys, torch
sys.path.insert(0, os.getcwd())
import bench
w = dict(bench.WORKLOADS["lotka_volterra_l2pgd_rkl"])
from rabi_b200.models import ZScoreLayer
dev = torch.device("cuda:0")
model = bench.make_weights(w)
model.embedding_net = torch.nn.Sequential(ZScoreLayer(torch.zeros(100), torch.ones(100)), model.embedding_net)
model = model.to(dev).eval()
eng = model.engine()
zm, zs = torch.zeros(100, device=dev), torch.ones(100, device=dev)
def t(fn, n=100):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for S in (5,):
    for B in (1000, 5000, 10000):
        if S == 1: B *= 4
        x = torch.randn(B, 100, device=dev)
        z = torch.randn(S, B, 4, device=dev)
        delta = torch.zeros_like(x)
        A = eng.ctx_project(x, zm, zs)
        r = []
        for f2 in ("0", "1"):
            os.environ["RABI_FAST2"] = f2
            r.append(t(lambda: eng.pgd_step(x, delta, zm, zs, A, z, 0.5, 0.05, -3., 3., 1e-3)))
        print(f"S={S} B={B:6d} pairs={S*B:6d}  pgd_step k_fast {r[0]:7.1f} us   k_fast2 {r[1]:7.1f} us   ratio {r[0]/r[1]:.2f}")
