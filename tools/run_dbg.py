"""Section timing of the persistent attack kernel (RABI_TC_RUN_DBG=1 prints CTA 0's cycles per iteration to stderr)."""
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
os.environ.setdefault("RABI_TC_RUN_DBG", "0")
import bench  # noqa: E402
from rabi_b200.models import ZScoreLayer  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "lotka_volterra_l2pgd_rkl"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
N_IT = int(sys.argv[3]) if len(sys.argv) > 3 else 20
w = dict(bench.WORKLOADS[wl])
dx, D = w["dx"], w["D"]
dev = torch.device("cuda:0")
model = bench.make_weights(w)
model.embedding_net = torch.nn.Sequential(ZScoreLayer(torch.zeros(dx), torch.ones(dx)), model.embedding_net)
model = model.to(dev).eval()
eng = model.engine()
zm, zs = torch.zeros(dx, device=dev), torch.ones(dx, device=dev)
x = torch.randn(B, dx, device=dev)
for n_it in (N_IT, N_IT):
    z = torch.randn(n_it, 5, B, D, device=dev)
    delta = torch.zeros_like(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.pgd_run(x, delta, zm, zs, None, z, 0.5, 0.05, -3.0, 3.0, 1.0 / 1000)
    e1.record()
    torch.cuda.synchronize()
    print("ms per iteration", e0.elapsed_time(e1) / n_it)
