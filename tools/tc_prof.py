"""One attack gradient launch group at config-2 size on the tensor-core path (profiling target for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
import bench  # noqa: E402
from rabi_b200.models import ZScoreLayer  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "lotka_volterra_l2pgd_rkl"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
n = int(sys.argv[3]) if len(sys.argv) > 3 else 8
w = dict(bench.WORKLOADS[wl])
dx, D = w["dx"], w["D"]
dev = torch.device("cuda:0")
model = bench.make_weights(w)
model.embedding_net = torch.nn.Sequential(ZScoreLayer(torch.zeros(dx), torch.ones(dx)), model.embedding_net)
model = model.to(dev).eval()
eng = model.engine()
zm, zs = torch.zeros(dx, device=dev), torch.ones(dx, device=dev)
x = torch.randn(B, dx, device=dev)
z = torch.randn(5, B, D, device=dev)
delta = torch.zeros_like(x)
A = eng.ctx_project(x, zm, zs)
for _ in range(n):
    eng.pgd_step(x, delta, zm, zs, A, z, 0.5, 0.05, -3.0, 3.0, 1.0 / 1000)
torch.cuda.synchronize()
print("ok", float(delta.abs().sum()))
