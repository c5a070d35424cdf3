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
B = 10000
x = torch.randn(B, 100, device=dev)
zm, zs = torch.zeros(100, device=dev), torch.ones(100, device=dev)
gA = torch.randn(3, 100, B, device=dev)
z = torch.randn(5, B, 4, device=dev)
delta = torch.zeros_like(x)
def t(fn, n=200):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
A = eng.ctx_project(x, zm, zs)
print("ctx_project      %.1f us" % t(lambda: eng.ctx_project(x, zm, zs)))
print("ctx_project_bwd  %.1f us" % t(lambda: eng.ctx_project_bwd(gA, zs)))
print("rkl_input_grad   %.1f us (ctx_project + k_fast + ctx_bwd + copy)" % t(lambda: eng.rkl_input_grad(x, zm, zs, A, z, 1e-3, want_kl=False)))
print("pgd_step         %.1f us" % t(lambda: eng.pgd_step(x, delta, zm, zs, A, z, 0.5, 0.05, -3., 3., 1e-3)))
ze = torch.randn(256, B, 4, device=dev)
print("rkl eval S=256   %.1f us" % t(lambda: eng.rkl(A, A, ze), n=5))
