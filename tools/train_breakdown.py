"""Where the training-step time goes (LV shapes, batch 512): NLL with / without CUDA-graph replay, the Fisher closed form eagerly
and as the regulariser's captured graph."""
import os, sys, time, torch
sys.path.insert(0, os.getcwd())
from rabi_b200.models import InverseAffineAutoregressiveModel, ZScoreLayer
from rabi_b200.loss import NLLLoss
from rabi_b200.defenses import FIMTraceRegularizer
from rabi_b200.fisher import fisher_trace_weight_grad
dev = torch.device("cuda:0")
dx, D, hidden, K, B = 100, 4, [100, 100], 3, 512
torch.manual_seed(0)
model = InverseAffineAutoregressiveModel(dx, D, hidden_dims=hidden, num_transforms=K, shuffle=False, log_scale_min_clip=-7, log_scale_max_clip=5, stable=False)
model.embedding_net = torch.nn.Sequential(ZScoreLayer(torch.zeros(dx), torch.ones(dx)), model.embedding_net)
model = model.to(dev).train()
x, th = torch.randn(B, dx, device=dev), torch.randn(B, D, device=dev)

def timeit(fn, n=50, warm=6):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); host = (time.perf_counter() - t) / n * 1e3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return host, e0.elapsed_time(e1) / n

for graph in (True, False):
    lf = NLLLoss(model, cuda_graph=graph).train()
    def step():
        model.zero_grad(set_to_none=True); lf(x, th).backward()
    print("NLL loss + backward, cuda_graph=%s: wall %.3f ms, device-event %.3f ms" % ((graph,) + timeit(step)))
def cf():
    fisher_trace_weight_grad(model, x, 5, beta=0.05)
print("Fisher closed form (value + all weight gradients), eager: wall %.3f ms, device-event %.3f ms" % timeit(cf, 30))
for graph in (True, False):
    lf = NLLLoss(model, cuda_graph=graph).train()
    d = FIMTraceRegularizer(model, lf, beta=0.05, mc_samples=50, ema_mc_samples=5, algorithm="ema", ema_decay=0.85, cuda_graph=graph)
    d.activate()
    def step():
        model.zero_grad(set_to_none=True); lf(x, th).backward()
    print("FIM step (loss + backward), cuda_graph=%s: wall %.3f ms, device-event %.3f ms" % ((graph,) + timeit(step, 30)))
    d.deactivate()
