# Where the host time of one NLL loss + backward goes (cProfile over 300 steps) and the device time of the same steps
import cProfile, os, pstats, sys, time, torch
sys.path.insert(0, os.getcwd())
from rabi_b200.models import InverseAffineAutoregressiveModel, ZScoreLayer
from rabi_b200.loss import NLLLoss
dev = torch.device("cuda:0")
dx, D, hidden, K, B = 100, 4, [100, 100], 3, 512
torch.manual_seed(0)
model = InverseAffineAutoregressiveModel(dx, D, hidden_dims=hidden, num_transforms=K, shuffle=False, log_scale_min_clip=-7, log_scale_max_clip=5, stable=False)
model.embedding_net = torch.nn.Sequential(ZScoreLayer(torch.zeros(dx), torch.ones(dx)), model.embedding_net)
model = model.to(dev).train()
x, th = torch.randn(B, dx, device=dev), torch.randn(B, D, device=dev)
loss_fn = NLLLoss(model).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
def step():
    opt.zero_grad(); l = loss_fn(x, th); l.backward()
for _ in range(20): step()
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(300): step()
t_host = time.perf_counter() - t
torch.cuda.synchronize()
t_all = time.perf_counter() - t
print(f"300 steps: host loop {t_host / 300 * 1e3:.3f} ms per step, with the final synchronize {t_all / 300 * 1e3:.3f} ms")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g = torch.cuda.CUDAGraph()
pr = cProfile.Profile(); pr.enable()
for _ in range(300): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
