import os, sys, torch, time
sys.path.insert(0, os.getcwd())
from rabi_b200.models import InverseAffineAutoregressiveModel, ZScoreLayer
from rabi_b200.loss import NLLLoss
from rabi_b200.defenses import L2PGDrKLTrades, FIMTraceRegularizer
dev = torch.device("cuda:0")
dx, D, hidden, K, B = 100, 4, [100, 100], 3, 512
torch.manual_seed(0)
model = InverseAffineAutoregressiveModel(dx, D, hidden_dims=hidden, num_transforms=K, shuffle=False, log_scale_min_clip=-7, log_scale_max_clip=5, stable=False)
model.embedding_net = torch.nn.Sequential(ZScoreLayer(torch.zeros(dx), torch.ones(dx)), model.embedding_net)
model = model.to(dev).train()
x, th = torch.randn(B, dx, device=dev), torch.randn(B, D, device=dev)
eps = 0.5 * float(x.std(1).mean())
which = sys.argv[1] if len(sys.argv) > 1 else "trades"
for name in which.split(","):
    loss_fn = NLLLoss(model).train()
    d = None
    if name == "trades":
        d = L2PGDrKLTrades(model, loss_fn, eps=eps, eps_iter=eps / 5, nb_iter=20, beta=0.1)
    elif name == "fim":
        d = FIMTraceRegularizer(model, loss_fn, beta=0.05, mc_samples=50, ema_mc_samples=5, algorithm="ema", ema_decay=0.85)
    if d: d.activate()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    def step(full=True):
        opt.zero_grad(); l = loss_fn(x, th); l.backward()
        if full:
            torch.nn.utils.clip_grad_norm_(model.parameters(), 75.0); opt.step()
    for i in range(6):
        step(); torch.cuda.synchronize(); print(name, "step", i, "ok", flush=True)
    for full in (True, False):
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(30): step(full)
        torch.cuda.synchronize(); print(f"{name}: {'whole step (loss, backward, clip, Adam)' if full else 'loss + backward only'} {(time.perf_counter() - t) / 30 * 1e3:.3f} ms", flush=True)
    if d: d.deactivate()
