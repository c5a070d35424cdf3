"""Training-side configs 3 / 4 (LV shapes, batch 512): step time of rabi_b200 vs the oracle (reference restatement)
on CPU and eagerly on the same GPU."""
import os, sys, time, torch
sys.path.insert(0, os.getcwd())
from oracle import rbi_oracle as O
from rabi_b200.models import InverseAffineAutoregressiveModel, ZScoreLayer
from rabi_b200.loss import NLLLoss
from rabi_b200.defenses import FIMTraceRegularizer, L2PGDrKLTrades
dev = torch.device("cuda:0")
dx, D, hidden, K, B = 100, 4, [100, 100], 3, 512
kw = dict(log_scale_min_clip=-7, log_scale_max_clip=5, stable=False)
torch.manual_seed(0)
mean, std = torch.zeros(dx), torch.ones(dx)
def mk_ref(device):
    torch.manual_seed(0)
    m = O.OracleMAF(dx, D, hidden_dims=hidden, num_transforms=K, shuffle=False,
                    embedding_net=torch.nn.Sequential(O.ZScoreLayer(mean.clone(), std.clone()), torch.nn.Identity()), **kw)
    return m.to(device).train()
ref_cpu = mk_ref("cpu")
model = InverseAffineAutoregressiveModel(dx, D, hidden_dims=hidden, num_transforms=K, shuffle=False, **kw)
model.embedding_net = torch.nn.Sequential(ZScoreLayer(mean.clone(), std.clone()), model.embedding_net)
model.load_state_dict(ref_cpu.state_dict()); model = model.to(dev).train()
x, th = torch.randn(B, dx), torch.randn(B, D)
eps = 0.5 * float(x.std(1).mean())

def timeit(fn, n, sync, warm=1):
    for _ in range(warm): fn()
    sync()
    t = time.perf_counter()
    for _ in range(n): fn()
    sync()
    return (time.perf_counter() - t) / n * 1e3

def mine_step(defense):
    loss_fn = NLLLoss(model).train()
    d = None
    if defense == "fim":
        d = FIMTraceRegularizer(model, loss_fn, beta=0.05, mc_samples=50, ema_mc_samples=5, algorithm="ema", ema_decay=0.85)
    elif defense == "trades":
        d = L2PGDrKLTrades(model, loss_fn, eps=eps, eps_iter=eps / 5, nb_iter=20, beta=0.1)
    if d: d.activate()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    xd, td = x.to(dev), th.to(dev)
    def step():
        opt.zero_grad(); l = loss_fn(xd, td); l.backward(); torch.nn.utils.clip_grad_norm_(model.parameters(), 75.0); opt.step()
    ms = timeit(step, 20, torch.cuda.synchronize, warm=5)
    if d: d.deactivate()
    return ms

def ref_step(ref, defense, device, n):
    ema = O.EMA(0.85)
    xd, td = x.to(device), th.to(device)
    opt = torch.optim.Adam(ref.parameters(), lr=1e-4)
    def reg(xi, qi, ti):
        if defense == "fim":
            O.fim_trace_regularizer_ema(ref, xi, qi, 0.05, ema, 5); return torch.zeros(1, device=device)
        r, _ = O.trades_rkl_regularizer(ref, xi, 0.1, eps, eps / 5, 20, 1); return r
    def step():
        opt.zero_grad(); l = O.train_loss_with_regularizer(ref, xd, td, None if defense == "nll" else reg) if device == "cpu" else _gpu_loss(ref, xd, td, None if defense == "nll" else reg)
        l.backward(); torch.nn.utils.clip_grad_norm_(ref.parameters(), 75.0); opt.step()
    return timeit(step, n, (torch.cuda.synchronize if device != "cpu" else (lambda: None)))

def _gpu_loss(ref, xd, td, regularizer):
    q = ref(xd); loss = -q.log_prob(td)
    post = torch.zeros(1, device=xd.device)
    if regularizer is not None: post = post + regularizer(xd, q, td)
    return loss.mean() + (post - post.detach())

torch.set_num_threads(os.cpu_count())
print(f"LV shapes, batch {B}, host threads {os.cpu_count()}")
for defense in ("nll", "fim", "trades"):
    m = mine_step(defense)
    rc = ref_step(ref_cpu, defense, "cpu", 2 if defense != "nll" else 5)
    try:
        rg = ref_step(mk_ref(dev), defense, dev, 3)
    except Exception as e:
        rg = float("nan"); print("oracle on GPU failed:", str(e)[:100])
    print(f"{defense:7s} rabi_b200 {m:8.1f} ms/step | oracle CPU {rc:9.1f} ms/step | oracle eager on the same B200 {rg:8.1f} ms/step")
