"""Data-parallel training step (BASELINE configs 3 / 4: FIMTraceRegularizer and L2PGDrKLTrades at lotka_volterra shapes) on N
GPUs: every rank takes a 512-row batch shard, runs loss + backward on its GPU and the gradient is all-reduced over NCCL
(rabi_b200.dist.allreduce_grads) before clip + Adam.  Launch: torchrun --nproc-per-node N tools/train_step_multi.py.
Prints one JSON line on rank 0: ms per step (max over ranks, CUDA events) with and without the all-reduce."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.getcwd())
from rabi_b200 import dist as rdist  # noqa: E402
from rabi_b200.defenses import FIMTraceRegularizer, L2PGDrKLTrades  # noqa: E402
from rabi_b200.loss import NLLLoss  # noqa: E402
from rabi_b200.models import InverseAffineAutoregressiveModel, ZScoreLayer  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
dx, D, hidden, K, B = 100, 4, [100, 100], 3, 512
torch.manual_seed(0)   # same weights on every rank
model = InverseAffineAutoregressiveModel(dx, D, hidden_dims=hidden, num_transforms=K, shuffle=False, log_scale_min_clip=-7,
                                         log_scale_max_clip=5, stable=False)
model.embedding_net = torch.nn.Sequential(ZScoreLayer(torch.zeros(dx), torch.ones(dx)), model.embedding_net)
model = model.to(dev).train()
g = torch.Generator().manual_seed(100 + rank)   # every rank its own shard
x, th = torch.randn(B, dx, generator=g).to(dev), torch.randn(B, D, generator=g).to(dev)
eps = 0.5 * float(x.std(1).mean())
out = {"n_gpus": world, "batch_per_gpu": B, "shapes": "lotka_volterra"}
for name in ("nll", "fim", "trades"):
    loss_fn = NLLLoss(model).train()
    d = None
    if name == "fim":
        d = FIMTraceRegularizer(model, loss_fn, beta=0.05, mc_samples=50, ema_mc_samples=5, algorithm="ema", ema_decay=0.85)
    elif name == "trades":
        d = L2PGDrKLTrades(model, loss_fn, eps=eps, eps_iter=eps / 5, nb_iter=20, beta=0.1)
    if d:
        d.activate()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)

    def step(reduce):
        opt.zero_grad()
        loss_fn(x, th).backward()
        if reduce:
            rdist.allreduce_grads(model.parameters())
        torch.nn.utils.clip_grad_norm_(model.parameters(), 75.0)
        opt.step()

    for reduce in (False, True):
        for _ in range(6):
            step(reduce)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            step(reduce)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 30], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out[f"{name}_ms_per_step" + ("" if reduce else "_no_allreduce")] = float(ms)
    if d:
        d.deactivate()

if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
