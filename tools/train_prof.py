"""A few FIM-regularised training steps at lotka_volterra shapes, batch 512, eagerly (profiling target for ncu)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from rabi_b200.models import InverseAffineAutoregressiveModel, ZScoreLayer
from rabi_b200.loss import NLLLoss
from rabi_b200.defenses import FIMTraceRegularizer
dev = torch.device("cuda:0")
dx, D, hidden, K, B = 100, 4, [100, 100], 3, 512
torch.manual_seed(0)
model = InverseAffineAutoregressiveModel(dx, D, hidden_dims=hidden, num_transforms=K, shuffle=False, log_scale_min_clip=-7, log_scale_max_clip=5, stable=False)
model.embedding_net = torch.nn.Sequential(ZScoreLayer(torch.zeros(dx), torch.ones(dx)), model.embedding_net)
model = model.to(dev).train()
x, th = torch.randn(B, dx, device=dev), torch.randn(B, D, device=dev)
lf = NLLLoss(model, cuda_graph=False).train()
d = FIMTraceRegularizer(model, lf, beta=0.05, mc_samples=50, ema_mc_samples=5, algorithm="ema", ema_decay=0.85, cuda_graph=False)
d.activate()
for _ in range(3):
    model.zero_grad(set_to_none=True)
    lf(x, th).backward()
torch.cuda.synchronize()
print("ok")
