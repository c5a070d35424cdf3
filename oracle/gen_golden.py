"""ORACLE (test infrastructure): golden-vector generator.  Run in the BUILD container only:

    python -m oracle.gen_golden            # writes tests/golden/*.pt

It executes the UNMODIFIED reference ``rbi`` package from /root/reference/src/rbi (made importable by
``oracle/thirdparty/install_stubs``: pyro / advertorch arithmetic restated, sbi / zuko placeholders) on
seeded inputs, records every base-noise draw, and for every quantity

  1. asserts that the standalone restatement ``oracle/rbi_oracle.py`` reproduces the reference result
     BIT-FOR-BIT when fed the recorded noise (this is what pins the restatement), and
  2. stores inputs, noise and the reference's outputs as a fixture.

Quantities per case (SURVEY 8c "golden vectors to create"): log_prob, rsample, rKL (+ d/dx), x~ after
1/5/20 PGD steps, Fisher trace (+ param grads of beta*mean), TRADES regulariser (+ param grads), NLL
(+ param grads), the FIM-EMA side effect on ``p.grad`` over two steps.
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import rbi_oracle as O  # noqa: E402
from oracle.thirdparty import install_stubs  # noqa: E402

CASES = {
    # name: dx, D, hidden, B, zscore, output_transform box, train-ish weights?
    "tiny": dict(dx=6, D=3, hidden=[16, 16], B=5, zscore=False, box=False, shuffle=False),
    "gaussian_linear": dict(dx=10, D=10, hidden=[100, 100], B=6, zscore=True, box=False, shuffle=False),
    "lotka_volterra": dict(dx=100, D=4, hidden=[100, 100], B=8, zscore=True, box=False, shuffle=False),
    "pyloric_small": dict(dx=20, D=31, hidden=[64, 64, 64], B=4, zscore=True, box=True, shuffle=False),
    "tiny_shuffle": dict(dx=6, D=5, hidden=[16, 16], B=5, zscore=False, box=False, shuffle=True),
}
K = 3
KW = dict(log_scale_min_clip=-7, log_scale_max_clip=5, stable=False)  # config/model/maf_pyro.yaml:6-8


def _bitwise(a, b, what):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    same = torch.equal(a, b) or torch.equal(torch.nan_to_num(a), torch.nan_to_num(b))
    assert same, f"restatement differs from the reference for {what}: max abs {(a - b).abs().max().item():.3e}"


def _perturb_weights(model, seed, scale=0.25):
    """Default nn.Linear init gives nearly-Gaussian q; add seeded noise so log-scales/means are non-trivial."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            p.add_(scale * torch.randn(p.shape, generator=g) * (p.abs().mean() + 0.05))


def build_case(name, cfg):
    import rbi.models as RM
    import rbi.loss as RL
    import rbi.attacks as RA
    import rbi.defenses as RD
    from rbi.defenses.trades_regularizers import L2PGDrKLTrades
    from rbi.models.module import ZScoreLayer as RefZ
    from rbi.utils.fisher_info import monte_carlo_fisher
    from torch.distributions import biject_to, constraints

    dx, D, hidden, B = cfg["dx"], cfg["D"], cfg["hidden"], cfg["B"]
    out = {"cfg": dict(cfg, K=K, **KW)}

    out_tf = None
    if cfg["box"]:
        low = -torch.ones(D) * 2.0
        high = torch.ones(D) * 3.0
        out_tf = biject_to(constraints.independent(constraints.interval(low, high), 1))
        out["box_low"], out["box_high"] = low, high

    torch.manual_seed(1234)
    ref = RM.InverseAffineAutoregressiveModel(dx, D, hidden_dims=hidden, num_transforms=K,
                                              shuffle=cfg["shuffle"], output_transform=out_tf, **KW)
    _perturb_weights(ref, 99)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, dx, generator=g)
    if cfg["zscore"]:
        mean = 0.3 * torch.randn(dx, generator=g)
        std = 0.05 + 1.95 * torch.rand(dx, generator=g)
        x = x * std + mean
        ref.embedding_net = torch.nn.Sequential(RefZ(mean=mean, std=std), ref.embedding_net)  # run_train.py:104-111
        out["zs_mean"], out["zs_std"] = mean, std
    ref.eval()
    state = {k: v.clone() for k, v in ref.state_dict().items()}
    # masks are 0/1: store them as bool to keep the fixture small (tests cast back to float32)
    out["state_dict"] = {k: (v.bool() if k.endswith(".mask") else v) for k, v in state.items()}
    out["x"] = x

    # the standalone restatement, same weights
    emb = None
    if cfg["zscore"]:
        emb = torch.nn.Sequential(O.ZScoreLayer(mean, std), torch.nn.Identity())
    torch.manual_seed(4321)  # different init on purpose; the state_dict must carry everything
    mine = O.OracleMAF(dx, D, hidden_dims=hidden, num_transforms=K, shuffle=cfg["shuffle"],
                       output_transform=out_tf, embedding_net=emb, **KW)
    missing = mine.load_state_dict(state, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    mine.eval()

    # ---- log_prob of arbitrary theta ------------------------------------------------------------
    S = 3
    if cfg["box"]:
        theta = low + (high - low) * torch.rand(S, B, D, generator=g)
    else:
        theta = 1.5 * torch.randn(S, B, D, generator=g)
    with torch.no_grad():
        lp_ref = ref(x).log_prob(theta)
        lp_mine = mine(x).log_prob(theta)
    _bitwise(lp_ref, lp_mine, f"{name}: log_prob")
    out["theta"], out["log_prob"] = theta, lp_ref

    # ---- rsample + log_prob of own samples (cache-hit path, Q9) ------------------------------------
    rec = []
    with torch.no_grad(), O.recorded_base_noise(rec):
        q = ref(x)
        th_ref = q.rsample((4,))
        lps_ref = q.log_prob(th_ref)
    with torch.no_grad(), O.injected_base_noise(rec):
        q2 = mine(x)
        th_mine = q2.rsample((4,))
        lps_mine = q2.log_prob(th_mine)
    _bitwise(th_ref, th_mine, f"{name}: rsample")
    _bitwise(lps_ref, lps_mine, f"{name}: log_prob(own samples)")
    out["z_rsample"], out["rsample"], out["rsample_log_prob"] = rec[0], th_ref, lps_ref

    # ---- reverse KL (metric form, no grad) and d(mean rKL)/dx~ (attack form) ------------------------
    x_t = x + 0.2 * torch.randn(B, dx, generator=g) * (std if cfg["zscore"] else 1.0)
    out["x_tilde"] = x_t
    for S_kl in (5, 64):
        rec = []
        with torch.no_grad(), O.recorded_base_noise(rec):
            kl_ref = RL.ReverseKLLoss(mc_samples=S_kl, reduction=None)(ref(x_t), ref(x))
        with torch.no_grad(), O.injected_base_noise(rec):
            kl_mine = O.ReverseKLLoss(mc_samples=S_kl, reduction=None)(mine(x_t), mine(x))
        _bitwise(kl_ref, kl_mine, f"{name}: rKL S={S_kl}")
        out[f"z_rkl{S_kl}"], out[f"rkl{S_kl}"] = rec[0], kl_ref

    rec = []
    xt1 = x_t.clone().requires_grad_()
    with O.recorded_base_noise(rec):
        with torch.no_grad():
            y = ref(x)
        loss = RL.ReverseKLLoss(mc_samples=5)(ref(xt1), y)
        loss.backward()
    xt2 = x_t.clone().requires_grad_()
    with O.injected_base_noise(rec):
        with torch.no_grad():
            y2 = mine(x)
        loss2 = O.ReverseKLLoss(mc_samples=5)(mine(xt2), y2)
        loss2.backward()
    _bitwise(loss.detach(), loss2.detach(), f"{name}: mean rKL")
    _bitwise(xt1.grad, xt2.grad, f"{name}: d rKL / dx~")
    for p in ref.parameters():
        p.grad = None
    for p in mine.parameters():
        p.grad = None
    out["z_rkl_grad"], out["rkl_mean"], out["rkl_grad_x"] = rec[0], loss.detach(), xt1.grad.clone()

    # ---- forward KL (samples from the clean posterior) ------------------------------------------------
    rec = []
    with torch.no_grad(), O.recorded_base_noise(rec):
        fkl_ref = RL.ForwardKLLoss(mc_samples=5, reduction=None)(ref(x_t), ref(x))
    with torch.no_grad(), O.injected_base_noise(rec):
        fkl_mine = O.ForwardKLLoss(mc_samples=5, reduction=None)(mine(x_t), mine(x))
    _bitwise(fkl_ref, fkl_mine, f"{name}: fKL")
    out["z_fkl5"], out["fkl5"] = rec[0], fkl_ref

    # ---- L2-PGD, rKL loss, S=5 (run_rob_eval.py:104-153 conventions) ---------------------------------
    eps = float(0.5 * x.std(1).mean())
    nb = 20
    eps_iter = eps / 10
    cmin, cmax = float(x.min()), float(x.max())
    out["pgd"] = dict(eps=eps, eps_iter=eps_iter, clip_min=cmin, clip_max=cmax, S=5)
    for n_it in (1, 5, nb):
        atk = RA.L2PGDAttack(ref, loss_fn=RL.ReverseKLLoss(mc_samples=5), eps=eps, nb_iter=n_it,
                             eps_iter=eps_iter, clip_min=cmin, clip_max=cmax, targeted=False)
        # capture the delta0 the reference draws (advertorch rand_init_delta; NB every model(x) call also
        # consumes D normals from the global RNG via ConstantConditionalDistribution.condition,
        # utils/distributions.py:457, so delta0 cannot simply be re-drawn from the seed)
        captured = {}
        orig_pi = O.adv.perturb_iterative

        def _spy(xvar, yvar, predict, *a, **k):
            captured["d0"] = k["delta_init"].detach().clone()
            return orig_pi(xvar, yvar, predict, *a, **k)

        O.adv.perturb_iterative = _spy
        rec = []
        torch.manual_seed(555)
        try:
            with O.recorded_base_noise(rec):
                xa_ref = atk.perturb(x)
        finally:
            O.adv.perturb_iterative = orig_pi
        d0 = captured["d0"]
        with O.injected_base_noise(rec):
            xa_mine = O.l2pgd_perturb(mine, x, O.ReverseKLLoss(mc_samples=5), eps, n_it, eps_iter, cmin, cmax,
                                      delta0=d0)
        _bitwise(xa_ref, xa_mine, f"{name}: x~ after {n_it} PGD steps")
        assert float((xa_ref - x).norm(dim=-1).max()) <= eps * (1 + 1e-5)
        out["pgd"][f"delta0"] = d0
        out["pgd"][f"z_{n_it}"] = torch.stack(rec)          # [n_it, S, B, D]
        out["pgd"][f"x_adv_{n_it}"] = xa_ref
        for p in list(ref.parameters()) + list(mine.parameters()):
            p.grad = None

    # ---- training-side quantities (train mode == eval mode for this model; output_transform stripped) --
    ref.output_transform = None   # run_train.py:122-128
    mine.output_transform = None
    theta_t = 1.2 * torch.randn(B, D, generator=g)
    out["theta_train"] = theta_t

    def grads(model):
        return [p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p) for p in model.parameters()]

    def zero(model):
        for p in model.parameters():
            p.grad = None

    # NLL
    loss_ref = RL.NLLLoss(ref)
    l = loss_ref(x, theta_t); l.backward()
    l2 = O.train_loss_with_regularizer(mine, x, theta_t); l2.backward()
    _bitwise(l.detach().reshape(-1), l2.detach().reshape(-1), f"{name}: NLL")
    for a, b in zip(grads(ref), grads(mine)):
        _bitwise(a, b, f"{name}: NLL grads")
    out["nll"], out["nll_grads"] = l.detach(), grads(ref)
    zero(ref); zero(mine)

    # Fisher trace (S=5) and grad of beta*mean(trace)
    beta = 0.05
    rec = []
    with O.recorded_base_noise(rec):
        q = ref(x)
        tr_ref = monte_carlo_fisher(ref, x, q, mc_samples=5, create_graph=True, retain_graph=True, typ="trace")
        g_ref = torch.autograd.grad(beta * tr_ref.mean(), list(ref.parameters()))
    with O.injected_base_noise(rec):
        q2 = mine(x)
        tr_mine = O.mc_fisher_trace(mine, x, q2, 5)
        g_mine = torch.autograd.grad(beta * tr_mine.mean(), list(mine.parameters()))
    _bitwise(tr_ref.detach(), tr_mine.detach(), f"{name}: Fisher trace")
    for a, b in zip(g_ref, g_mine):
        _bitwise(a, b, f"{name}: Fisher trace grads")
    out["z_fim"], out["fim_trace"], out["fim_grads"] = rec[0], tr_ref.detach(), [t.detach() for t in g_ref]

    # FIMTraceRegularizer (algorithm=ema) through TrainLoss, two consecutive steps (EMA state + .grad side effect)
    defense = RD.FIMTraceRegularizer(ref, loss_ref, beta=beta, mc_samples=50, ema_mc_samples=5, algorithm="ema",
                                     ema_decay=0.85, reduce="mean")
    defense.activate()
    ema = O.EMA(0.85)
    fim_steps = []
    for step in range(2):
        rec = []
        zero(ref); zero(mine)
        with O.recorded_base_noise(rec):
            l = loss_ref(x, theta_t); l.backward()
        with O.injected_base_noise(rec):
            l2 = O.train_loss_with_regularizer(
                mine, x, theta_t,
                lambda xi, qi, ti: (O.fim_trace_regularizer_ema(mine, xi, qi, beta, ema, 5), torch.zeros(1))[1])
            l2.backward()
        _bitwise(l.detach(), l2.detach(), f"{name}: FIM-reg loss step {step}")
        for a, b in zip(grads(ref), grads(mine)):
            _bitwise(a, b, f"{name}: FIM-reg p.grad step {step}")
        fim_steps.append(dict(z=rec[0], loss=l.detach(), grads=grads(ref)))
    defense.deactivate()
    out["fim_reg_steps"] = fim_steps
    zero(ref); zero(mine)

    # TRADES (L2PGDrKLTrades), 3 inner PGD steps, S=1; inner attack backward pollutes p.grad (Q3)
    tr_eps, tr_eps_iter, tr_nb, tr_beta = 0.5 * float(x.std(1).mean()), 0.1 * float(x.std(1).mean()), 3, 0.1
    trades = L2PGDrKLTrades(ref, loss_ref, eps=tr_eps, eps_iter=tr_eps_iter, nb_iter=tr_nb, beta=tr_beta)
    trades.activate()
    captured = {}
    orig_pi = O.adv.perturb_iterative

    def _spy2(xvar, yvar, predict, *a, **k):
        captured["d0"] = k["delta_init"].detach().clone()
        return orig_pi(xvar, yvar, predict, *a, **k)

    O.adv.perturb_iterative = _spy2
    rec = []
    torch.manual_seed(777)
    try:
        with O.recorded_base_noise(rec):
            l = loss_ref(x, theta_t); l.backward()
    finally:
        O.adv.perturb_iterative = orig_pi
    d0 = captured["d0"]
    holder = {}

    def _reg(xi, qi, ti):
        r, xp = O.trades_rkl_regularizer(mine, xi, tr_beta, tr_eps, tr_eps_iter, tr_nb, 1, delta0=d0)
        holder["x_pert"] = xp
        return r

    with O.injected_base_noise(rec):
        l2 = O.train_loss_with_regularizer(mine, x, theta_t, _reg); l2.backward()
    _bitwise(l.detach(), l2.detach(), f"{name}: TRADES loss")
    for a, b in zip(grads(ref), grads(mine)):
        _bitwise(a, b, f"{name}: TRADES grads (incl. attack-backward pollution)")
    trades.deactivate()
    out["trades"] = dict(eps=tr_eps, eps_iter=tr_eps_iter, nb_iter=tr_nb, beta=tr_beta, delta0=d0,
                         z=[r.clone() for r in rec], loss=l.detach(), grads_polluted=grads(ref),
                         x_pert=holder["x_pert"])
    zero(ref); zero(mine)
    # the same regulariser without the pollution (function-level gradient: NLL + beta*KL only)
    with O.injected_base_noise(rec[tr_nb:]):
        q_clean = mine(x)
        nllv = -q_clean.log_prob(theta_t)
        reg = tr_beta * O.ReverseKLLoss(mc_samples=1)(mine(x), mine(holder["x_pert"]))
        (nllv.mean() + reg).backward()
    out["trades"]["reg_value"] = reg.detach()
    out["trades"]["grads_clean"] = grads(mine)
    zero(mine)
    return out


def main():
    install_stubs.install()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    torch.set_num_threads(1)  # bit-reproducible reductions
    for name, cfg in CASES.items():
        out = build_case(name, cfg)
        path = os.path.join(ROOT, "tests", "golden", f"{name}.pt")
        torch.save(out, path)
        print(f"{name}: ok -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
