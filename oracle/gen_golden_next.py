"""ORACLE (test infrastructure): golden vectors for the SURVEY 8f "next" rows.  Run in the BUILD container only:

    python -m oracle.gen_golden_next       # writes tests/golden/next_*.pt

Same method as oracle/gen_golden.py: the UNMODIFIED reference (``rbi`` and, here, ``rbibm.metric``) is executed over the
third-party stubs with every base-noise draw recorded, the restatement in oracle/rbi_oracle.py must reproduce each result
BIT-FOR-BIT from the recorded noise, and inputs / noise / reference outputs are stored.

Rows covered:
  f1  forward-KL L2-PGD attack (x~ after 1/5/20 steps), ``ForwardKLRobMetric`` and ``ReverseKLRobMetric`` end to end
      (several attack attempts, best-of selection)
  f2  ``NegativeLogLikelihoodMetric``, ``ExpectedCoverageMetric`` (batched walk + |coverage - alpha| area)
  f3  ``L2PGDAdversarialTraining``, ``L2UniformNoiseTraining`` (+ ``L2UniformNoiseAttack``), ``L2PGDTrades`` (fKL)
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import rbi_oracle as O  # noqa: E402
from oracle.gen_golden import CASES, K, KW, _bitwise, _perturb_weights  # noqa: E402
from oracle.thirdparty import install_stubs  # noqa: E402

NEXT_CASES = ("tiny", "lotka_volterra", "tiny_shuffle")


class _SpyDelta0:
    """Collects the delta0 of every advertorch ``perturb_iterative`` call made inside the block."""

    def __enter__(self):
        self.d0 = []
        self._orig = O.adv.perturb_iterative

        def spy(xvar, yvar, predict, *a, **k):
            self.d0.append(k["delta_init"].detach().clone())
            return self._orig(xvar, yvar, predict, *a, **k)

        O.adv.perturb_iterative = spy
        return self

    def __exit__(self, *exc):
        O.adv.perturb_iterative = self._orig
        return False


def _grads(model):
    return [p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p) for p in model.parameters()]


def _zero(*models):
    for m in models:
        for p in m.parameters():
            p.grad = None


def build_case(name, cfg):
    import rbi.attacks as RA
    import rbi.defenses as RD
    import rbi.loss as RL
    import rbi.models as RM
    import rbibm.metric.approximation_metric as AM
    import rbibm.metric.robustness_metric as RMet
    from rbi.models.module import ZScoreLayer as RefZ

    dx, D, hidden, B = cfg["dx"], cfg["D"], cfg["hidden"], cfg["B"]
    out = {"cfg": dict(cfg, K=K, **KW)}
    torch.manual_seed(1234)
    ref = RM.InverseAffineAutoregressiveModel(dx, D, hidden_dims=hidden, num_transforms=K, shuffle=cfg["shuffle"],
                                              output_transform=None, **KW)
    _perturb_weights(ref, 99)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, dx, generator=g)
    emb = None
    if cfg["zscore"]:
        mean = 0.3 * torch.randn(dx, generator=g)
        std = 0.05 + 1.95 * torch.rand(dx, generator=g)
        x = x * std + mean
        ref.embedding_net = torch.nn.Sequential(RefZ(mean=mean, std=std), ref.embedding_net)
        out["zs_mean"], out["zs_std"] = mean, std
        emb = torch.nn.Sequential(O.ZScoreLayer(mean, std), torch.nn.Identity())
    ref.eval()
    state = {k: v.clone() for k, v in ref.state_dict().items()}
    out["state_dict"] = {k: (v.bool() if k.endswith(".mask") else v) for k, v in state.items()}
    out["x"] = x
    torch.manual_seed(4321)
    mine = O.OracleMAF(dx, D, hidden_dims=hidden, num_transforms=K, shuffle=cfg["shuffle"], output_transform=None,
                       embedding_net=emb, **KW)
    mine.load_state_dict(state, strict=True)
    mine.eval()

    eps = float(0.5 * x.std(1).mean())
    eps_iter = eps / 10
    cmin, cmax = float(x.min()), float(x.max())
    out["pgd"] = dict(eps=eps, eps_iter=eps_iter, clip_min=cmin, clip_max=cmax, S=5)

    # ---- f1a: forward-KL L2-PGD attack (the default attack loss of config/eval_rob/attack/l2pgd.yaml:5) ----------------
    for n_it in (1, 5, 20):
        atk = RA.L2PGDAttack(ref, loss_fn=RL.ForwardKLLoss(mc_samples=5), eps=eps, nb_iter=n_it, eps_iter=eps_iter,
                             clip_min=cmin, clip_max=cmax, targeted=False)
        rec = []
        torch.manual_seed(321)
        with _SpyDelta0() as spy, O.recorded_base_noise(rec):
            xa_ref = atk.perturb(x)
        with O.injected_base_noise(rec):
            xa_mine = O.l2pgd_perturb(mine, x, O.ForwardKLLoss(mc_samples=5), eps, n_it, eps_iter, cmin, cmax,
                                      delta0=spy.d0[0])
        _bitwise(xa_ref, xa_mine, f"{name}: fKL x~ after {n_it} PGD steps")
        out["pgd"]["delta0"] = spy.d0[0]
        out["pgd"][f"z_fkl_{n_it}"] = torch.stack(rec)
        out["pgd"][f"x_adv_fkl_{n_it}"] = xa_ref
        _zero(ref, mine)

    # ---- f1b: the robustness metrics end to end (3 attempts x 5 PGD iterations, 16 evaluation samples) ------------------
    for kind, ref_cls, ref_loss, my_loss in (("fkl", RMet.ForwardKLRobMetric, RL.ForwardKLLoss, O.ForwardKLLoss),
                                             ("rkl", RMet.ReverseKLRobMetric, RL.ReverseKLLoss, O.ReverseKLLoss)):
        n_it, attempts, S_eval = 5, 3, 16
        atk = RA.L2PGDAttack(ref, loss_fn=ref_loss(mc_samples=5), eps=eps, nb_iter=n_it, eps_iter=eps_iter,
                             clip_min=cmin, clip_max=cmax, targeted=False)
        met = ref_cls(ref, atk, attack_attemps=attempts, device="cpu", reduction=None, batch_size=None,
                      mc_samples=S_eval)
        rec = []
        torch.manual_seed(99)
        with _SpyDelta0() as spy, O.recorded_base_noise(rec):
            loss_ref = met.eval(x)
        xs_ref = met._cached_xs_tilde
        assert len(spy.d0) == attempts and len(rec) == attempts * (n_it + 1)
        with O.injected_base_noise(rec):
            loss_mine, xs_mine = O.rob_metric_eval(
                mine, x, my_loss(mc_samples=S_eval, reduction=None),
                lambda i: O.l2pgd_perturb(mine, x, my_loss(mc_samples=5), eps, n_it, eps_iter, cmin, cmax,
                                          delta0=spy.d0[i]),
                attack_attemps=attempts)
        _bitwise(loss_ref, loss_mine, f"{name}: {kind} robustness metric")
        _bitwise(xs_ref, xs_mine, f"{name}: {kind} robustness metric x~")
        out[f"rob_{kind}"] = dict(nb_iter=n_it, attempts=attempts, S_eval=S_eval, delta0=spy.d0,
                                  z_attack=[torch.stack(rec[i * (n_it + 1): i * (n_it + 1) + n_it]) for i in range(attempts)],
                                  z_eval=[rec[i * (n_it + 1) + n_it] for i in range(attempts)],
                                  losses=loss_ref, xs_tilde=xs_ref)
        _zero(ref, mine)

    # ---- f2: approximation metrics ----------------------------------------------------------------------------------------
    with torch.no_grad():
        theta = ref(x).sample() + 0.3 * torch.randn(B, D, generator=g) * (torch.arange(B) % 2)[:, None]
    out["theta"] = theta
    nll_ref = AM.NegativeLogLikelihoodMetric(ref, reduction=None, device="cpu").eval(x, theta)
    _bitwise(nll_ref, O.nll_metric(mine, x, theta), f"{name}: NLL metric")
    out["nll_metric"] = nll_ref
    alphas = torch.linspace(0, 1, 11)
    mc, bs = 50, (B + 1) // 2
    cov = AM.ExpectedCoverageMetric(ref, mc_samples=mc, alphas=alphas, device="cpu", batch_size=bs,
                                    reduction="absAUC_full")
    rec = []
    with torch.no_grad(), O.recorded_base_noise(rec):
        auc_ref, full_ref = cov.eval(x, theta)
    with torch.no_grad(), O.injected_base_noise(rec):
        curves = torch.vstack([O.expected_coverage(mine, xb, tb, mc, alphas)
                               for xb, tb in zip(torch.split(x, bs), torch.split(theta, bs))])
    _bitwise(auc_ref, O.coverage_abs_auc(curves, alphas), f"{name}: coverage |AUC|")
    _bitwise(full_ref["quantiles"], curves.mean(0), f"{name}: coverage curve")
    out["coverage"] = dict(mc_samples=mc, batch_size=bs, alphas=alphas, z=rec, abs_auc=auc_ref,
                           curve=full_ref["quantiles"])

    # ---- f3: defenses behind the pre-loss seam, and forward-KL TRADES -------------------------------------------------------
    theta_t = 1.2 * torch.randn(B, D, generator=g)
    out["theta_train"] = theta_t
    loss_ref = RL.NLLLoss(ref)
    tr_eps, tr_eps_iter, tr_nb = 0.5 * float(x.std(1).mean()), 0.1 * float(x.std(1).mean()), 3

    # L2PGDAdversarialTraining (attack loss fKL, S=1); the attack's inner backward pollutes p.grad
    adv_train = RD.L2PGDAdversarialTraining(ref, loss_ref, eps=tr_eps, eps_iter=tr_eps_iter, nb_iter=tr_nb)
    adv_train.activate()
    rec = []
    torch.manual_seed(2024)
    with _SpyDelta0() as spy, O.recorded_base_noise(rec):
        l = loss_ref(x, theta_t); l.backward()
    adv_train.deactivate()
    with O.injected_base_noise(rec):
        l2, x_pert = O.train_loss_with_pre_regularizer(
            mine, x, theta_t,
            lambda xi, ti: (O.l2pgd_perturb(mine, xi, O.ForwardKLLoss(mc_samples=1), tr_eps, tr_nb, tr_eps_iter,
                                            delta0=spy.d0[0]).detach(), ti))
        l2.backward()
    _bitwise(l.detach().reshape(-1), l2.detach().reshape(-1), f"{name}: adversarial-training loss")
    for a, b in zip(_grads(ref), _grads(mine)):
        _bitwise(a, b, f"{name}: adversarial-training grads (incl. attack-backward pollution)")
    entry = dict(eps=tr_eps, eps_iter=tr_eps_iter, nb_iter=tr_nb, delta0=spy.d0[0], z=torch.stack(rec),
                 loss=l.detach(), grads_polluted=_grads(ref), x_pert=x_pert)
    _zero(ref, mine)
    (-mine(x_pert).log_prob(theta_t)).mean().backward()
    entry["grads_clean"] = _grads(mine)
    out["adv_training"] = entry
    _zero(ref, mine)

    # L2UniformNoiseAttack / L2UniformNoiseTraining
    import rbi.attacks.noise_attacks as NA
    torch.manual_seed(31)
    n_ref = NA.sample_lp_uniformly(B, dx, p=2.0, eps=tr_eps)
    torch.manual_seed(31)
    _bitwise(n_ref, O.sample_l2_uniformly(B, dx, tr_eps), f"{name}: l2-ball noise")
    assert float(n_ref.norm(dim=-1).max()) <= tr_eps * (1 + 1e-5)
    captured = {}
    orig = NA.sample_lp_uniformly

    def spy_noise(*a, **k):
        captured["noise"] = orig(*a, **k)
        return captured["noise"]

    NA.sample_lp_uniformly = spy_noise
    noise_train = RD.L2UniformNoiseTraining(ref, loss_ref, eps=tr_eps, clip_min=cmin, clip_max=cmax)
    noise_train.activate()
    try:
        torch.manual_seed(32)
        l = loss_ref(x, theta_t); l.backward()
    finally:
        NA.sample_lp_uniformly = orig
        noise_train.deactivate()
    l2, x_noisy = O.train_loss_with_pre_regularizer(
        mine, x, theta_t, lambda xi, ti: (O.l2_uniform_noise_perturb(xi, tr_eps, cmin, cmax, captured["noise"]), ti))
    l2.backward()
    _bitwise(l.detach().reshape(-1), l2.detach().reshape(-1), f"{name}: noise-training loss")
    for a, b in zip(_grads(ref), _grads(mine)):
        _bitwise(a, b, f"{name}: noise-training grads")
    out["noise_training"] = dict(eps=tr_eps, clip_min=cmin, clip_max=cmax, noise=captured["noise"], x_noisy=x_noisy,
                                 loss=l.detach(), grads=_grads(ref))
    _zero(ref, mine)

    # L2PGDTrades (forward KL)
    tr_beta = 0.1
    trades = RD.L2PGDTrades(ref, loss_ref, eps=tr_eps, eps_iter=tr_eps_iter, nb_iter=tr_nb, beta=tr_beta)
    trades.activate()
    rec = []
    torch.manual_seed(778)
    with _SpyDelta0() as spy, O.recorded_base_noise(rec):
        l = loss_ref(x, theta_t); l.backward()
    trades.deactivate()
    holder = {}

    def _reg(xi, qi, ti):
        r, xp = O.trades_regularizer(mine, xi, tr_beta, tr_eps, tr_eps_iter, tr_nb, O.ForwardKLLoss, 1,
                                     delta0=spy.d0[0])
        holder["x_pert"] = xp
        return r

    with O.injected_base_noise(rec):
        l2 = O.train_loss_with_regularizer(mine, x, theta_t, _reg); l2.backward()
    _bitwise(l.detach(), l2.detach(), f"{name}: fKL-TRADES loss")
    for a, b in zip(_grads(ref), _grads(mine)):
        _bitwise(a, b, f"{name}: fKL-TRADES grads (incl. attack-backward pollution)")
    entry = dict(eps=tr_eps, eps_iter=tr_eps_iter, nb_iter=tr_nb, beta=tr_beta, delta0=spy.d0[0],
                 z=[r.clone() for r in rec], loss=l.detach(), grads_polluted=_grads(ref), x_pert=holder["x_pert"])
    _zero(ref, mine)
    with O.injected_base_noise(rec[tr_nb:]):
        nllv = -mine(x).log_prob(theta_t)
        reg = tr_beta * O.ForwardKLLoss(mc_samples=1)(mine(x), mine(holder["x_pert"]))
        (nllv.mean() + reg).backward()
    entry["reg_value"] = reg.detach()
    entry["grads_clean"] = _grads(mine)
    out["trades_fkl"] = entry
    _zero(ref, mine)
    return out


def build_checkpoint_fixture():
    """f4: a whole pickled reference model, exactly as ``save_model_by_id`` writes it (utils_data.py:688-705:
    ``torch.save(model.to("cpu"), path)``), with every feature a converter has to carry over: Permute layers, the
    z-score embedding of run_train.py:104-111 and a box output transform."""
    import rbi.models as RM
    from rbi.models.module import ZScoreLayer as RefZ
    from torch.distributions import biject_to, constraints

    dx, D, hidden = 8, 5, [24, 24]
    g = torch.Generator().manual_seed(5)
    low, high = -torch.ones(D) * 2.0, torch.ones(D) * 3.0
    out_tf = biject_to(constraints.independent(constraints.interval(low, high), 1))
    torch.manual_seed(77)
    ref = RM.InverseAffineAutoregressiveModel(dx, D, hidden_dims=hidden, num_transforms=K, shuffle=True,
                                              output_transform=out_tf, **KW)
    _perturb_weights(ref, 3)
    mean, std = 0.3 * torch.randn(dx, generator=g), 0.05 + 1.95 * torch.rand(dx, generator=g)
    ref.embedding_net = torch.nn.Sequential(RefZ(mean=mean, std=std), ref.embedding_net)
    ref.eval()
    x = torch.randn(6, dx, generator=g) * std + mean
    theta = low + (high - low) * torch.rand(3, 6, D, generator=g)
    with torch.no_grad():
        lp = ref(x).log_prob(theta)
    path = os.path.join(ROOT, "tests", "golden", "ref_checkpoint.pkl")
    torch.save(ref.to("cpu"), path)
    state = {k: (v.bool() if k.endswith(".mask") else v.clone()) for k, v in ref.state_dict().items()}
    torch.save(dict(x=x, theta=theta, log_prob=lp, state_dict=state, box_low=low, box_high=high,
                    cfg=dict(dx=dx, D=D, hidden=hidden, K=K, shuffle=True, **KW)),
               os.path.join(ROOT, "tests", "golden", "ref_checkpoint_expected.pt"))
    print(f"ref_checkpoint: ok -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def main():
    install_stubs.install_rbibm()
    torch.set_num_threads(1)
    build_checkpoint_fixture()
    for name in NEXT_CASES:
        out = build_case(name, CASES[name])
        path = os.path.join(ROOT, "tests", "golden", f"next_{name}.pt")
        torch.save(out, path)
        print(f"next_{name}: ok -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
