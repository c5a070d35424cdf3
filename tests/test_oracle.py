"""CPU tests of the oracle: the standalone restatement reproduces the golden fixtures (outputs of the unmodified
reference, see oracle/gen_golden.py) and satisfies the known-answer properties SURVEY 8c lists."""
import math

import pytest
import torch

from oracle import rbi_oracle as O
from oracle.thirdparty import pyro_restated as pr
from tests.util import CASES, build_oracle, load_golden

torch.set_num_threads(1)


def close(a, b, tol=2e-6):
    scale = max(float(b.abs().max()), 1e-30)
    assert float((a - b).abs().max()) <= tol * scale, float((a - b).abs().max()) / scale


@pytest.mark.parametrize("name", CASES)
def test_restatement_reproduces_reference_fixture(name):
    g = load_golden(name)
    m = build_oracle(g)
    x = g["x"]
    with torch.no_grad():
        close(m(x).log_prob(g["theta"]), g["log_prob"])
        with O.injected_base_noise([g["z_rsample"]]):
            q = m(x)
            th = q.rsample((4,))
            close(th, g["rsample"])
            close(q.log_prob(th), g["rsample_log_prob"])
        with O.injected_base_noise([g["z_rkl5"]]):
            kl = O.ReverseKLLoss(mc_samples=5, reduction=None)(m(g["x_tilde"]), m(x))
        assert float((kl - g["rkl5"]).abs().max()) < 1e-5
    p = g["pgd"]
    with O.injected_base_noise(list(p["z_5"])):
        xa = O.l2pgd_perturb(m, x, O.ReverseKLLoss(mc_samples=p["S"]), p["eps"], 5, p["eps_iter"], p["clip_min"],
                             p["clip_max"], delta0=p["delta0"])
    close(xa - x, p["x_adv_5"] - x, 1e-5)


@pytest.mark.parametrize("name", ["tiny", "lotka_volterra"])
def test_restatement_training_side_fixture(name):
    g = load_golden(name)
    m = build_oracle(g, with_output_transform=False)
    x, th = g["x"], g["theta_train"]
    loss = O.train_loss_with_regularizer(m, x, th)
    loss.backward()
    close(loss.detach().reshape(()), g["nll"].reshape(()))
    for p, ref in zip(m.parameters(), g["nll_grads"]):
        close(p.grad, ref, 1e-5)
    with O.injected_base_noise([g["z_fim"]]):
        tr = O.mc_fisher_trace(m, x, m(x), 5)
    close(tr.detach(), g["fim_trace"], 1e-5)


def test_masks_are_strictly_autoregressive_in_permutation_order():
    torch.manual_seed(0)
    D, C = 5, 3
    arn = pr.ConditionalAutoRegressiveNN(D, C, [16, 16])
    ctx = torch.randn(C)
    x = torch.randn(D, requires_grad=True)
    J = torch.autograd.functional.jacobian(lambda v: torch.cat(arn(v, context=ctx)), x)  # [2D, D]
    pos = torch.empty(D, dtype=torch.long)
    pos[arn.permutation] = torch.arange(D)
    for out in range(2 * D):
        for inp in range(D):
            if pos[inp] >= pos[out % D]:
                assert J[out, inp] == 0, (out, inp)
    Jc = torch.autograd.functional.jacobian(lambda c: torch.cat(arn(x.detach(), context=c)), ctx)
    assert (Jc.abs().sum(1) > 0).all()  # every output depends on the context


def test_inverse_roundtrip_and_cached_log_scale():
    torch.manual_seed(1)
    D, C = 4, 6
    arn = pr.ConditionalAutoRegressiveNN(D, C, [20, 20])
    from functools import partial
    cond = partial(arn, context=torch.randn(7, C))
    cond.permutation = arn.permutation
    t = pr.AffineAutoregressive(cond, log_scale_min_clip=-7, log_scale_max_clip=5)
    y = torch.randn(7, D)
    x = t._inverse(y)
    ls_inv = t._cached_log_scale.clone()
    y2 = t._call(x)
    assert torch.allclose(y, y2, atol=2e-6)
    assert torch.allclose(ls_inv, t._cached_log_scale, atol=1e-6)  # last _inverse pass already has the full log-scale


def test_zero_weights_give_a_fixed_gaussian():
    """Known answer: with all weights zero the flow is an affine map of N(0, I) fixed by the biases."""
    torch.manual_seed(2)
    D, dx = 3, 5
    m = O.OracleMAF(dx, D, hidden_dims=[8, 8], num_transforms=2, shuffle=False, log_scale_min_clip=-7,
                    log_scale_max_clip=5)
    with torch.no_grad():
        for n, p in m.named_parameters():
            p.zero_()
        for k in range(2):
            b = getattr(m, f"t{k}").nn.layers[-1].bias
            b[:D] = 0.5 * (k + 1)          # mean
            b[D:] = 0.1 * (k + 1)          # log-scale
    x = torch.randn(4, dx)
    th = torch.randn(6, 4, D)
    # density pass: u1 = e^{.2} th + 1.0 ; u0 = e^{.1} u1 + .5
    u0 = math.exp(0.1) * (math.exp(0.2) * th + 1.0) + 0.5
    expected = (-0.5 * u0 ** 2 - 0.5 * math.log(2 * math.pi)).sum(-1) + D * 0.3
    with torch.no_grad():
        assert torch.allclose(m(x).log_prob(th), expected, atol=1e-5)
        kl = O.ReverseKLLoss(mc_samples=4, reduction=None)(m(x), m(x + 1.0))
    assert float(kl.abs().max()) < 1e-5  # context has no influence -> KL == 0


def test_kl_of_same_object_is_exactly_zero():
    g = load_golden("tiny")
    m = build_oracle(g)
    with torch.no_grad():
        q = m(g["x"])
        assert float(O.mc_kl(q, q, 16).abs().max()) == 0.0


def test_pgd_respects_eps_ball_and_clip_box():
    g = load_golden("tiny")
    m = build_oracle(g)
    x = g["x"]
    torch.manual_seed(0)
    xa = O.l2pgd_perturb(m, x, O.ReverseKLLoss(mc_samples=2), 0.3, 4, 0.1, -1.0, 1.5)
    assert float((xa - x.clamp(-1.0, 1.5)).norm(dim=-1).max()) <= 0.3 * 1.001 + float((x - x.clamp(-1.0, 1.5)).norm(dim=-1).max())
    assert float(xa.min()) >= -1.0 and float(xa.max()) <= 1.5


def test_ema_estimator_matches_reference_semantics():
    """decay weights the NEW value and the bias correction is 1 - (1 - decay)^t (streaming_estimators.py:44-72)."""
    e = O.EMA(0.85)
    a, b = torch.tensor([1.0, 2.0]), torch.tensor([3.0, -1.0])
    e([a])
    assert torch.allclose(e.value[0], 0.85 * a / (1 - 0.15))
    e([b])
    assert torch.allclose(e.value[0], (0.85 * b + 0.15 * 0.85 * a) / (1 - 0.15 ** 2))


# ---- SURVEY 8f "next" rows: fixtures of oracle/gen_golden_next.py ------------------------------------------------------
NEXT_CASES = ["tiny", "lotka_volterra", "tiny_shuffle"]


@pytest.mark.parametrize("name", NEXT_CASES)
def test_restatement_reproduces_next_row_fixtures(name):
    g = load_golden(f"next_{name}")
    m = build_oracle(g, with_output_transform=False)
    x, p = g["x"], g["pgd"]
    # f1: forward-KL attack and both robustness metrics (3 attempts, best-of)
    with O.injected_base_noise(list(p["z_fkl_5"])):
        xa = O.l2pgd_perturb(m, x, O.ForwardKLLoss(mc_samples=5), p["eps"], 5, p["eps_iter"], p["clip_min"],
                             p["clip_max"], delta0=p["delta0"])
    close(xa - x, p["x_adv_fkl_5"] - x, 1e-5)
    for kind, loss_cls in (("fkl", O.ForwardKLLoss), ("rkl", O.ReverseKLLoss)):
        r = g[f"rob_{kind}"]
        noise = [z for i in range(r["attempts"]) for z in (list(r["z_attack"][i]) + [r["z_eval"][i]])]
        with O.injected_base_noise(noise):
            losses, xs = O.rob_metric_eval(
                m, x, loss_cls(mc_samples=r["S_eval"], reduction=None),
                lambda i: O.l2pgd_perturb(m, x, loss_cls(mc_samples=5), p["eps"], r["nb_iter"], p["eps_iter"],
                                          p["clip_min"], p["clip_max"], delta0=r["delta0"][i]),
                attack_attemps=r["attempts"])
        close(losses, r["losses"], 1e-5)
        close(xs - x, r["xs_tilde"] - x, 1e-5)
    # f2: approximation metrics
    close(O.nll_metric(m, x, g["theta"]), g["nll_metric"])
    c = g["coverage"]
    with torch.no_grad(), O.injected_base_noise(list(c["z"])):
        curves = torch.vstack([O.expected_coverage(m, xb, tb, c["mc_samples"], c["alphas"]) for xb, tb in
                               zip(torch.split(x, c["batch_size"]), torch.split(g["theta"], c["batch_size"]))])
    close(curves.mean(0), c["curve"])
    close(O.coverage_abs_auc(curves, c["alphas"]), c["abs_auc"])
    # f3: pre-loss defenses and forward-KL TRADES
    th = g["theta_train"]
    a = g["adv_training"]
    m.zero_grad()  # the attacks above left their weight-gradient pollution behind
    with O.injected_base_noise(list(a["z"])):
        loss, x_pert = O.train_loss_with_pre_regularizer(
            m, x, th, lambda xi, ti: (O.l2pgd_perturb(m, xi, O.ForwardKLLoss(mc_samples=1), a["eps"], a["nb_iter"],
                                                      a["eps_iter"], delta0=a["delta0"]).detach(), ti))
        loss.backward()
    close(loss.detach().reshape(()), a["loss"].reshape(()))
    for prm, ref in zip(m.parameters(), a["grads_polluted"]):
        close(prm.grad, ref, 1e-5)
    m.zero_grad()
    n = g["noise_training"]
    assert float(n["noise"].norm(dim=-1).max()) <= n["eps"] * (1 + 1e-5)
    loss, x_noisy = O.train_loss_with_pre_regularizer(
        m, x, th, lambda xi, ti: (O.l2_uniform_noise_perturb(xi, n["eps"], n["clip_min"], n["clip_max"], n["noise"]), ti))
    close(loss.detach().reshape(()), n["loss"].reshape(()))
    close(x_noisy, n["x_noisy"])
    t = g["trades_fkl"]
    with O.injected_base_noise(list(t["z"])):
        reg, xp = O.trades_regularizer(m, x, t["beta"], t["eps"], t["eps_iter"], t["nb_iter"], O.ForwardKLLoss, 1,
                                       delta0=t["delta0"])
    close(xp - x, t["x_pert"] - x, 1e-5)
    close(reg.detach().reshape(()), t["reg_value"].reshape(()), 1e-5)


def test_l2_ball_noise_is_uniform_in_the_ball():
    """Known answer: for a uniform draw in the d-ball of radius eps, P(||n|| <= r) = (r/eps)^d."""
    torch.manual_seed(0)
    d, eps = 3, 0.7
    n = O.sample_l2_uniformly(20000, d, eps).norm(dim=-1)
    assert float(n.max()) <= eps * (1 + 1e-6)
    for r in (0.3, 0.5, 0.6):
        assert abs(float((n <= r).float().mean()) - (r / eps) ** d) < 0.015
    v = O.sample_l2_uniformly(20000, d, eps)
    assert float(v.mean(0).abs().max()) < 0.01


def test_closed_form_fisher_trace_gradient_equals_autograd_double_backward():
    """The derivation behind the product's Fisher-trace weight gradient (dual density pass, two masked reverse passes per
    transform, reverse of the sampling chain) against autograd's double backward on the oracle, in fp64."""
    from oracle.fim_closed_form_check import check

    assert check() < 1e-9
