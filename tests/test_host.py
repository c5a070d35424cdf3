"""CPU tests of the host-side mirror of the reference interface (no GPU, no kernels)."""
import copy
import io

import pytest
import torch

from oracle import rbi_oracle as O
from oracle.thirdparty import pyro_restated as pr
from tests.util import CASES, load_golden


def _model(**kw):
    from rabi_b200.models import InverseAffineAutoregressiveModel

    args = dict(hidden_dims=[16, 16], num_transforms=3, shuffle=False, log_scale_min_clip=-7, log_scale_max_clip=5,
                stable=False)
    args.update(kw)
    return InverseAffineAutoregressiveModel(6, 3, **args)


@pytest.mark.parametrize("D,C,hidden", [(3, 6, [16, 16]), (4, 100, [100, 100]), (10, 10, [100, 100]),
                                        (31, 100, [200, 200, 200]), (5, 2, [7])])
def test_made_masks_equal_restated_pyro_create_mask(D, C, hidden):
    from rabi_b200.models import made_masks

    torch.manual_seed(D)
    perm = torch.randperm(D)
    mine = made_masks(D, C, hidden, perm)
    ref, _ = pr.create_mask(D, C, hidden, perm, 2)
    assert len(mine) == len(ref)
    for a, b in zip(mine, ref):
        assert torch.equal(a, b)


def test_state_dict_layout_and_init_match_the_reference_restated():
    """Same keys, shapes AND the same RNG consumption at construction as the reference class."""
    torch.manual_seed(123)
    ref = O.OracleMAF(6, 3, hidden_dims=[16, 16], num_transforms=3, shuffle=False, log_scale_min_clip=-7,
                      log_scale_max_clip=5)
    torch.manual_seed(123)
    mine = _model()
    sr, sm = ref.state_dict(), mine.state_dict()
    assert list(sr.keys()) == list(sm.keys())
    for k in sr:
        assert torch.equal(sr[k], sm[k]), k


@pytest.mark.parametrize("name", CASES)
def test_reference_checkpoints_load(name):
    from tests.util import build_product

    g = load_golden(name)
    m = build_product(g, "cpu", with_output_transform=False)
    for k, v in g["state_dict"].items():
        assert torch.equal(m.state_dict()[k].float(), v.float()), k


def test_model_surface_deepcopy_pickle_and_output_transform_property():
    from rabi_b200.models import ZScoreLayer

    m = _model()
    assert m.input_dim == 6 and m.output_dim == 3 and m.context_dim == 6
    m.embedding_net = torch.nn.Sequential(ZScoreLayer(torch.zeros(6), torch.ones(6)), m.embedding_net)  # run_train.py:111
    t = torch.distributions.transforms.ExpTransform()
    m.output_transform = t
    assert m.output_transform is t
    m.output_transform = None
    assert m.output_transform is None
    m2 = copy.deepcopy(m)
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
    buf = io.BytesIO()
    torch.save(m, buf)                      # utils_data.py:699 pickles the whole model
    buf.seek(0)
    m3 = torch.load(buf, weights_only=False)
    assert list(m3.state_dict().keys()) == list(m.state_dict().keys())


def test_forward_returns_distribution_with_reference_batch_semantics():
    m = _model()
    for shape in [(6,), (10, 6), (1, 6), (2, 1, 3, 6), (10, 1, 6)]:
        q = m(torch.randn(*shape))
        assert isinstance(q, torch.distributions.Distribution)
        assert tuple(q.batch_shape) == shape[:-1] and tuple(q.event_shape) == (3,)
        assert q.has_rsample
        assert tuple(q.expand((4,) + shape[:-1]).batch_shape) == (4,) + shape[:-1]


def test_no_cpu_fallback_is_loud():
    from rabi_b200._lib import RabiError

    q = _model()(torch.randn(4, 6))
    with pytest.raises(RabiError):
        q.log_prob(torch.randn(4, 3))
    with pytest.raises(RabiError):
        with torch.no_grad():
            q.sample((2,))


def test_hidden_smaller_than_dim_raises_like_pyro():
    from rabi_b200.models import InverseAffineAutoregressiveModel

    with pytest.raises(ValueError, match="Hidden dimension must not be less than input dimension"):
        InverseAffineAutoregressiveModel(6, 8, hidden_dims=[4, 4], num_transforms=1, shuffle=False)


def test_zscore_embedding_is_recognised_for_fusion():
    from rabi_b200.models import ZScoreLayer, _split_embedding

    z = ZScoreLayer(torch.zeros(6), torch.ones(6))
    assert _split_embedding(torch.nn.Identity()) == (None, None)
    assert _split_embedding(torch.nn.Sequential(z, torch.nn.Identity())) == (z, None)
    assert _split_embedding(torch.nn.Sequential(z, torch.nn.Sequential(torch.nn.Identity()))) == (z, None)
    lin = torch.nn.Linear(6, 6)
    zl, rest = _split_embedding(torch.nn.Sequential(z, lin))
    assert zl is z and rest is lin
    assert _split_embedding(lin) == (None, lin)
    zo = O.ZScoreLayer(torch.zeros(6), torch.ones(6))  # the reference's own class name is recognised too
    assert _split_embedding(torch.nn.Sequential(zo, torch.nn.Identity()))[0] is zo


def test_summary_reduce_matches_reference_metric_reduce():
    from rabi_b200.metric import summary_reduce

    torch.manual_seed(0)
    m = torch.randn(1000)
    m[3] = float("nan")
    m[7] = float("inf")
    out, supp = summary_reduce(m.clone(), "mean_summary")
    f = m[torch.isfinite(m)]
    assert torch.isclose(out, f.mean())
    assert torch.isclose(supp["q50"], f.quantile(0.5)) and torch.isclose(supp["q99_5"], f.quantile(0.995))
    assert torch.isclose(supp["std"], f.std()) and set(supp) >= {"min", "max", "q00_5", "q97_5"}
    assert summary_reduce(m, None) is m


def test_loss_constraint_replaces_overflowed_kl_by_median():
    from rabi_b200.metric import ReverseKLRobMetric

    loss = torch.tensor([0.1, 0.2, -2e4, 0.4, 0.3])
    out = ReverseKLRobMetric._ensure_loss_constraint(loss.clone())
    assert out[2] == torch.tensor([0.1, 0.2, 0.4, 0.3]).median()


def test_ema_estimator_matches_reference_fixture_semantics():
    from rabi_b200.fisher import ExponentialMovingAverageEstimator

    mine, ref = ExponentialMovingAverageEstimator(0.85), O.EMA(0.85)
    torch.manual_seed(1)
    for _ in range(4):
        g = [torch.randn(3, 2), torch.randn(5)]
        g[0][0, 0] = float("nan")
        mine(g)
        ref([t.clone() for t in g])
        for a, b in zip(mine.value, ref.value):
            assert torch.allclose(a, b)


def test_noise_injection_queue():
    from rabi_b200 import noise

    z = torch.arange(12.0).reshape(2, 2, 3)
    with noise.injected([z]):
        got = noise.standard_normal((2, 2, 3), "cpu")
        assert torch.equal(got, z)
        with pytest.raises(RuntimeError):
            noise.standard_normal((1,), "cpu")
    assert noise.standard_normal((4,), "cpu").shape == (4,)


def test_l2pgd_attack_defaults_follow_the_reference_overrides():
    from rabi_b200.attacks import L2PGDAttack, rand_init_delta_l2
    from rabi_b200.loss import ReverseKLLoss

    m = _model()
    a = L2PGDAttack(m, loss_fn=ReverseKLLoss(mc_samples=5), eps=0.5)
    assert (a.clip_min, a.clip_max, a.ord, a.rand_init, a.targeted) == (-1000.0, 1000.0, 2, True, False)
    x = torch.randn(50, 6)
    d = rand_init_delta_l2(x, 0.3, -1.0, 1.5)
    assert float(d.norm(dim=1).max()) <= 0.3 * (1 + 1e-5) + float((x - x.clamp(-1.0, 1.5)).norm(dim=1).max())
    assert float((x + d).min()) >= -1.0 - 1e-6 and float((x + d).max()) <= 1.5 + 1e-6


def test_coverage_curve_equals_the_reference_level_loop():
    """``ExpectedCoverageMetric.coverage_curve`` (all levels from one sort) against the per-level restatement."""
    from rabi_b200.metric import ExpectedCoverageMetric

    torch.manual_seed(0)
    S, B = 40, 17
    alphas = torch.linspace(0, 1, 21)
    met = ExpectedCoverageMetric(model=None, mc_samples=S, alphas=alphas, device="cpu")
    logq_samples, logq_theta = torch.randn(S, B), torch.randn(B)
    mine = met.coverage_curve(logq_theta, logq_samples)
    srt, _ = torch.sort(logq_samples, dim=0)
    for i, a in enumerate(alphas):
        if float(a) in (0.0, 1.0):
            assert float(mine[i]) == float(a)
        else:
            cut = int(S * (1 - a))
            assert float(mine[i]) == float((srt[cut:].min(0).values < logq_theta).float().mean())
    # a calibrated posterior has coverage(alpha) = alpha: here theta and the samples are exchangeable draws
    cov = ExpectedCoverageMetric(None, mc_samples=200, alphas=alphas, device="cpu").coverage_curve(
        torch.randn(4000), torch.randn(200, 4000))
    assert float((cov - alphas).abs().max()) < 0.04
    val, full = ExpectedCoverageMetric(None, mc_samples=S, alphas=alphas, device="cpu",
                                       reduction="absAUC_full")._reduce(torch.stack([mine, mine]))
    assert torch.isclose(val, torch.trapz((mine - alphas).abs(), alphas)) and set(full) == {"alphas", "quantiles"}
    with pytest.raises(ValueError):
        ExpectedCoverageMetric(None, device="cpu").reduction = "mean"


def test_metric_reduction_setter_validates_like_the_reference():
    from rabi_b200.metric import Metric

    m = Metric(None, reduction="mean_quantile95", device="cpu")  # not validated at construction (rKL.yaml default)
    assert m._reduce(torch.tensor([1.0, 3.0])) == 2.0
    m.reduction = "median_summary"
    with pytest.raises(ValueError):
        m.reduction = "mean_foo"
    with pytest.raises(ValueError):
        m.reduction = "mode"


def test_pre_and_post_loss_defenses_register_on_the_train_loss():
    from rabi_b200.defenses import L2PGDAdversarialTraining, L2PGDTrades, L2UniformNoiseTraining
    from rabi_b200.loss import ForwardKLLoss, NLLLoss

    m = _model()
    lf = NLLLoss(m)
    adv = L2PGDAdversarialTraining(m, lf, eps=0.3, eps_iter=0.05, nb_iter=4)
    assert type(adv.attack.loss_fn) is ForwardKLLoss and adv.attack.loss_fn.mc_samples == 1 and adv.attack.nb_iter == 4
    noise_def = L2UniformNoiseTraining(m, lf, eps=0.2)
    trades = L2PGDTrades(m, lf, eps=0.3, beta=0.1, nb_iter=2)
    assert type(trades.reg_loss_fn) is ForwardKLLoss and trades.reg_loss_fn is not trades.attack.loss_fn
    adv.activate(); noise_def.activate(); trades.activate()
    assert set(lf.pre_loss_regularizers) == {"L2PGDAdversarialTraining", "L2UniformNoiseTraining"}
    assert set(lf.post_loss_regularizers) == {"L2PGDTrades"}
    adv.deactivate(); noise_def.deactivate(); trades.deactivate()
    assert not lf.pre_loss_regularizers and not lf.post_loss_regularizers
    x = torch.randn(64, 6)
    xn = noise_def.attack.perturb(x)  # pure elementwise torch: runs anywhere
    assert float((xn - x).norm(dim=-1).max()) <= 0.2 * (1 + 1e-5)
