"""GPU parity tests (run on the B200 with ``-m gpu``): the sm_100a kernels, driven through the C ABI, against
(a) the committed golden fixtures -- outputs of the UNMODIFIED reference executed in the build container -- and
(b) the live CPU oracle on fresh seeded inputs, at the reference's shapes."""
import pytest
import torch

from tests.util import (CASES, RTOL, assert_trajectories_match, build_oracle, build_product, load_golden,
                        pgd_near_kink_rows, rel_err, relu_margins)
from tests.util import TAU, assert_kl_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(scope="module", params=CASES)
def case(request, dev):
    g = load_golden(request.param)
    return request.param, g, build_product(g, dev)


def test_log_prob_matches_reference(case, dev):
    name, g, model = case
    with torch.no_grad():
        lp = model(g["x"].to(dev)).log_prob(g["theta"].to(dev))
    assert lp.shape == g["log_prob"].shape
    assert rel_err(lp, g["log_prob"]) < RTOL


def test_rsample_and_own_log_prob_match_reference(case, dev):
    from rabi_b200 import noise

    name, g, model = case
    with torch.no_grad(), noise.injected([g["z_rsample"]]):
        q = model(g["x"].to(dev))
        th = q.rsample((4,))
        lp_own = q.log_prob(th)          # "free" log-prob of the samples just drawn (reference cache hit, Q9)
        lp_again = q.log_prob(th.clone())  # recomputed through the density pass
    assert rel_err(th, g["rsample"]) < RTOL
    assert rel_err(lp_own, g["rsample_log_prob"]) < RTOL
    assert rel_err(lp_again, g["rsample_log_prob"]) < RTOL


@pytest.mark.parametrize("S", [5, 64])
def test_reverse_kl_matches_reference(case, dev, S):
    from rabi_b200 import noise
    from rabi_b200.loss import ReverseKLLoss

    name, g, model = case
    with torch.no_grad(), noise.injected([g[f"z_rkl{S}"]]):
        kl = ReverseKLLoss(mc_samples=S, reduction=None)(model(g["x_tilde"].to(dev)), model(g["x"].to(dev)))
    ref = g[f"rkl{S}"]
    assert_kl_close(kl, ref, g["rsample_log_prob"].abs().max(), f"{name}: rKL S={S}")


def test_forward_kl_matches_reference(case, dev):
    from rabi_b200 import noise
    from rabi_b200.loss import ForwardKLLoss

    name, g, model = case
    with torch.no_grad(), noise.injected([g["z_fkl5"]]):
        kl = ForwardKLLoss(mc_samples=5, reduction=None)(model(g["x_tilde"].to(dev)), model(g["x"].to(dev)))
    assert_kl_close(kl, g["fkl5"], g["rsample_log_prob"].abs().max(), f"{name}: fKL S=5")


def test_reverse_kl_input_gradient_matches_reference(case, dev):
    """d mean_B KL(q(.|x~) || q(.|x)) / d x~ -- what perturb_iterative's loss.backward() leaves in delta.grad."""
    name, g, model = case
    x, xt = g["x"].to(dev), g["x_tilde"].to(dev)
    eng = model.engine()
    with torch.no_grad():
        q = model(x)
        _, zs, _ = model.condition_inputs(xt)
        zs = zs or (None, None)
        gx, kl = eng.rkl_input_grad(xt, zs[0], zs[1], q.A, g["z_rkl_grad"].to(dev), 1.0 / x.shape[0])
    assert rel_err(gx, g["rkl_grad_x"]) < RTOL
    assert_kl_close(kl.mean(), g["rkl_mean"], g["rsample_log_prob"].abs().max(), f"{name}: mean rKL of the attack loss")


@pytest.mark.parametrize("n_it", [1, 5, 20])
def test_l2pgd_matches_reference(case, dev, n_it):
    from rabi_b200 import noise
    from rabi_b200.attacks import L2PGDAttack
    from rabi_b200.loss import ReverseKLLoss

    name, g, model = case
    p = g["pgd"]
    x = g["x"].to(dev)
    atk = L2PGDAttack(model, loss_fn=ReverseKLLoss(mc_samples=p["S"]), eps=p["eps"], nb_iter=n_it,
                      eps_iter=p["eps_iter"], clip_min=p["clip_min"], clip_max=p["clip_max"], targeted=False)
    with noise.injected(list(p[f"z_{n_it}"])):
        xa = atk.perturb(x, delta_init=p["delta0"].to(dev))
    ref = p[f"x_adv_{n_it}"]
    # compare the perturbations (x~ - x), relative to their size
    # a row may leave the tolerance only with the fp64 proof of a near-kink event on its trajectory
    prone = pgd_near_kink_rows(build_oracle(g), g["x"], p["delta0"], list(p[f"z_{n_it}"]), p)
    e = assert_trajectories_match(xa.cpu() - g["x"], ref - g["x"], RTOL, prone=prone)
    print(f"\n[{name}] x~ after {n_it} PGD steps: max row rel. err {float(e.max()):.2e}, near-kink rows {int(prone.sum())}/{e.numel()}")
    assert float((xa.cpu() - g["x"]).norm(dim=-1).max()) <= p["eps"] * (1 + 1e-5)
    assert float(xa.min()) >= p["clip_min"] - 1e-6 and float(xa.max()) <= p["clip_max"] + 1e-6


# ------------------------------------------------------------------------------------------------------------
# live oracle at the reference's full shapes (fresh seeded inputs; sizes the oracle finishes in seconds)
SHAPES = {
    "gaussian_linear": dict(dx=10, D=10, hidden=[100, 100], B=64),
    "lotka_volterra": dict(dx=100, D=4, hidden=[100, 100], B=96),
    "pyloric": dict(dx=100, D=31, hidden=[200, 200, 200], B=24),
}


def _fresh(shape, dev, seed=0):
    from oracle import rbi_oracle as O
    from rabi_b200.models import InverseAffineAutoregressiveModel, ZScoreLayer

    torch.manual_seed(seed)
    kw = dict(log_scale_min_clip=-7, log_scale_max_clip=5, stable=False)
    dx, D, hidden, B = shape["dx"], shape["D"], shape["hidden"], shape["B"]
    mean, std = 0.3 * torch.randn(dx), 0.05 + 1.95 * torch.rand(dx)
    ref = O.OracleMAF(dx, D, hidden_dims=hidden, num_transforms=3, shuffle=False,
                      embedding_net=torch.nn.Sequential(O.ZScoreLayer(mean, std), torch.nn.Identity()), **kw)
    with torch.no_grad():
        for p in ref.parameters():
            p.add_(0.3 * torch.randn_like(p) * (p.abs().mean() + 0.05))
    model = InverseAffineAutoregressiveModel(dx, D, hidden_dims=hidden, num_transforms=3, shuffle=False, **kw)
    model.embedding_net = torch.nn.Sequential(ZScoreLayer(mean, std), model.embedding_net)
    model.load_state_dict(ref.state_dict())
    x = torch.randn(B, dx) * std + mean
    return ref.eval(), model.to(dev).eval(), x


@pytest.mark.parametrize("shape", list(SHAPES))
def test_full_shape_pipeline_vs_oracle(shape, dev):
    from oracle import rbi_oracle as O
    from rabi_b200 import noise
    from rabi_b200.attacks import L2PGDAttack
    from rabi_b200.loss import ReverseKLLoss

    ref, model, x = _fresh(SHAPES[shape], dev)
    B, D = x.shape[0], SHAPES[shape]["D"]
    eps = float(0.5 * x.std(1).mean())
    cmin, cmax = float(x.min()), float(x.max())
    nb, S, S_eval = 10, 5, 32
    g = torch.Generator().manual_seed(3)
    zs = [torch.randn(S, B, D, generator=g) for _ in range(nb)]
    z_eval = torch.randn(S_eval, B, D, generator=g)
    d0 = O.adv.clamp_by_pnorm(torch.randn(B, x.shape[1], generator=g), 2, eps)
    d0 = O.adv.clamp(x + d0, cmin, cmax) - x
    with O.injected_base_noise(zs):
        xa_ref = O.l2pgd_perturb(ref, x, O.ReverseKLLoss(mc_samples=S), eps, nb, eps / 10, cmin, cmax, delta0=d0)
    with O.injected_base_noise([z_eval]):
        kl_ref = O.rkl_rob_metric(ref, x, xa_ref, S_eval)
    atk = L2PGDAttack(model, loss_fn=ReverseKLLoss(mc_samples=S), eps=eps, nb_iter=nb, eps_iter=eps / 10,
                      clip_min=cmin, clip_max=cmax)
    with noise.injected(zs):
        xa = atk.perturb(x.to(dev), delta_init=d0.to(dev))
    prone = pgd_near_kink_rows(ref, x, d0, zs, dict(eps=eps, eps_iter=eps / 10, clip_min=cmin, clip_max=cmax, S=S))
    assert_trajectories_match(xa.cpu() - x, xa_ref - x, RTOL, prone=prone)
    with torch.no_grad(), noise.injected([z_eval]):
        kl = ReverseKLLoss(mc_samples=S_eval, reduction=None)(model(xa_ref.to(dev)), model(x.to(dev)))
    with torch.no_grad():
        lp_scale = float(ref(x).log_prob(ref(x).sample()).abs().max())
    assert_kl_close(kl, kl_ref, lp_scale, f"{shape}: rKL metric S={S_eval} at full shape")


def test_size_independent_properties_at_config2_scale(dev):
    """BASELINE config 2 sizes (LV shapes, B=1000 chunk, S=256): properties that need no oracle run.
    KL(q,q) == 0 exactly when both arguments are the same conditioned object; ||x~ - x|| <= eps and inside the clip
    box; log_prob(rsample) through the density pass equals the sampling pass' own log-density."""
    from rabi_b200.attacks import L2PGDAttack
    from rabi_b200.loss import ReverseKLLoss

    _, model, _ = _fresh(SHAPES["lotka_volterra"], dev, seed=5)
    torch.manual_seed(11)
    x = torch.randn(1000, 100, device=dev)
    with torch.no_grad():
        q = model(x)
        kl_same = ReverseKLLoss(mc_samples=256, reduction=None)(q, q)
        assert float(kl_same.abs().max()) < 2e-4  # same weights, same context: only fp32 re-association noise
        th = q.rsample((64,))
        lp_free = q.log_prob(th)
        lp_dens = q.log_prob(th.clone())
        assert torch.isfinite(lp_free).all()
        assert float((lp_free - lp_dens).abs().max()) < 1e-4 * float(lp_free.abs().max())
    eps = float(0.5 * x.std(1).mean())
    atk = L2PGDAttack(model, loss_fn=ReverseKLLoss(mc_samples=5), eps=eps, nb_iter=20, eps_iter=eps / 10,
                      clip_min=float(x.min()), clip_max=float(x.max()))
    xa = atk.perturb(x)
    assert xa.shape == x.shape and (xa != x).any()
    assert float((xa - x).norm(dim=-1).max()) <= eps * (1 + 1e-5)
    assert float(xa.min()) >= float(x.min()) and float(xa.max()) <= float(x.max())
    with torch.no_grad():
        kl_adv = ReverseKLLoss(mc_samples=256, reduction=None)(model(xa), model(x))
        kl_rand = ReverseKLLoss(mc_samples=256, reduction=None)(
            model(x + eps * torch.nn.functional.normalize(torch.randn_like(x), dim=-1)), model(x))
    assert float(kl_adv.mean()) > float(kl_rand.mean())  # the attack beats a random direction of the same norm


@pytest.mark.parametrize("shape", ["gaussian_linear", "lotka_volterra", "pyloric"])
def test_forward_kl_attack_vs_oracle(shape, dev):
    """ForwardKLLoss as the attack loss (the default of config/eval_rob/attack/l2pgd.yaml):
    gradient of mean_B KL(q(.|x) || q(.|x~)) w.r.t. x~ and the L2-PGD trajectory, against the live oracle."""
    from oracle import rbi_oracle as O
    from rabi_b200 import noise
    from rabi_b200.attacks import L2PGDAttack
    from rabi_b200.loss import ForwardKLLoss

    ref, model, x = _fresh(SHAPES[shape], dev, seed=2)
    B, D = x.shape[0], SHAPES[shape]["D"]
    S, nb = 5, 8
    eps = float(0.5 * x.std(1).mean())
    cmin, cmax = float(x.min()), float(x.max())
    g = torch.Generator().manual_seed(9)
    zs = [torch.randn(S, B, D, generator=g) for _ in range(nb)]
    xt = x + 0.2 * eps * torch.randn(B, x.shape[1], generator=g)
    xt_ = xt.clone().requires_grad_()
    with O.injected_base_noise([zs[0]]):
        with torch.no_grad():
            y = ref(x)
        O.ForwardKLLoss(mc_samples=S)(ref(xt_), y).backward()
    with torch.no_grad():
        q = model(x.to(dev))
        _, zsc, _ = model.condition_inputs(xt.to(dev))
        gx, _ = model.engine().rkl_input_grad(xt.to(dev), zsc[0], zsc[1], q.A, zs[0].to(dev), 1.0 / B, forward_kl=True)
    assert rel_err(gx, xt_.grad) < RTOL
    d0 = O.adv.clamp_by_pnorm(torch.randn(B, x.shape[1], generator=g), 2, eps)
    d0 = O.adv.clamp(x + d0, cmin, cmax) - x
    with O.injected_base_noise(zs):
        xa_ref = O.l2pgd_perturb(ref, x, O.ForwardKLLoss(mc_samples=S), eps, nb, eps / 10, cmin, cmax, delta0=d0)
    atk = L2PGDAttack(model, loss_fn=ForwardKLLoss(mc_samples=S), eps=eps, nb_iter=nb, eps_iter=eps / 10,
                      clip_min=cmin, clip_max=cmax)
    with noise.injected(zs):
        xa = atk.perturb(x.to(dev), delta_init=d0.to(dev))
    prone = pgd_near_kink_rows(ref, x, d0, zs, dict(eps=eps, eps_iter=eps / 10, clip_min=cmin, clip_max=cmax, S=S), forward_kl=True)
    assert_trajectories_match(xa.cpu() - x, xa_ref - x, RTOL, prone=prone)


def test_edge_cases_empty_ragged_and_large_sample_counts(dev):
    """Empty batches, S = 1, row counts that are not multiples of the tile / vector sizes, more samples per row than a
    tile holds (falls back to the generic kernel), and equality of the fast and the generic kernels."""
    import os

    from rabi_b200 import noise
    from rabi_b200.attacks import L2PGDAttack
    from rabi_b200.loss import ReverseKLLoss

    ref, model, x = _fresh(dict(dx=10, D=4, hidden=[24, 24], B=37), dev, seed=4)
    xd = x.to(dev)
    with torch.no_grad():
        q = model(xd)
        # empty sample shape and empty batch
        assert q.rsample((0,)).shape == (0, 37, 4)
        assert model(xd[:0]).log_prob(torch.zeros(3, 0, 4, device=dev)).shape == (3, 0)
        assert ReverseKLLoss(mc_samples=4, reduction=None)(model(xd[:0]), model(xd[:0])).shape == (0,)
        # S = 1 and a ragged row count, fast vs generic kernel
        g = torch.Generator().manual_seed(0)
        for S in (1, 3, 700):
            z = torch.randn(S, 37, 4, generator=g)
            outs = []
            for flag in ("0", "1"):
                os.environ["RABI_DISABLE_FAST"] = flag
                with noise.injected([z, z]):
                    th = q.rsample((S,))
                    lp = q.log_prob(th.clone())
                    kl = ReverseKLLoss(mc_samples=S, reduction=None)(model(xd + 0.1), q)
                outs.append((th, lp, kl))
            os.environ["RABI_DISABLE_FAST"] = "0"
            (th_f, lp_f, kl_f), (th_g, lp_g, kl_g) = outs
            assert rel_err(th_f, th_g) < 1e-5 and rel_err(lp_f, lp_g) < 1e-5
            # KL values are differences of O(10) log-densities: compare on that scale
            assert float((kl_f - kl_g).abs().max()) < 1e-5 * float(lp_g.abs().max())
            with O_injected([z]):
                th_ref = ref(x).rsample((S,))
            assert rel_err(outs[0][0], th_ref) < RTOL
    # an attack whose S exceeds one tile's capacity still works (generic kernel) and respects the constraints
    atk = L2PGDAttack(model, loss_fn=ReverseKLLoss(mc_samples=600), eps=0.3, nb_iter=2, eps_iter=0.1, clip_min=-2.0,
                      clip_max=2.0)
    xa = atk.perturb(xd[:3])
    assert float((xa - xd[:3].clamp(-2, 2)).norm(dim=-1).max()) <= 0.3 * 1.001 + float((xd[:3] - xd[:3].clamp(-2, 2)).norm(dim=-1).max())


def O_injected(noise_list):
    from oracle import rbi_oracle as O

    return O.injected_base_noise(noise_list)


def test_sharded_equals_chunked_attack(dev):
    """attacking all rows in one call with chunk_size=c gives every row the trajectory it gets when
    the rows are attacked in independent chunks of c rows (the reference's batched walk)."""
    from rabi_b200 import noise
    from rabi_b200.attacks import L2PGDAttack
    from rabi_b200.loss import ReverseKLLoss

    _, model, x = _fresh(SHAPES["lotka_volterra"], dev, seed=6)
    xd = x.to(dev)[:64]
    B, D, S, nb, c = 64, 4, 5, 6, 16
    eps = float(0.5 * xd.std(1).mean())
    g = torch.Generator().manual_seed(5)
    zs = [torch.randn(S, B, D, generator=g) for _ in range(nb)]
    d0 = (0.1 * eps * torch.randn(B, xd.shape[1], generator=g)).to(dev)
    atk = L2PGDAttack(model, loss_fn=ReverseKLLoss(mc_samples=S), eps=eps, nb_iter=nb, eps_iter=eps / 10,
                      clip_min=-4.0, clip_max=4.0)
    with noise.injected(zs):
        whole = atk.perturb(xd, delta_init=d0, chunk_size=c)
    parts = []
    for i in range(0, B, c):
        with noise.injected([z[:, i:i + c].contiguous() for z in zs]):
            parts.append(atk.perturb(xd[i:i + c], delta_init=d0[i:i + c]))
    assert_trajectories_match(whole.cpu() - xd.cpu(), torch.cat(parts).cpu() - xd.cpu(), 1e-5)


def test_l2pgd_through_a_nonidentity_embedding_vs_oracle(dev):
    """x-space PGD with a learned embedding net in front of the flow (pyloric-style: z-score -> network -> context):
    the gradient is chained through the torch embedding, everything else runs on the kernels."""
    from oracle import rbi_oracle as O
    from rabi_b200 import noise
    from rabi_b200.attacks import L2PGDAttack
    from rabi_b200.loss import ReverseKLLoss
    from rabi_b200.models import InverseAffineAutoregressiveModel, ZScoreLayer

    torch.manual_seed(8)
    dx, C, D, B, S, nb = 40, 12, 5, 24, 5, 6
    kw = dict(log_scale_min_clip=-7, log_scale_max_clip=5, stable=False)
    mean, std = 0.2 * torch.randn(dx), 0.5 + torch.rand(dx)

    def emb(zcls):
        torch.manual_seed(3)
        return torch.nn.Sequential(zcls(mean.clone(), std.clone()), torch.nn.Linear(dx, 32), torch.nn.Tanh(),
                                   torch.nn.Linear(32, C))

    ref = O.OracleMAF(dx, D, hidden_dims=[32, 32], num_transforms=3, shuffle=False, embedding_net=emb(O.ZScoreLayer), **kw)
    with torch.no_grad():
        for n, p in ref.named_parameters():
            if n.startswith("t"):
                p.add_(0.3 * torch.randn_like(p) * (p.abs().mean() + 0.05))
    model = InverseAffineAutoregressiveModel(dx, D, hidden_dims=[32, 32], num_transforms=3, shuffle=False,
                                             embedding_net=emb(ZScoreLayer), **kw)
    model.load_state_dict(ref.state_dict())
    ref, model = ref.eval(), model.to(dev).eval()
    assert model.context_dim == C
    x = torch.randn(B, dx) * std + mean
    eps = float(0.5 * x.std(1).mean())
    cmin, cmax = float(x.min()), float(x.max())
    g = torch.Generator().manual_seed(1)
    zs = [torch.randn(S, B, D, generator=g) for _ in range(nb)]
    d0 = O.adv.clamp_by_pnorm(torch.randn(B, dx, generator=g), 2, eps)
    d0 = O.adv.clamp(x + d0, cmin, cmax) - x
    with O.injected_base_noise(zs):
        xa_ref = O.l2pgd_perturb(ref, x, O.ReverseKLLoss(mc_samples=S), eps, nb, eps / 10, cmin, cmax, delta0=d0)
    atk = L2PGDAttack(model, loss_fn=ReverseKLLoss(mc_samples=S), eps=eps, nb_iter=nb, eps_iter=eps / 10,
                      clip_min=cmin, clip_max=cmax)
    with noise.injected(zs):
        xa = atk.perturb(x.to(dev), delta_init=d0.to(dev))
    prone = pgd_near_kink_rows(ref, x, d0, zs, dict(eps=eps, eps_iter=eps / 10, clip_min=cmin, clip_max=cmax, S=S))
    assert_trajectories_match(xa.cpu() - x, xa_ref - x, RTOL, prone=prone)
    with torch.no_grad(), O.injected_base_noise([zs[0]]):
        kl_ref = O.ReverseKLLoss(mc_samples=S, reduction=None)(ref(xa_ref), ref(x))
    with torch.no_grad(), noise.injected([zs[0]]):
        kl = ReverseKLLoss(mc_samples=S, reduction=None)(model(xa_ref.to(dev)), model(x.to(dev)))
    assert float((kl.cpu() - kl_ref).abs().max()) < 1e-3


# ------------------------------------------------------------------------------------------------------------
# deep path (maf_deep.cuh: per-pair state in a global-memory scratch, one launch per (transform, pass)); it serves the
# models the fast path cannot hold on chip (pyloric_small above already runs on it) and can be forced for any model
@pytest.fixture()
def force_deep():
    import os

    os.environ["RABI_FORCE_DEEP"] = "1"
    yield
    os.environ.pop("RABI_FORCE_DEEP", None)


def test_deep_path_matches_reference_fixtures(case, dev, force_deep):
    from rabi_b200 import noise
    from rabi_b200.attacks import L2PGDAttack
    from rabi_b200.loss import ReverseKLLoss

    name, g, model = case
    x, xt = g["x"].to(dev), g["x_tilde"].to(dev)
    launches0 = model.engine().launches
    with torch.no_grad(), noise.injected([g["z_rkl64"]]):
        kl = ReverseKLLoss(mc_samples=64, reduction=None)(model(xt), model(x))
    assert_kl_close(kl, g["rkl64"], g["rsample_log_prob"].abs().max(), f"{name} (deep path): rKL S=64")
    eng = model.engine()
    with torch.no_grad():
        q = model(x)
        _, zs, _ = model.condition_inputs(xt)
        zs = zs or (None, None)
        gx, klm = eng.rkl_input_grad(xt, zs[0], zs[1], q.A, g["z_rkl_grad"].to(dev), 1.0 / x.shape[0])
    assert rel_err(gx, g["rkl_grad_x"]) < RTOL
    assert_kl_close(klm.mean(), g["rkl_mean"], g["rsample_log_prob"].abs().max(), f"{name} (deep path): mean rKL")
    p = g["pgd"]
    atk = L2PGDAttack(model, loss_fn=ReverseKLLoss(mc_samples=p["S"]), eps=p["eps"], nb_iter=5, eps_iter=p["eps_iter"],
                      clip_min=p["clip_min"], clip_max=p["clip_max"])
    with noise.injected(list(p["z_5"])):
        xa = atk.perturb(x, delta_init=p["delta0"].to(dev))
    prone = pgd_near_kink_rows(build_oracle(g), g["x"], p["delta0"], list(p["z_5"]), p)
    assert_trajectories_match(xa.cpu() - g["x"], p["x_adv_5"] - g["x"], RTOL, prone=prone)
    assert model.engine().launches > launches0
    # the single-chain modes: density of given theta, sampling (+ free log-density)
    with torch.no_grad():
        assert rel_err(model(x).log_prob(g["theta"].to(dev)), g["log_prob"]) < RTOL
        with noise.injected([g["z_rsample"]]):
            qq = model(x)
            th = qq.rsample((4,))
            lp_own = qq.log_prob(th)
    assert rel_err(th, g["rsample"]) < RTOL and rel_err(lp_own, g["rsample_log_prob"]) < RTOL


def test_deep_path_equals_generic_kernel_across_batches(dev):
    """3 hidden layers, D = 9, more pairs than one batch of the deep path holds (RABI_DEEP_T=64 => 148 x 64 pairs per
    batch), ragged sizes: deep == generic per-pair kernel for the metric and for the attack gradient."""
    import os

    ref, model, x = _fresh(dict(dx=12, D=9, hidden=[40, 36, 44], B=1201), dev, seed=7)
    xd = x.to(dev)
    eng = model.engine()
    g = torch.Generator().manual_seed(1)
    S = 9
    z = torch.randn(S, xd.shape[0], 9, generator=g).to(dev)
    with torch.no_grad():
        q = model(xd)
        _, zs, _ = model.condition_inputs(xd)
        outs = []
        for env in ({"RABI_DEEP_T": "64"}, {"RABI_DISABLE_DEEP": "1"}):
            os.environ.update(env)
            try:
                Ap = model(xd + 0.05).A
                kl = eng.rkl(Ap, q.A, z)
                gx, klm = eng.rkl_input_grad(xd + 0.05, zs[0], zs[1], q.A, z, 1.0 / xd.shape[0])
            finally:
                for k in env:
                    os.environ.pop(k, None)
            outs.append((kl, gx, klm))
    (kl_d, gx_d, klm_d), (kl_g, gx_g, klm_g) = outs
    assert float((kl_d - kl_g).abs().max()) < 1e-5 * max(1.0, float(kl_g.abs().max()) * 10)
    assert float((klm_d - klm_g).abs().max()) < 1e-5 * max(1.0, float(klm_g.abs().max()) * 10)
    assert_trajectories_match(gx_d, gx_g, 1e-4, prone=relu_margins(ref, x + 0.05, x, z.cpu()) < TAU)


def test_size_independent_properties_at_config5_scale(dev):
    """BASELINE config 5 sizes (pyloric shapes: D = 31, 3 x 200 hidden; 12 500 rows = one GPU's shard; deep path):
    KL(q, q') == 0 for identical contexts, the density pass reproduces the sampling pass' own log-density, the attack
    stays inside the eps-ball and the clip box and beats a random direction, for the reverse- and the forward-KL loss,
    and attacking the shard in one call equals attacking it in chunks (Q2) on the rows compared."""
    from rabi_b200.attacks import L2PGDAttack
    from rabi_b200.loss import ForwardKLLoss, ReverseKLLoss

    _, model, _ = _fresh(SHAPES["pyloric"], dev, seed=6)
    torch.manual_seed(12)
    x = torch.randn(12500, 100, device=dev)
    with torch.no_grad():
        q = model(x)
        kl_same = ReverseKLLoss(mc_samples=16, reduction=None)(model(x.clone()), q)
        assert float(kl_same.abs().max()) < 2e-3  # two separately projected but identical contexts: rounding only
        th = q.rsample((8,))
        lp_free, lp_dens = q.log_prob(th), q.log_prob(th.clone())
        assert torch.isfinite(lp_free).all()
        assert float((lp_free - lp_dens).abs().max()) < 1e-4 * float(lp_free.abs().max())
    eps = float(0.5 * x.std(1).mean())
    cmin, cmax = float(x.min()), float(x.max())
    for loss_cls in (ReverseKLLoss, ForwardKLLoss):
        atk = L2PGDAttack(model, loss_fn=loss_cls(mc_samples=5), eps=eps, nb_iter=10, eps_iter=eps / 5, clip_min=cmin,
                          clip_max=cmax)
        torch.manual_seed(13)
        xa = atk.perturb(x, chunk_size=500)
        assert float((xa - x).norm(dim=-1).max()) <= eps * (1 + 1e-5)
        assert float(xa.min()) >= cmin and float(xa.max()) <= cmax
        with torch.no_grad():
            kl_adv = loss_cls(mc_samples=64, reduction=None)(model(xa), model(x))
            kl_rand = loss_cls(mc_samples=64, reduction=None)(
                model(x + eps * torch.nn.functional.normalize(torch.randn_like(x), dim=-1)), model(x))
        assert torch.isfinite(kl_adv).all()
        assert float(kl_adv.mean()) > float(kl_rand.mean())


@pytest.mark.parametrize("shape", [dict(dx=100, D=4, hidden=[100, 100], B=301), dict(dx=10, D=10, hidden=[100, 100], B=97),
                                   dict(dx=20, D=31, hidden=[64, 64, 64], B=45), dict(dx=12, D=9, hidden=[40, 36, 44], B=130)])
def test_warp_per_pair_kernel_equals_thread_per_pair_kernels(shape, dev):
    """maf_warp.cuh (launches of <= 1024 pairs: training-time attacks) against the thread-per-pair kernels, reverse- and
    forward-KL gradient, with the weights staged in shared memory and read from global memory, ragged pair counts."""
    import os

    ref, model, x = _fresh(shape, dev, seed=8)
    xd = x.to(dev)
    eng = model.engine()
    g = torch.Generator().manual_seed(2)
    S = 3
    z = torch.randn(S, xd.shape[0], shape["D"], generator=g).to(dev)
    with torch.no_grad():
        q = model(xd)
        _, zs, _ = model.condition_inputs(xd)
        th = q.rsample((2,))
        lp_scale = max(1.0, float(q.log_prob(th.clone()).abs().max()))  # KL values are differences of log-densities
        for fkl in (False, True):
            outs = []
            for env in ({"RABI_WARP_MAX_PAIRS": "0"}, {}, {"RABI_WARP_STAGE": "0"}):
                env = dict(env, RABI_TC="0")   # the FMA kernels among themselves; the tensor-core path has its own test below
                os.environ.update(env)
                try:
                    gx, kl = eng.rkl_input_grad(xd + 0.05, zs[0], zs[1], q.A, z, 1.0 / xd.shape[0], forward_kl=fkl)
                finally:
                    for k in env:
                        os.environ.pop(k, None)
                outs.append((gx, kl))
            prone = (relu_margins(ref, x, x + 0.05, z.cpu()) if fkl else relu_margins(ref, x + 0.05, x, z.cpu())) < TAU
            for gx, kl in outs[1:]:
                assert float((kl - outs[0][1]).abs().max()) < 1e-5 * lp_scale
                assert_trajectories_match(gx, outs[0][0], 1e-4, prone=prone)


@pytest.mark.parametrize("shape", [dict(dx=100, D=4, hidden=[100, 100], B=301), dict(dx=100, D=4, hidden=[100, 100], B=2000),
                                   dict(dx=12, D=3, hidden=[64, 48], B=130), dict(dx=7, D=2, hidden=[20, 36], B=77),
                                   dict(dx=9, D=4, hidden=[104, 97], B=50)])
def test_two_lanes_per_pair_kernel_equals_thread_per_pair_kernels(shape, dev):
    """maf_fast2.cuh (the attack gradient with two lanes per pair, by default only for launches of at most 40 pairs per SM) against
    k_fast (RABI_FAST2=0) and the generic kernel (RABI_DISABLE_FAST=1): lotka_volterra shapes (compile-time leading
    dimension), D < 4, widths that are not multiples of 8, ragged tiles."""
    import os

    ref, model, x = _fresh(shape, dev, seed=9)
    xd = x.to(dev)
    eng = model.engine()
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        q = model(xd)
        _, zs, _ = model.condition_inputs(xd)
        th = q.rsample((2,))
        lp_scale = max(1.0, float(q.log_prob(th.clone()).abs().max()))
        for S in (5, 1):
            z = torch.randn(S, xd.shape[0], shape["D"], generator=g).to(dev)
            outs = []
            # k_fast, k_fast2 (forced), generic, and the tensor-core kernel k_tc (forced for every launch size)
            for env in ({"RABI_FAST2": "0", "RABI_TC": "0"}, {"RABI_FAST2": "2", "RABI_TC": "0"},
                        {"RABI_DISABLE_FAST": "1", "RABI_TC": "0"}, {"RABI_TC": "1", "RABI_TC_GRAD_MIN_PAIRS": "1"}):
                env = dict(env, RABI_WARP_MAX_PAIRS="0")
                os.environ.update(env)
                try:
                    gx, kl = eng.rkl_input_grad(xd + 0.05, zs[0], zs[1], q.A, z, 1.0 / xd.shape[0])
                finally:
                    for k in env:
                        os.environ.pop(k, None)
                outs.append((gx, kl))
            prone = relu_margins(ref, x + 0.05, x, z.cpu()) < TAU
            for gx, kl in outs[1:]:
                assert float((kl - outs[0][1]).abs().max()) < 1e-5 * lp_scale
                assert_trajectories_match(gx, outs[0][0], 1e-4, prone=prone)


@pytest.mark.parametrize("shape", [dict(dx=20, D=31, hidden=[64, 64, 64], B=45), dict(dx=12, D=9, hidden=[40, 36, 44], B=130),
                                   dict(dx=100, D=31, hidden=[200, 200, 200], B=300), dict(dx=16, D=5, hidden=[72, 200, 33], B=1031)])
def test_three_layer_tensor_core_kernel_equals_fma_paths(shape, dev):
    """maf_tcd.cu (three hidden layers up to 200 wide on tcgen05.mma, all per-pair activations in tensor memory) against
    the global-scratch deep path (RABI_TC=0) in every mode: metric, reverse- / forward-KL attack gradient, log_prob,
    rsample + own log-density; pyloric shapes, ragged tiles, unequal widths, several tiles per CTA round."""
    import os

    ref, model, x = _fresh(shape, dev, seed=11)
    xd = x.to(dev)
    eng = model.engine()
    g = torch.Generator().manual_seed(4)
    D = shape["D"]
    with torch.no_grad():
        q = model(xd)
        _, zs, _ = model.condition_inputs(xd)
        for S in (3, 1):
            z = torch.randn(S, xd.shape[0], D, generator=g).to(dev)
            outs = []
            for env in ({"RABI_TC": "0"}, {"RABI_TC": "1", "RABI_TC_GRAD_MIN_PAIRS": "1"}):
                env = dict(env, RABI_WARP_MAX_PAIRS="0")
                os.environ.update(env)
                try:
                    Ap = model(xd + 0.05).A
                    kl = eng.rkl(Ap, q.A, z)
                    gx, klm = eng.rkl_input_grad(xd + 0.05, zs[0], zs[1], q.A, z, 1.0 / xd.shape[0])
                    gxf, klf = eng.rkl_input_grad(xd + 0.05, zs[0], zs[1], q.A, z, 1.0 / xd.shape[0], forward_kl=True)
                    th, lq = eng.rsample(q.A, z)
                    lp = eng.log_prob(Ap, th)
                finally:
                    for k in env:
                        os.environ.pop(k, None)
                outs.append((kl, gx, klm, gxf, klf, th, lq, lp))
            a, b = outs
            lp_scale = max(1.0, float(a[6].abs().max()))
            assert rel_err(b[5], a[5]) < 1e-5, "rsample"
            assert float((b[6] - a[6]).abs().max()) < 1e-5 * lp_scale, "own log-density"
            assert float((b[7] - a[7]).abs().max()) < 1e-5 * lp_scale, "log_prob"
            assert float((b[0] - a[0]).abs().max()) < 1e-5 * lp_scale, "metric"
            assert float((b[2] - a[2]).abs().max()) < 1e-5 * lp_scale and float((b[4] - a[4]).abs().max()) < 1e-5 * lp_scale
            assert_trajectories_match(b[1], a[1], 1e-4, prone=relu_margins(ref, x + 0.05, x, z.cpu()) < TAU)
            assert_trajectories_match(b[3], a[3], 1e-4, prone=relu_margins(ref, x, x + 0.05, z.cpu()) < TAU)


@pytest.mark.parametrize("shape", [dict(dx=100, D=4, hidden=[100, 100], B=301, S=5), dict(dx=10, D=10, hidden=[100, 100], B=530, S=3),
                                   dict(dx=12, D=3, hidden=[64, 48], B=1300, S=1), dict(dx=100, D=4, hidden=[100, 100], B=12007, S=5),
                                   dict(dx=37, D=2, hidden=[20, 36], B=77, S=7)])
def test_persistent_attack_equals_per_iteration_launches(shape, dev):
    """rabi_tc_run (maf_tc.cu: the whole L2-PGD attack as ONE persistent launch -- projection, flow passes, gradient
    back-projection and the perturb_iterative(ord=2) update inside the kernel, a CTA owning its row blocks for all iterations)
    against (a) the per-iteration launch groups of the same library and (b) the oracle, for the reverse- and the forward-KL
    loss: ragged row counts, S = 1, context width != hidden width, D = 10, more than one round of row blocks per CTA."""
    import os

    from oracle import rbi_oracle as O

    ref, model, x = _fresh(shape, dev, seed=21)
    xd = x.to(dev)
    eng = model.engine()
    S, B, D, n_it = shape["S"], shape["B"], shape["D"], 6
    g = torch.Generator().manual_seed(5)
    z_all = torch.randn(n_it, S, B, D, generator=g)
    d0 = 0.01 * torch.randn(B, shape["dx"], generator=g)
    eps = float(0.5 * x.std(1).mean())
    cmin, cmax = float(x.min()), float(x.max())
    with torch.no_grad():
        _, zs, _ = model.condition_inputs(xd)
    launches = {}
    for fkl in (False, True):
        outs = []
        for env in ({"RABI_TC_RUN": "0"}, {"RABI_TC_RUN_MIN_PAIRS": "1"}):
            os.environ.update(env)
            try:
                delta = d0.to(dev).clone()
                n0 = eng.launches
                with torch.no_grad():
                    xa = eng.pgd_run(xd, delta, zs[0], zs[1], None, z_all.to(dev), eps, eps / 4, cmin, cmax, 1.0 / B, forward_kl=fkl)
                launches[tuple(env)] = eng.launches - n0
            finally:
                for k in env:
                    os.environ.pop(k, None)
            outs.append((xa.cpu(), delta.cpu()))
        nb = min(B, 400)   # the oracle walks a bounded number of rows (rows are independent; 1/B is explicit)
        loss = (O.ForwardKLLoss if fkl else O.ReverseKLLoss)(mc_samples=S)
        p = dict(eps=eps, eps_iter=eps / 4, clip_min=cmin, clip_max=cmax, S=S)
        prone = pgd_near_kink_rows(ref, x[:nb], d0[:nb], [z[:, :nb] for z in z_all], p, forward_kl=fkl)
        # (b) vs the oracle on the first rows: advertorch divides the loss by the batch it sees, the engine by B: same direction,
        # the 1e-6 norm floor is far away at these sizes
        with O.injected_base_noise([z[:, :nb] for z in z_all]):
            xa_ref = O.l2pgd_perturb(ref, x[:nb], loss, eps, n_it, eps / 4, cmin, cmax, delta0=d0[:nb])
        for q in ref.parameters():
            q.grad = None
        e = assert_trajectories_match(outs[1][0][:nb] - x[:nb], xa_ref - x[:nb], RTOL, prone=prone)
        # (a) vs the per-iteration path: every row, same tolerance; rows proven near a kink (first rows) or, beyond them, a
        # bounded fraction may sit on the other side of a ReLU kink
        assert_trajectories_match(outs[1][1][:nb], outs[0][1][:nb], RTOL, prone=prone)
        if B > nb:
            assert_trajectories_match(outs[1][1][nb:], outs[0][1][nb:], RTOL, max_flip_frac=0.02)
        assert float((outs[1][0] - x).norm(dim=-1).max()) <= eps * (1 + 1e-5)
        print(f"\n[persistent attack {shape}, {'fKL' if fkl else 'rKL'}] max row rel. err vs oracle {float(e.max()):.2e} "
              f"(near-kink rows {int(prone.sum())}/{nb})")
    # one launch for the whole attack (+ the projection of the clean x, the final clamp and, on first use, the image pack)
    assert launches[("RABI_TC_RUN_MIN_PAIRS",)] <= 4 < launches[("RABI_TC_RUN",)], launches


@pytest.mark.parametrize("shape", ["lotka_volterra", "pyloric"])
def test_batched_eps_sweep_equals_one_attack_per_eps(shape, dev):
    """rows x attack radii in ONE call (rabi_pgd_set_row_scales: per-row eps / eps_iter; L2PGDAttack.perturb_sweep,
    RobustnessMetric.eval_sweep) against one attack per eps -- what the reference runs as one process per value of
    config/experiment/eval_l2_pyloric.yaml:13-16 -- on the same starting points and base noise."""
    from rabi_b200 import noise
    from rabi_b200.attacks import L2PGDAttack
    from rabi_b200.loss import ReverseKLLoss
    from rabi_b200.metric import ReverseKLRobMetric

    sh = dict(SHAPES[shape], B=300)
    ref, model, x = _fresh(sh, dev, seed=31)
    xd = x.to(dev)
    B, D, S, n_it = sh["B"], sh["D"], 5, 6
    std = float(x.std(1).mean())
    eps_rel = [0.1, 0.5, 2.0]
    cmin, cmax = float(x.min()), float(x.max())
    g = torch.Generator().manual_seed(9)
    zs = [torch.randn(n_it, S, B, D, generator=g) for _ in eps_rel]
    d0 = [0.02 * (i + 1) * torch.randn(B, sh["dx"], generator=g) for i in range(len(eps_rel))]
    singles = []
    for i, e in enumerate(eps_rel):
        eps = e * std
        atk = L2PGDAttack(model, loss_fn=ReverseKLLoss(mc_samples=S), eps=eps, nb_iter=n_it, eps_iter=min(eps / n_it * 20, eps),
                          clip_min=cmin, clip_max=cmax)
        with noise.injected(list(zs[i])):
            singles.append(atk.perturb(xd, delta_init=d0[i].to(dev)).cpu())
    atk = L2PGDAttack(model, loss_fn=ReverseKLLoss(mc_samples=S), eps=1.0, nb_iter=n_it, eps_iter=0.1, clip_min=cmin, clip_max=cmax)
    with noise.injected([torch.cat([z[t] for z in zs], dim=1) for t in range(n_it)]):
        sweep = atk.perturb_sweep(xd, [e * std for e in eps_rel], delta_init=torch.cat(d0).to(dev)).cpu()
    assert sweep.shape == (len(eps_rel), B, sh["dx"]) and (atk.eps, atk.eps_iter) == (1.0, 0.1)
    for i, e in enumerate(eps_rel):
        err = assert_trajectories_match(sweep[i] - x, singles[i] - x, RTOL, max_flip_frac=0.02)
        assert float((sweep[i] - x).norm(dim=-1).max()) <= e * std * (1 + 1e-5)
        print(f"\n[eps-sweep {shape}] eps = {e} x std: max row rel. err vs the single-eps attack {float(err.max()):.2e}")
    # the metric over the sweep: one value per eps, increasing with the radius
    met = ReverseKLRobMetric(model, atk, mc_samples=32, attack_attemps=1, device=dev, batch_size=None, reduction="mean")
    torch.manual_seed(3)
    vals = met.eval_sweep(xd, eps_rel)
    v = [float(vals[e]) for e in eps_rel]
    assert all(torch.isfinite(torch.tensor(v))) and v[0] < v[1] < v[2], v


def test_two_row_lanes_equal_one_launch_group(dev):
    """MafEngine.pgd_run splits the rows of an attack on the three-hidden-layer path into two halves on two handles and
    streams when a launch per iteration would end in a ragged round of tiles (rows are independent).  Same x_adv / delta as
    the single-lane run, with a scalar radius, with per-row radii (eps-sweep) and with a column mask; the second lane's
    launches are counted."""
    import os

    from rabi_b200 import noise
    from rabi_b200.attacks import L2PGDAttack
    from rabi_b200.loss import ReverseKLLoss

    sh = dict(SHAPES["pyloric"], B=4003)
    ref, model, x = _fresh(sh, dev, seed=41)
    xd = x.to(dev)
    B, D, S, n_it = sh["B"], sh["D"], 5, 3
    eng = model.engine()
    assert eng._lanes(S, B) == 2 and eng._lanes(S, 64) == 1
    std = float(x.std(1).mean())
    cmin, cmax = float(x.min()), float(x.max())
    g = torch.Generator().manual_seed(12)
    z = [torch.randn(S, B, D, generator=g) for _ in range(n_it)]
    d0 = 0.02 * torch.randn(B, sh["dx"], generator=g)
    mask = torch.ones(sh["dx"], dtype=torch.bool)
    mask[::3] = False
    eps_rows = (0.2 + 0.8 * torch.rand(B, generator=g)) * std
    cases = {"scalar": dict(eps=0.5 * std, eps_iter=0.1 * std, kw={}),
             "per-row radii": dict(eps=eps_rows.to(dev), eps_iter=(eps_rows / 4).to(dev), kw={}),
             "column mask": dict(eps=0.5 * std, eps_iter=0.1 * std, kw={"feature_mask": mask})}
    for name, c in cases.items():
        outs, launches = [], []
        for lanes in ("1", "2"):
            os.environ["RABI_PGD_LANES"] = lanes
            try:
                atk = L2PGDAttack(model, loss_fn=ReverseKLLoss(mc_samples=S), eps=c["eps"], nb_iter=n_it, eps_iter=c["eps_iter"],
                                  clip_min=cmin, clip_max=cmax)
                n0 = eng.launches
                with noise.injected(list(z)):
                    outs.append(atk.perturb(xd, delta_init=d0.to(dev), **c["kw"]).cpu())
                launches.append(eng.launches - n0)
            finally:
                os.environ.pop("RABI_PGD_LANES", None)
        err = assert_trajectories_match(outs[1] - x, outs[0] - x, RTOL, max_flip_frac=0.02)
        assert launches[1] > launches[0], launches      # both lanes launch their own groups
        if name == "column mask":
            assert torch.equal(outs[1][:, ~mask], x[:, ~mask])
        print(f"\n[two row lanes, {name}] max row rel. err vs one lane {float(err.max()):.2e}; launches {launches}")


@pytest.mark.parametrize("kernel", ["k_warp", "k_fast2", "k_fast", "k_tc"])
def test_every_attack_gradient_kernel_against_the_oracle_at_its_own_launch_size(kernel, dev):
    """d[1/B sum_b KL(q(.|x~_b) || q(.|x_b))]/dx~ from each kernel family, at a launch size the dispatcher gives to THAT family,
    against the oracle's autograd gradient on the same base noise (lotka_volterra shapes): k_warp (<= 1024 pairs), k_fast2 (two
    lanes per pair, up to 40 pairs per SM), k_fast (thread per pair) and the tensor-core k_tc."""
    import os

    from oracle import rbi_oracle as O

    B, env = {"k_warp": (200, {}), "k_fast2": (400, {"RABI_TC": "0"}), "k_fast": (2000, {"RABI_TC": "0", "RABI_FAST2": "0"}),
              "k_tc": (2000, {})}[kernel]
    shape = dict(SHAPES["lotka_volterra"], B=B)
    ref, model, x = _fresh(shape, dev, seed=51)
    S, D = 5, shape["D"]
    z = torch.randn(S, B, D, generator=torch.Generator().manual_seed(6))
    xt = x + 0.05 * torch.randn(B, shape["dx"], generator=torch.Generator().manual_seed(7))
    # oracle: autograd through the MC reverse KL
    xr = xt.clone().requires_grad_()
    with O.injected_base_noise([z]):
        loss = O.ReverseKLLoss(mc_samples=S)(ref(xr), ref(x))
    loss.backward()
    for q in ref.parameters():
        q.grad = None
    eng = model.engine()
    os.environ.update(env)
    try:
        with torch.no_grad():
            q = model(x.to(dev))
            _, zs, _ = model.condition_inputs(x.to(dev))
            gx, kl = eng.rkl_input_grad(xt.to(dev), zs[0], zs[1], q.A, z.to(dev), 1.0 / B)
    finally:
        for k in env:
            os.environ.pop(k, None)
    prone = relu_margins(ref, xt, x, z) < TAU
    e = assert_trajectories_match(gx.cpu(), xr.grad, RTOL, prone=prone)
    with torch.no_grad(), O.injected_base_noise([z]):
        kl_ref = O.ReverseKLLoss(mc_samples=S, reduction=None)(ref(xt), ref(x))
        lp_scale = float(ref(x).log_prob(ref(x).sample()).abs().max())
    assert_kl_close(kl, kl_ref, lp_scale, f"{kernel}: per-row rKL of the attack loss")
    print(f"\n[{kernel}, {S * B} pairs] input gradient vs oracle autograd: max row rel. err {float(e.max()):.2e} "
          f"(near-kink rows {int(prone.sum())}/{B})")
