"""Shared helpers for the test-suite (checker side: may import ``oracle``)."""
from __future__ import annotations

import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = ["tiny", "gaussian_linear", "lotka_volterra", "pyloric_small", "tiny_shuffle"]

# north_star: "log_prob, KL, Fisher trace and PGD perturbations agree within 1e-4 relative error in fp32".
# Relative error is measured against the magnitude of the reference tensor (max-abs), which is the meaningful scale
# for quantities that are differences of O(1..10) log-densities (KL values, gradients).
RTOL = 1e-4


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    finite = torch.isfinite(b)
    assert torch.equal(torch.isfinite(a), finite), "finite pattern differs"
    if not finite.any():
        return 0.0
    scale = b[finite].abs().max().clamp_min(1e-30)
    return float((a[finite] - b[finite]).abs().max() / scale)


def assert_kl_close(kl, ref, lp_scale, what, rtol=RTOL):
    """KL values against the reference, relative to the KL ITSELF (north_star: 'KL ... within 1e-4 relative'): max abs
    error <= rtol * max|KL_ref| + 32 fp32 ulps of the log-density scale (a KL estimate is a mean of differences of two
    fp32 log-densities of magnitude ``lp_scale``; the reference's own value carries that rounding).  The achieved error is
    printed into the test log."""
    kl, ref = kl.detach().double().cpu().reshape(-1), ref.detach().double().cpu().reshape(-1)
    err = float((kl - ref).abs().max())
    scale = float(ref.abs().max())
    floor = 32 * 1.1920929e-07 * float(lp_scale)
    print(f"\n[{what}] max |KL - KL_ref| = {err:.2e} on max |KL_ref| = {scale:.3e}: {err / max(scale, 1e-30):.2e} relative "
          f"(tolerance {rtol:g} relative + {floor:.1e} rounding floor)")
    assert err <= rtol * scale + floor, (what, err, scale)


def row_rel_err(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """per-row ||a - b|| / ||b||"""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).norm(dim=-1) / b.norm(dim=-1).clamp_min(1e-30)


TAU = 4e-6   # near-kink threshold on the normalised ReLU margin (relu_margins below): about 64 fp32 ulps


def assert_trajectories_match(delta, delta_ref, tol, max_flip_frac=0.05, prone=None):
    """Multi-step PGD perturbations (or input gradients), compared per observation row.

    A ReLU network is only piecewise smooth: when a pre-activation lands within rounding distance of 0, ANY change of
    fp32 summation order can flip that unit's mask; the input gradient then jumps and that row departs by O(1e-3..1e-2)
    while every other row agrees to O(1e-6).  A row may exceed ``tol`` only
      * with its PROOF -- ``prone`` [B] bool marks the rows for which an fp64 run of the oracle found a pre-activation
        within TAU of its kink in the gradient evaluation(s) compared (relu_margins / pgd_near_kink_rows); or,
      * where no oracle is at hand (product-vs-product comparisons), as a fraction of the rows with NO floor:
        int(max_flip_frac * rows), i.e. zero on fixtures of fewer than 20 rows.
    All other rows must meet the tolerance and the typical error must be tiny."""
    e = row_rel_err(delta, delta_ref)
    bad = e > tol
    if prone is not None:
        unproven = bad & ~prone.cpu()
        assert not bool(unproven.any()), (f"rows {unproven.nonzero().flatten().tolist()} differ by {e[unproven].tolist()} "
                                          f"(> {tol}) without a near-kink event")
    else:
        assert int(bad.sum()) <= int(max_flip_frac * e.numel()), f"{int(bad.sum())}/{e.numel()} rows differ by more than {tol}: {e.max():.3e}"
    assert float(e.median()) < tol / 10, float(e.median())
    return e


def pgd_near_kink_rows(oracle, x, delta0, zs, p, forward_kl=False, tau=TAU):
    """Rows of an L2-PGD run (oracle arithmetic, fp32 states; margins in fp64) that pass within ``tau`` of a ReLU kink in at
    least one of the ``len(zs)`` gradient evaluations: [B] bool.  ``p``: dict(eps, eps_iter, clip_min, clip_max, S)."""
    from oracle import rbi_oracle as O

    states, calls = [], {"n": 0}

    def hook(_m, args):
        if calls["n"] >= 1:   # call 0 is the target y = model(x)
            states.append(args[0].detach().clone())
        calls["n"] += 1

    h = oracle.register_forward_pre_hook(hook)
    loss = (O.ForwardKLLoss if forward_kl else O.ReverseKLLoss)(mc_samples=p["S"])
    try:
        with O.injected_base_noise([z.cpu() for z in zs]):
            O.l2pgd_perturb(oracle, x, loss, p["eps"], len(zs), p["eps_iter"], p["clip_min"], p["clip_max"], delta0=delta0)
    finally:
        h.remove()
        for q in oracle.parameters():
            q.grad = None
    prone = torch.zeros(x.shape[0], dtype=torch.bool)
    for xt, z in zip(states, zs):
        m = relu_margins(oracle, x, xt, z.cpu()) if forward_kl else relu_margins(oracle, xt, x, z.cpu())
        prone |= m < tau
    return prone


def load_golden(name):
    g = torch.load(os.path.join(GOLDEN, f"{name}.pt"), weights_only=False)
    g["state_dict"] = {k: (v.float() if v.dtype == torch.bool else v) for k, v in g["state_dict"].items()}
    return g


def box_transform(g):
    from torch.distributions import biject_to, constraints

    if not g["cfg"]["box"]:
        return None
    return biject_to(constraints.independent(constraints.interval(g["box_low"], g["box_high"]), 1))


def box_transform_on(g, device):
    from torch.distributions import biject_to, constraints

    if not g["cfg"]["box"]:
        return None
    return biject_to(constraints.independent(constraints.interval(g["box_low"].to(device), g["box_high"].to(device)), 1))


def build_oracle(g, with_output_transform=True):
    from oracle import rbi_oracle as O

    cfg = g["cfg"]
    emb = None
    if cfg["zscore"]:
        emb = torch.nn.Sequential(O.ZScoreLayer(g["zs_mean"], g["zs_std"]), torch.nn.Identity())
    m = O.OracleMAF(cfg["dx"], cfg["D"], hidden_dims=cfg["hidden"], num_transforms=cfg["K"], shuffle=cfg["shuffle"],
                    output_transform=box_transform(g) if with_output_transform else None, embedding_net=emb,
                    log_scale_min_clip=cfg["log_scale_min_clip"], log_scale_max_clip=cfg["log_scale_max_clip"],
                    stable=cfg["stable"])
    m.load_state_dict(g["state_dict"], strict=True)
    return m.eval()


def build_product(g, device, with_output_transform=True):
    from rabi_b200.models import InverseAffineAutoregressiveModel, ZScoreLayer

    cfg = g["cfg"]
    m = InverseAffineAutoregressiveModel(cfg["dx"], cfg["D"], hidden_dims=cfg["hidden"], num_transforms=cfg["K"],
                                         shuffle=cfg["shuffle"], log_scale_min_clip=cfg["log_scale_min_clip"],
                                         log_scale_max_clip=cfg["log_scale_max_clip"], stable=cfg["stable"])
    if cfg["zscore"]:
        # exactly what run_train does (rbibm/runs/run_train.py:104-111)
        m.embedding_net = torch.nn.Sequential(ZScoreLayer(g["zs_mean"], g["zs_std"]), m.embedding_net)
    missing = m.load_state_dict(g["state_dict"], strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    m = m.to(device).eval()
    if with_output_transform:
        m.output_transform = box_transform_on(g, device)
    return m


def relu_margins(oracle_model, x_sample, x_density, z):
    """ReLU margins of ONE reverse-KL attack-gradient evaluation, in fp64 on the CPU oracle (checker side).

    For every hidden unit the conditioners evaluate while q(.|x_sample) is sampled with base noise ``z`` [S, B, D] and
    the samples are scored under q(.|x_density) (reverse KL attack: x_sample = x + delta, x_density = x; forward KL: swapped), margin = |pre-activation| / (sum_i |w_i in_i| + |b|): the distance of
    the unit from its kink in units of the magnitude its fp32 rounding error scales with.  Returns the minimum over
    units, layers, transforms, both posteriors and the S samples: [B].  A row whose margin is below a few fp32 ulps is
    one where two correct fp32 implementations with different summation orders may take different sides of a kink
    (the input gradient then jumps by O(1 %), the normalised PGD step of that row by O(1e-3))."""
    import copy

    m = copy.deepcopy(oracle_model).double().eval()
    x_pert, x_clean, z = x_sample.double(), x_density.double(), z.double()
    best = torch.full(z.shape[:2], float("inf"), dtype=torch.float64)

    def arn_margin(arn, ctx, xin):
        nonlocal best
        inp = torch.cat([ctx.expand(xin.shape[:-1] + (ctx.shape[-1],)), xin], dim=-1)
        for layer in arn.layers[:-1]:
            W = layer.weight * layer.mask
            pre = inp @ W.T + layer.bias
            scale = inp.abs() @ W.abs().T + layer.bias.abs()
            best = torch.minimum(best, (pre.abs() / scale.clamp_min(1e-300)).amin(dim=-1))
            inp = torch.relu(pre)

    with torch.no_grad():
        h_p, h_q = m.embedding_net(x_pert), m.embedding_net(x_clean)
        v = z
        for t in m.flow_transforms:      # sampling chain under x_pert: the conditioner's final input is the transform's output
            if not hasattr(t, "condition"):      # Permute layer (shuffle=True): no ReLU units
                v = t(v)
                continue
            v = t.condition(h_p)(v)
            arn_margin(t.nn, h_p, v)
        w = v
        for t in reversed(m.flow_transforms):   # density chain under x_clean
            if not hasattr(t, "condition"):
                w = t.inv(w)
                continue
            arn_margin(t.nn, h_q, w)
            w = t.condition(h_q).inv(w)
    return best.amin(dim=0)
