"""The reference's RUNNERS over the product (run on the B200 with ``-m gpu``).

oracle/rbibm_runs_restated.py restates ``rbibm.runs.run_rob_eval.run_rob_eval`` and ``rbibm.runs.run_train.run_train``; the
restatements are pinned bit-for-bit against the unmodified runners in the build container (oracle/gen_golden_runs.py).  Here
the same functions drive ``rabi_b200`` classes, resolved by module path + class name like the script does, on the fixture's
data, initial weights, PGD starting points and base noise; results are compared with what the unmodified reference produced
(tests/golden/runs_small.pt)."""
import importlib
import json
import os

import pytest
import torch

from tests.util import GOLDEN, RTOL, row_rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def g():
    return torch.load(os.path.join(GOLDEN, "runs_small.pt"), weights_only=False)


@pytest.fixture(scope="module")
def cfgs():
    with open(os.path.join(GOLDEN, "ref_configs.json")) as f:
        return json.load(f)


def _cls(module, name):
    return getattr(importlib.import_module(module), name)


def _model(g, state, with_zscore):
    from torch.distributions import biject_to, constraints

    s = g["shape"]
    cls = _cls("rabi_b200.models", "InverseAffineAutoregressiveModel")
    tf = biject_to(constraints.independent(constraints.interval(g["low"], g["high"]), 1))
    m = cls(s["dx"], s["D"], hidden_dims=s["hidden"], num_transforms=s["K"], shuffle=False, output_transform=tf,
            **g["model_kw"])
    if with_zscore:
        Z = _cls("rabi_b200.models", "ZScoreLayer")
        m.embedding_net = torch.nn.Sequential(Z(torch.zeros(s["dx"]), torch.ones(s["dx"])), m.embedding_net)
    m.load_state_dict({k: (v.float() if v.dtype == torch.bool else v) for k, v in state.items()}, strict=True)
    return m


def _loaders(g):
    from torch.utils.data import DataLoader, TensorDataset

    mk = lambda k: DataLoader(TensorDataset(g[f"{k}_x"], g[f"{k}_theta"]), batch_size=g["shape"]["batch"], shuffle=False)
    return mk("train"), mk("val"), mk("test")


@pytest.mark.parametrize("tag", ["nll", "fim"])
def test_run_train_over_the_product(g, cfgs, dev, tag):
    from oracle import rbibm_runs_restated as R
    from rabi_b200 import noise

    ref = g[f"train_{tag}"]
    model = _model(g, g[f"train_{tag}_init"], with_zscore=False)
    hp = {k: v for k, v in g["train_hp"].items() if k not in ("initialize_as_prior", "initialize_as_prior_rounds")}
    defense_cls = None
    if tag == "fim":
        d = cfgs["defense_fisher_trace"]
        defense_cls = _cls(d["defense_module"].replace("rbi.defenses", "rabi_b200.defenses"), d["defense"])
        assert ref["defense_hp"] == d["params"]
    tr, va, te = _loaders(g)
    with noise.injected([z.to(dev) for z in ref["z"]]) as left:
        model, defense, tl, vl, tel = R.run_train(model, _cls("rabi_b200.loss", "NLLLoss"), _cls("rabi_b200.models", "ZScoreLayer"),
                                                  tr, va, te, defense=defense_cls, optimizer=torch.optim.Adam,
                                                  defense_hyper_parameters=dict(ref["defense_hp"]), device=dev, **hp)
        assert not left, "the product drew fewer MC samples than the reference"
    tl, vl = torch.stack(tl), torch.cat(vl)
    print(f"\n[run_train {tag}] train {tl.tolist()} vs reference {ref['train_loss'].tolist()}; val {vl.tolist()} vs {ref['val_loss'].tolist()}; "
          f"test {float(tel):.6f} vs {float(ref['test_loss']):.6f}")
    assert len(tl) == ref["epochs"]
    assert float((tl - ref["train_loss"]).abs().max()) <= RTOL * float(ref["train_loss"].abs().max())
    assert float((vl - ref["val_loss"]).abs().max()) <= RTOL * float(ref["val_loss"].abs().max())
    assert abs(float(tel) - float(ref["test_loss"])) <= RTOL * abs(float(ref["test_loss"]))
    # final weights after 6 Adam steps (z-score buffers included): the surface run_train leaves behind
    sd = model.state_dict()
    assert set(sd) == set(ref["final"])
    worst = 0.0
    for k, v in ref["final"].items():
        a, b = sd[k].detach().float().cpu(), v.float()
        worst = max(worst, float((a - b).abs().max()) / max(float(b.abs().max()), 1e-3))
    print(f"[run_train {tag}] worst relative weight deviation after training: {worst:.2e}")
    assert worst < 20 * RTOL     # Adam's sign-like first steps amplify fp32 gradient rounding; losses above are at 1e-4
    assert model.output_transform is not None and type(model.embedding_net[0]).__name__ == "ZScoreLayer"
    if tag == "fim":
        assert type(defense).__name__ == "FIMTraceRegularizer" and defense.ema.t == 6


def test_run_rob_eval_over_the_product(g, cfgs, dev, monkeypatch):
    from oracle import rbibm_runs_restated as R
    from rabi_b200 import noise
    import rabi_b200.attacks as A

    r = g["rob"]
    a, m = cfgs["attack_l2pgd"], cfgs["metric_rKL"]
    model = _model(g, g["rob_model"], with_zscore=True)
    attack_cls = _cls(a["attack_module"].replace("rbi.attacks.advertorch_attack", "rabi_b200.attacks"), a["attack_class"])
    metric_cls = _cls(m["metric_module"].replace("rbibm.metric.robustness_metric", "rabi_b200.metric"), m["metric_class"])
    loss_cls = _cls(a["loss_module"].replace("rbi.loss", "rabi_b200.loss"), "ReverseKLLoss")
    q = [r["delta0"].to(dev)]
    monkeypatch.setattr(A, "rand_init_delta_l2", lambda x, eps, lo, hi: q.pop(0).clone())
    with noise.injected(list(r["z_attack"]) + [r["z_eval"]]) as left:
        out = R.run_rob_eval(model, metric_cls, attack_cls, g["rob_x"], g["rob_theta"], attack_loss_fn=loss_cls, device=dev,
                             **r["kw"])
        assert not left
    assert out["metric_name"] == r["name"] == "ReverseKLRobMetric"
    assert out["eps_abs"] == r["eps_abs"]
    for k, v in r["attack_params"].items():      # eps_iter auto, automatic clip box, targeted flag: the runner's conventions
        assert out["attack_params"][k] == v, k
    val, ref = float(out["value"]), float(r["value"])
    print(f"\n[run_rob_eval] value {val:.6e} vs reference {ref:.6e}; eps_abs {out['eps_abs']:.4f}")
    assert abs(val - ref) <= 10 * RTOL * abs(ref)
    assert set(out["additional_values"]) == set(r["additional"])
    for k, v in r["additional"].items():
        assert abs(float(out["additional_values"][k]) - float(v)) <= 20 * RTOL * max(abs(float(v)), abs(ref)), k
    # the most affected rows and their adversarial examples
    assert out["x_idx"].tolist() == r["x_idx"].tolist()
    e = row_rel_err(out["x_adversarial"] - out["x_unsecure"], r["x_adversarial"] - r["x_unsecure"])
    print(f"[run_rob_eval] adversarial examples: max row rel. err {float(e.max()):.2e}")
    assert float(e.max()) < 10 * RTOL
    assert torch.equal(out["x_unsecure"], r["x_unsecure"]) and torch.equal(out["theta_unsecure"], r["theta_unsecure"])
