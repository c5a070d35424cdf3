"""The plugin seam, exercised the way the reference's script does it (CPU, no compute): classes are resolved by
``importlib.import_module(module_path)`` + ``getattr(module, class_name)`` and constructed from the ``params`` of the reference's
own hydra yaml files (rbibm/rbibm/scripts/rbibm_script.py:597-636,753-764).  The yaml VALUES are committed as
tests/golden/ref_configs.json (written by oracle/gen_golden_runs.py from src/rbibm/config/...); where /root/reference exists the
copy is checked against the live files.  Only the module paths are overridden -- the drop-in recipe of INTEGRATION.md."""
import importlib
import json
import os

import pytest
import torch

from tests.util import GOLDEN

# the four overrides
OVERRIDES = {
    "rbi.models": "rabi_b200.models",
    "rbi.defenses": "rabi_b200.defenses",
    "rbi.defenses.trades_regularizers": "rabi_b200.defenses",
    "rbi.attacks.advertorch_attack": "rabi_b200.attacks",
    "rbi.loss": "rabi_b200.loss",
    "rbibm.metric.robustness_metric": "rabi_b200.metric",
}


@pytest.fixture(scope="module")
def cfgs():
    with open(os.path.join(GOLDEN, "ref_configs.json")) as f:
        return json.load(f)


def _resolve(module_path, class_name):
    return getattr(importlib.import_module(OVERRIDES[module_path]), class_name)


def test_committed_config_values_equal_the_reference_yaml(cfgs):
    root = "/root/reference"
    if not os.path.isdir(root):
        pytest.skip("the reference tree exists in the build container only")
    import yaml

    for key, c in cfgs.items():
        with open(os.path.join(root, c["__file__"])) as f:
            live = yaml.safe_load(f)
        assert {k: v for k, v in c.items() if k != "__file__"} == live, key


def _model(cfgs, dx=12, D=3):
    from torch.distributions import biject_to, constraints

    m = cfgs["model"]
    cls = _resolve(m["module_path"], m["class_name"])
    hp = dict(m["params"])
    hp["embedding_net"] = torch.nn.Identity()                    # config/model/embedding_net/identity.yaml
    assert hp["output_transform"] == "biject_to"                 # rbibm/utils/output_transform.py:6-16 on a box prior
    hp["output_transform"] = biject_to(constraints.independent(constraints.interval(-torch.ones(D), torch.ones(D)), 1))
    return cls(dx, D, **hp), hp


def test_model_resolves_and_takes_the_reference_params(cfgs):
    model, hp = _model(cfgs)
    assert type(model).__name__ == "InverseAffineAutoregressiveModel" and type(model).__module__ == "rabi_b200.models"
    keys = set(model.state_dict())
    for k in range(hp["num_transforms"]):
        for j in range(len(hp["hidden_dims"]) + 1):
            assert {f"t{k}.nn.layers.{j}.weight", f"t{k}.nn.layers.{j}.bias", f"t{k}.nn.layers.{j}.mask"} <= keys
        assert f"t{k}.nn.permutation" in keys
    assert model.output_transform is hp["output_transform"]
    # run_train.py:104-128: assignable embedding_net, settable output_transform
    model.embedding_net = torch.nn.Sequential(torch.nn.Identity(), model.embedding_net)
    tf = model.output_transform
    model.output_transform = None
    model.output_transform = tf


@pytest.mark.parametrize("key", ["defense_fisher_trace", "defense_l2pgdTrades_rKL"])
def test_defenses_resolve_and_take_the_reference_params(cfgs, key):
    model, _ = _model(cfgs)
    loss_cls = _resolve(cfgs["train"]["loss_module"], "NLLLoss")          # run_train.py:286-316 builds the NLL loss
    loss = loss_cls(model, **cfgs["train"]["params"]["loss_fn_hyper_parameters"])
    d = cfgs[key]
    cls = _resolve(d["defense_module"], d["defense"])
    hp = dict(d["params"])
    if hp.pop("scale_eps_beta_by_std", False):                            # run_train.py:147-162
        for k in ("eps", "eps_iter"):
            if k in hp:
                hp[k] = hp[k] * 1.7
    defense = cls(model=model, loss_fn=loss, **hp)
    n_post = len(loss.post_loss_regularizers)
    defense.activate()
    assert len(loss.post_loss_regularizers) == n_post + 1
    defense.deactivate()
    assert len(loss.post_loss_regularizers) == n_post
    if key == "defense_fisher_trace":
        assert (defense.beta, defense.ema_mc_samples, defense.algorithm) == (0.05, 5, "ema")
    else:
        assert defense.attack.nb_iter == 20 and abs(defense.attack.eps - 0.5 * 1.7) < 1e-6


def test_attack_and_metric_resolve_like_eval_rob(cfgs):
    model, _ = _model(cfgs)
    a, m, er = cfgs["attack_l2pgd"], cfgs["metric_rKL"], cfgs["eval_rob"]
    attack_cls = _resolve(a["attack_module"], a["attack_class"])
    loss_cls = _resolve(a["loss_module"], a["attack_loss_fn"])
    metric_cls = _resolve(m["metric_module"], m["metric_class"])
    assert (attack_cls.__name__, loss_cls.__name__, metric_cls.__name__) == ("L2PGDAttack", "ForwardKLLoss", "ReverseKLRobMetric")
    # run_rob_eval.py:115-160 on the yaml's params: eps relative to std, eps_iter auto, automatic clip box
    eps_abs = er["eps"] * 2.0
    hp = dict(a["params"])
    assert hp["eps_iter"] == "auto"
    hp["eps_iter"] = min(eps_abs / hp["nb_iter"] * 20, eps_abs)
    hp["clip_min"], hp["clip_max"] = -3.0, 4.0
    hp["targeted"] = er["targeted"]
    attack = attack_cls(model, loss_fn=loss_cls(mc_samples=a["attack_mc_budget"]), eps=eps_abs, **hp)
    assert attack.nb_iter == 200 and attack.predict is model and attack.targeted is False
    metric = metric_cls(model, attack=attack, mc_samples=m["eval_mc_budget"], attack_attemps=er["attack_attemps"],
                        targeted=er["targeted"], target_wrapper=lambda x: x, device="cpu", batch_size=er["batch_size"],
                        reduction="mean_summary", **dict(m["params"]))
    assert metric.loss_fn.mc_samples == 256 and metric.batch_size == 1000
    # the reverse-KL attack loss of eval_rob/metric_attack/l2pgd_rKL (BASELINE config 2) resolves from the same module
    assert _resolve(a["loss_module"], "ReverseKLLoss").__name__ == "ReverseKLLoss"
