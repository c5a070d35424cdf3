"""Checkpoint / result-file compatibility: a whole pickled reference model written exactly like
``save_model_by_id`` (rbibm/rbibm/utils/utils_data.py:688-705) loads without rbi / pyro installed; weights export back in
the reference's state-dict layout; ``metrics_rob.csv`` rows carry the reference's columns (utils_data.py:71-86)."""
import os

import pytest
import torch

from tests.util import GOLDEN, RTOL, rel_err

CKPT = os.path.join(GOLDEN, "ref_checkpoint.pkl")
# in the fixture the pyro classes live under oracle.thirdparty.pyro_restated (pyro is not installed in the container
# that produced it); the rbi classes carry their real module paths
EXTRA = ("oracle",)


def _expected():
    e = torch.load(os.path.join(GOLDEN, "ref_checkpoint_expected.pt"), weights_only=False)
    e["state_dict"] = {k: (v.float() if v.dtype == torch.bool else v) for k, v in e["state_dict"].items()}
    return e


def test_reference_checkpoint_loads_without_the_reference_installed():
    from rabi_b200.io import load_reference_checkpoint, reference_state_dict
    from rabi_b200.models import ZScoreLayer

    e = _expected()
    m = load_reference_checkpoint(CKPT, device="cpu", extra_packages=EXTRA)
    cfg = e["cfg"]
    assert (m.input_dim, m.output_dim, m.hidden_dims, m.num_transforms, m.shuffle) == \
        (cfg["dx"], cfg["D"], cfg["hidden"], cfg["K"], True)
    assert (m.log_scale_min_clip, m.log_scale_max_clip) == (-7.0, 5.0)
    assert isinstance(m.embedding_net[0], ZScoreLayer) and isinstance(m.embedding_net[1], torch.nn.Identity)
    sd = m.state_dict()
    assert set(sd) == set(e["state_dict"])
    for k, v in e["state_dict"].items():
        assert torch.equal(sd[k].float(), v.float()), k
    y = m.output_transform(torch.zeros(1, cfg["D"]))
    assert torch.allclose(y, (e["box_low"] + e["box_high"])[None] / 2)
    # export: the oracle's reference-layout model accepts the weights strictly
    from oracle import rbi_oracle as O

    emb = torch.nn.Sequential(O.ZScoreLayer(sd["embedding_net.0.mean"], sd["embedding_net.0.std"]), torch.nn.Identity())
    o = O.OracleMAF(cfg["dx"], cfg["D"], hidden_dims=cfg["hidden"], num_transforms=cfg["K"], shuffle=True,
                    embedding_net=emb, log_scale_min_clip=-7, log_scale_max_clip=5, stable=False)
    res = o.load_state_dict(reference_state_dict(m), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    o.output_transform = m.output_transform
    with torch.no_grad():
        assert rel_err(o.eval()(e["x"]).log_prob(e["theta"]), e["log_prob"]) < 1e-6


def test_unknown_model_family_or_embedding_is_refused():
    from rabi_b200.io import _Bag, convert_reference_model

    other = type("GaussianNet", (_Bag,), {"_ref_module": "rbi.models.parametric", "_ref_name": "GaussianNet"})()
    with pytest.raises(NotImplementedError):
        convert_reference_model(other, device=None)


def test_rob_metric_store_writes_the_reference_csv_layout(tmp_path):
    import pandas as pd

    from rabi_b200.io import ROB_EVAL_COLUMNS, RobMetricStore, tensors_to_floats

    store = RobMetricStore(str(tmp_path), "bench")
    x, th, xa = torch.randn(4, 3), torch.randn(4, 2), torch.randn(4, 3)
    value, extra = torch.tensor(0.73), {"q50": torch.tensor(0.5), "q95": torch.tensor(1.9)}
    ids = [store.add_entry(x_unsecure=x, theta_unsecure=th, x_adversarial=xa, id="model-1",
                           metric_rob="ReverseKLRobMetric", attack="L2PGDAttack", attack_loss_fn="ReverseKLLoss",
                           attack_attemps=5, eps=e, eps_abs=2 * e, main_value_rob=tensors_to_floats(value),
                           additional_value_rob=tensors_to_floats(extra), target_strategy=None, eval_time=1.5,
                           params={"a": 1}, not_a_column="ignored") for e in (0.1, 0.5)]
    df = pd.read_csv(os.path.join(str(tmp_path), "bench", "eval", "metrics_rob.csv"))
    assert list(df.columns) == ROB_EVAL_COLUMNS == [
        "id", "metric_rob", "attack", "attack_loss_fn", "attack_attemps", "target_strategy", "eps", "eps_abs",
        "main_value_rob", "additional_value_rob", "id_adversarial", "c2st_x_xtilde", "eval_time", "params"]
    assert len(df) == 2 and list(df["id_adversarial"]) == ids and ids[0] != ids[1]
    assert abs(df["main_value_rob"][0] - 0.73) < 1e-6 and df["eps"].tolist() == [0.1, 0.5]
    adv = torch.load(os.path.join(str(tmp_path), "bench", "eval", ids[1] + ".pkl"))
    assert set(adv) == {"xs", "thetas", "xs_tilde"} and torch.equal(adv["xs_tilde"], xa)
    assert tensors_to_floats([torch.ones(2), "x"]) == [[1.0, 1.0], None]


@pytest.mark.gpu
def test_loaded_reference_checkpoint_reproduces_reference_log_prob():
    from rabi_b200.io import load_reference_checkpoint

    e = _expected()
    m = load_reference_checkpoint(CKPT, device="cuda", extra_packages=EXTRA).eval()
    with torch.no_grad():
        lp = m(e["x"].cuda()).log_prob(e["theta"].cuda())
    assert rel_err(lp, e["log_prob"]) < RTOL
