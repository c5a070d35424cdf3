"""Checkpoint and result-file compatibility with the reference harness.

* ``convert_reference_model`` / ``load_reference_checkpoint``: the reference stores whole pickled modules
  (``torch.save(model.to("cpu"), models/<id>.pkl)``, rbibm/rbibm/utils/utils_data.py:609-621,688-705).  Such a file names
  classes of ``rbi`` and ``pyro``; here it is read WITHOUT either package: a re-mapping unpickler turns every object of
  those packages into an inert attribute bag, and the MAF is rebuilt from the bags (structure from the tensor shapes,
  clips from the transform kwargs, z-score layer, output transform, permutations, masks, weights).
* ``reference_state_dict``: weights back in the layout ``rbi.models.InverseAffineAutoregressiveModel.load_state_dict``
  accepts (the key layout is shared, rabi_b200/models.py).
* ``RobMetricStore``: appends rows to ``<root>/<name>/eval/metrics_rob.csv`` and stores the adversarial examples next to
  it with the reference's column set and file names (utils_data.py:71-86,430-470,540-565), so the unchanged notebooks
  can read results produced by this backend.
"""
from __future__ import annotations

import os
import pickle
import uuid
from collections import OrderedDict
from typing import Any, Dict, Iterable, Optional

import torch
from torch import nn

from .models import InverseAffineAutoregressiveModel, ZScoreLayer

_REFERENCE_PACKAGES = ("rbi", "pyro")


class _Bag:
    """Attribute bag standing in for an object of a package that is not installed (state only, no behaviour)."""

    _ref_module = ""
    _ref_name = ""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):  # (dict, slots) form
            state = {**(state[0] or {}), **state[1]}
        self.__dict__.update(state)

    def __call__(self, *a, **k):  # pickled functools.partial targets etc.
        return None

    def __repr__(self):
        return f"<reference {self._ref_module}.{self._ref_name}>"


def _make_pickle_module(packages: Iterable[str]):
    packages = tuple(packages)
    cache: Dict[tuple, type] = {}

    class Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            if module.split(".")[0] in packages:
                key = (module, name)
                if key not in cache:
                    cache[key] = type(name, (_Bag,), {"_ref_module": module, "_ref_name": name})
                return cache[key]
            return super().find_class(module, name)

    class _Mod:
        pass

    mod = _Mod()
    mod.Unpickler = Unpickler
    mod.load = lambda f, **kw: Unpickler(f, **kw).load()
    mod.__name__ = "rabi_b200_remap_pickle"
    return mod


def _is(obj, name: str) -> bool:
    return isinstance(obj, _Bag) and obj._ref_name == name or type(obj).__name__ == name


def _children(m) -> "OrderedDict[str, Any]":
    return m.__dict__.get("_modules", OrderedDict()) if isinstance(m, _Bag) else OrderedDict(m.named_children())


def _tensors(m) -> "OrderedDict[str, torch.Tensor]":
    """Own parameters + persistent buffers of a module or of its attribute bag."""
    out = OrderedDict()
    if isinstance(m, _Bag):
        skip = m.__dict__.get("_non_persistent_buffers_set", set())
        for k, v in m.__dict__.get("_parameters", {}).items():
            if v is not None:
                out[k] = v.data
        for k, v in m.__dict__.get("_buffers", {}).items():
            if v is not None and k not in skip:
                out[k] = v
    else:
        for k, v in m.named_parameters(recurse=False):
            out[k] = v.data
        for k, v in m.named_buffers(recurse=False):
            if k not in m._non_persistent_buffers_set:
                out[k] = v
    return out


def _flat_state(m, prefix="") -> "OrderedDict[str, torch.Tensor]":
    out = OrderedDict((prefix + k, v) for k, v in _tensors(m).items())
    for name, child in _children(m).items():
        if child is not None:
            out.update(_flat_state(child, prefix + name + "."))
    return out


def _convert_embedding(net):
    """Embedding nets made of torch modules load as they are; the reference's ZScoreLayer is re-created."""
    if net is None:
        return nn.Identity()
    if _is(net, "ZScoreLayer"):
        t = _tensors(net)
        return ZScoreLayer(t["mean"].clone(), t["std"].clone())
    if isinstance(net, _Bag):
        raise NotImplementedError(f"embedding net {net!r} is defined outside torch.nn; convert it by hand and pass it "
                                  "as `embedding_net=`")
    if isinstance(net, nn.Sequential):
        return nn.Sequential(*[_convert_embedding(c) for c in net])
    return net


def _get(m, name, default=None):
    return m.__dict__.get(name, default) if isinstance(m, _Bag) else getattr(m, name, default)


def convert_reference_model(ref, device="cuda", embedding_net: Optional[nn.Module] = None):
    """A loaded ``rbi.models.InverseAffineAutoregressiveModel`` (the live object, or its attribute bag from
    ``load_reference_checkpoint``) -> ``rabi_b200.models.InverseAffineAutoregressiveModel`` with identical weights,
    masks, permutations, z-score layer and output transform."""
    if not (_is(ref, "InverseAffineAutoregressiveModel")):
        raise NotImplementedError(f"only the maf_pyro family (InverseAffineAutoregressiveModel) is supported, got {ref!r}")
    kids = _children(ref)
    transforms = [k for k in kids if k.startswith("t") and k[1:].isdigit()]
    K = len(transforms)
    state = _flat_state(ref)
    state = OrderedDict((k, v) for k, v in state.items() if not k.startswith("embedding_net."))
    L = sum(1 for k in state if k.startswith("t0.nn.layers.") and k.endswith(".weight"))
    hidden = [int(state[f"t0.nn.layers.{j}.weight"].shape[0]) for j in range(L - 1)]
    D = int(state["base_dist_mean"].shape[0])
    kwargs = dict(_get(kids["t0"], "kwargs", None) or _get(ref, "_kwargs", None) or {})
    emb = embedding_net if embedding_net is not None else _convert_embedding(kids.get("embedding_net"))
    model = InverseAffineAutoregressiveModel(int(_get(ref, "input_dim")), D, hidden_dims=hidden, num_transforms=K,
                                             shuffle="permute0" in state, with_cache=bool(_get(ref, "with_cache", False)),
                                             output_transform=_get(ref, "_output_transform", None),
                                             embedding_net=emb, **kwargs)
    own = model.state_dict()
    emb_state = {k: v for k, v in own.items() if k.startswith("embedding_net.")}
    missing = model.load_state_dict({**emb_state, **{k: v.clone() for k, v in state.items()}}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return model.to(device) if device is not None else model


def load_reference_checkpoint(path: str, device="cuda", embedding_net: Optional[nn.Module] = None,
                              extra_packages: Iterable[str] = ()):
    """``models/<id>.pkl`` written by the reference -> a rabi_b200 model.  Needs neither ``rbi`` nor ``pyro``."""
    mod = _make_pickle_module(_REFERENCE_PACKAGES + tuple(extra_packages))
    ref = torch.load(path, map_location="cpu", pickle_module=mod, weights_only=False)
    return convert_reference_model(ref, device=device, embedding_net=embedding_net)


def reference_state_dict(model: InverseAffineAutoregressiveModel) -> "OrderedDict[str, torch.Tensor]":
    """CPU state dict for ``rbi.models.InverseAffineAutoregressiveModel.load_state_dict`` (same key layout)."""
    return OrderedDict((k, v.detach().cpu().clone()) for k, v in model.state_dict().items())


# ------------------------------------------------------------------------------------------------- result files
ROB_EVAL_COLUMNS = ["id", "metric_rob", "attack", "attack_loss_fn", "attack_attemps", "target_strategy", "eps", "eps_abs",
                    "main_value_rob", "additional_value_rob", "id_adversarial", "c2st_x_xtilde", "eval_time", "params"]


def tensors_to_floats(value):
    """rbibm/rbibm/utils/torchutils.py:34-50 (anything that is not a tensor / container of tensors becomes None)."""
    if isinstance(value, dict):
        return {k: tensors_to_floats(v) for k, v in value.items()}
    if isinstance(value, (list, tuple)):
        return [tensors_to_floats(v) for v in value]
    if isinstance(value, torch.Tensor):
        return float(value.cpu()) if value.numel() <= 1 else [tensors_to_floats(v) for v in list(value)]
    return None


class RobMetricStore:
    """``<root>/<name>/eval/metrics_rob.csv`` + ``<id_adversarial>.pkl`` in the reference's format."""

    def __init__(self, root: str, name: str):
        self.dir = os.path.join(root, name, "eval")
        self.csv = os.path.join(self.dir, "metrics_rob.csv")
        os.makedirs(self.dir, exist_ok=True)
        if not os.path.exists(self.csv):
            with open(self.csv, "w") as f:
                f.write(",".join(ROB_EVAL_COLUMNS) + "\n")

    def add_entry(self, x_unsecure=None, theta_unsecure=None, x_adversarial=None, **fields) -> str:
        import pandas as pd

        known = set(pd.read_csv(self.csv, usecols=["id_adversarial"])["id_adversarial"].astype(str))
        id_adv = str(uuid.uuid4())
        while id_adv in known:
            id_adv = str(uuid.uuid4())
        cpu = lambda t: t.detach().cpu() if isinstance(t, torch.Tensor) else t
        torch.save({"xs": cpu(x_unsecure), "thetas": cpu(theta_unsecure), "xs_tilde": cpu(x_adversarial)},
                   os.path.join(self.dir, id_adv + ".pkl"))
        row = {c: [pd.NA] for c in ROB_EVAL_COLUMNS}
        for k, v in fields.items():
            if k in row:
                row[k] = [v]
        row["id_adversarial"] = [id_adv]
        with open(self.csv, "a") as f:
            pd.DataFrame(row).to_csv(f, mode="a", header=f.tell() == 0, index=False)
        return id_adv

    def add_metric_result(self, model_id: str, metric, out, eps: float, eps_abs: float, eval_time: float,
                          target_strategy=None, params=None, c2st_x_xtilde=float("nan"), x=None, theta=None,
                          num_examples: int = 100) -> str:
        """Convenience wrapper with the field conventions of rbibm_script.py:803-823: ``out`` is what
        ``metric.eval`` returned (value or (value, supplement))."""
        value, extra = out if isinstance(out, tuple) else (out, None)
        x_unsec = th_unsec = x_adv = None
        if x is not None:
            idx, x_adv = metric.generate_adversarial_examples(x.to(metric.device), num_examples=num_examples)
            x_unsec = x[idx.cpu()]
            th_unsec = theta[idx.cpu()] if theta is not None else None
        loss_fn = getattr(metric.attack, "loss_fn", None)
        return self.add_entry(
            x_unsecure=x_unsec, theta_unsecure=th_unsec, x_adversarial=x_adv, id=model_id,
            metric_rob=type(metric).__name__, attack=type(metric.attack).__name__,
            attack_loss_fn=type(loss_fn).__name__ if loss_fn is not None else None,
            attack_attemps=metric.attack_attemps, eps=float(eps), eps_abs=float(eps_abs),
            main_value_rob=tensors_to_floats(value), additional_value_rob=tensors_to_floats(extra),
            target_strategy=target_strategy, params=params, c2st_x_xtilde=c2st_x_xtilde, eval_time=eval_time)
