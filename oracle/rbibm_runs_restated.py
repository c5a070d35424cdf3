"""ORACLE (test infrastructure, not product code): the two reference RUNNERS that sit on either side of the hot path,
restated so that they can be executed on the GPU box (where /root/reference does not exist) over ``rabi_b200`` classes.

  run_rob_eval  <- rbibm/rbibm/runs/run_rob_eval.py:23-241   (untargeted branch; the C2ST side output is dropped)
  run_train     <- rbibm/rbibm/runs/run_train.py:22-283, EarlyStopping :352-404 (NLL branch of construct_loss :286-316)

Pinned by oracle/gen_golden_runs.py: the UNMODIFIED reference runners are executed in the build container over the
unmodified ``rbi`` classes, and these restatements, executed over the oracle classes on the same data and base noise, must
return bit-identical results before tests/golden/runs_*.pt is written.  Classes are passed in, exactly like the
reference's script resolves them (importlib + getattr, rbibm_script.py:597-636,753-764), so the same functions drive the
oracle classes (pinning) and the product classes (tests/test_gpu_runs.py).
"""
from __future__ import annotations

from copy import deepcopy
from typing import Any, Callable, Optional

import torch
from torch import nn


# ------------------------------------------------------------------------------------------------- robustness eval
def run_rob_eval(model, metric_class, attack_class, xs, thetas, *, attack_loss_fn: Optional[Callable] = None,
                 attack_attemps: int = 10, attack_mc_budget: int = 4, eval_mc_budget: int = 100, eps: float = 0.5,
                 attack_hyper_parameters: Optional[dict] = None, metric_hyper_parameters: Optional[dict] = None,
                 num_adversarial_examples: int = 1000, reduction: Optional[str] = "mean_summary", device="cpu",
                 eps_relative_to_std: bool = True, clip_min_max_automatic: bool = True,
                 batch_size: Optional[int] = None):
    """Untargeted robustness evaluation on given (xs, thetas) -- the ``test_loader`` branch of the reference.

    Conventions that the product must honour (run_rob_eval.py:82-197):
      * eps is RELATIVE to the mean per-row std of xs (``xs.std(1).mean()``) when ``eps_relative_to_std``;
      * ``eps_iter == "auto"`` -> min(eps_abs / nb_iter * 20, eps_abs);
      * ``clip_min_max_automatic`` -> clip box = [xs.min(), xs.max()] (overrides the yaml's +-1000);
      * ``targeted`` is always passed to the attack constructor; the attack loss gets ``mc_samples=attack_mc_budget``;
      * the metric is built with ``mc_samples=eval_mc_budget`` and the metric yaml's ``params`` (mask: null -> None);
      * returns the adversarial examples of the ``num_adversarial_examples`` most affected rows and their (x, theta).
    """
    metric_params = dict(metric_hyper_parameters or {})
    mask = None
    if metric_params.get("mask") is not None:
        mask = torch.as_tensor(metric_params["mask"]).bool()
        metric_params["mask"] = mask
    if eps_relative_to_std:
        std = xs.std(1).mean() if mask is None else xs[..., mask].std(1).mean()
        eps_abs = float(eps * std)
    else:
        eps_abs = float(eps)
    model = model.to(device).eval()
    loss_fn_a = attack_loss_fn(mc_samples=attack_mc_budget) if attack_loss_fn is not None else None
    hp = dict(attack_hyper_parameters or {})
    hp["targeted"] = False
    if hp.get("eps_iter") == "auto":
        hp["eps_iter"] = min((eps_abs / hp.get("nb_iter", 1)) * 20, eps_abs)
    if clip_min_max_automatic:
        hp["clip_min"], hp["clip_max"] = float(xs.min()), float(xs.max())
    attack = attack_class(model, loss_fn=loss_fn_a, eps=eps_abs, **hp) if loss_fn_a is not None else \
        attack_class(model, eps=eps_abs, **hp)
    metric = metric_class(model, attack=attack, mc_samples=eval_mc_budget, attack_attemps=attack_attemps, targeted=False,
                          target_wrapper=lambda x: x, device=device, batch_size=batch_size, reduction=reduction,
                          **metric_params)
    out = metric.eval(xs)
    value, additional = out if isinstance(out, tuple) else (out, None)
    x_idx, x_adversarial = metric.generate_adversarial_examples(xs, num_examples=num_adversarial_examples)
    x_idx, x_adversarial, xs = x_idx.cpu(), x_adversarial.cpu(), xs.cpu()
    return dict(metric_name=metric_class.__name__, value=value, additional_values=additional, x_unsecure=xs[x_idx],
                theta_unsecure=thetas[x_idx], x_adversarial=x_adversarial, eps_abs=eps_abs, attack_params=hp,
                x_idx=x_idx)


# ------------------------------------------------------------------------------------------------- training
class EarlyStopping:
    """run_train.py:352-404: keeps the best state_dict; restores it on NaN (up to ``max_resets`` times) or when the
    patience counter (+1 on a non-improving epoch, -1 floored at 0 on an improving one) exceeds ``patience``."""

    def __init__(self, model, patience: int, min_epochs: int, max_resets: int = 10):
        self.model, self.patience, self.min_epochs, self.max_resets = model, patience, min_epochs, max_resets
        self.current_patience = 0.0
        self.best_state = model.state_dict()
        self.best_loss = torch.inf
        self.resets = 0

    def __call__(self, epoch: int, loss) -> bool:
        if loss < self.best_loss:
            self.best_loss = loss
            self.best_state = deepcopy(self.model.state_dict())
            self.current_patience = max(self.current_patience - 1, 0)
        else:
            self.current_patience += 1
        if torch.isnan(loss):
            self.model.load_state_dict(self.best_state)
            self.resets += 1
            return self.resets > self.max_resets
        if epoch < self.min_epochs:
            return False
        if self.current_patience > self.patience:
            self.model.load_state_dict(self.best_state)
            return True
        return False


def run_train(model, nll_loss_class, zscore_layer_class, train_loader, validation_loader=None, test_loader=None, *,
              defense=None, early_stopping: bool = True, patience: int = 5, min_epochs: int = 5, max_epochs: int = 100,
              optimizer: Any = torch.optim.Adam, lr_scheduler=None, lr_scheduler_kwargs=None, lr: float = 1e-4,
              grad_clip_value: Optional[float] = None, defense_hyper_parameters: Optional[dict] = None,
              loss_fn_hyper_parameters: Optional[dict] = None, device="cpu", z_score_x: bool = True,
              train_in_unconstrained_space: bool = False, on_batch: Optional[Callable] = None):
    """NLL training of an amortized posterior, optionally under a defense (run_train.py:92-283).

    Order of operations kept from the reference: optimizer over ``model.parameters()`` first; the z-score layer is
    PREPENDED to ``model.embedding_net`` (mean / std over the training xs, std clamped at 0.05); train mode, then
    ``.to(device)``; the output transform is stripped for training in unconstrained space and restored at the end (theta
    mapped through its inverse per batch, but NOT for the test loss); ``scale_eps_beta_by_std`` multiplies the defense's
    eps / eps_iter by the mean per-row std of the training xs; additive defenses are constructed on (model, loss_fn) and
    activated; per batch: zero_grad, loss, backward, clip_grad_norm_, step, scheduler step; validation under no_grad in eval
    mode of the LOSS (the model stays in train mode); early stopping on the validation loss.
    ``on_batch(i_epoch, i_batch)`` is a test hook (base-noise injection), not part of the reference.
    """
    defense_hp = dict(defense_hyper_parameters or {})
    optim = optimizer(model.parameters(), lr=lr)
    scheduler = lr_scheduler(optim, **(lr_scheduler_kwargs or {})) if lr_scheduler is not None else None
    loss_function = nll_loss_class(model, **(loss_fn_hyper_parameters or {}))
    if z_score_x:
        X = torch.vstack([xy[0] for xy in train_loader.dataset])
        mean = torch.nan_to_num(X.mean(0), nan=0.0, posinf=0.0, neginf=0.0)
        std = torch.nan_to_num(X.std(0), nan=1.0, posinf=1.0, neginf=1.0).clamp(min=0.05)
        model.embedding_net = nn.Sequential(zscore_layer_class(mean=mean, std=std), model.embedding_net)
    loss_function.train()
    model.train()
    loss_function.to(device)
    model.to(device)
    output_transform = None
    if train_in_unconstrained_space:
        output_transform = model.output_transform
        if output_transform is None:
            train_in_unconstrained_space = False
        model.output_transform = None
    checker = EarlyStopping(model, patience=patience, min_epochs=min_epochs) if early_stopping else None
    defense_model = None
    if defense is not None:
        if defense_hp.pop("scale_eps_beta_by_std", False):
            X = torch.vstack([xy[0] for xy in train_loader.dataset])
            std_rows = float(X.std(1).mean())
            for key in ("eps", "eps_iter"):
                if key in defense_hp:
                    defense_hp[key] = defense_hp[key] * std_rows
        defense_model = defense(model=model, loss_fn=loss_function, **defense_hp)
        defense_model.activate()
    train_loss, validation_loss = [], []
    for i in range(max_epochs):
        epoch_loss, n = 0, 0
        loss_function.train()
        for x, theta in train_loader:
            x, theta = x.to(device), theta.to(device)
            if train_in_unconstrained_space:
                theta = output_transform.inv(theta)
            if on_batch is not None:
                on_batch(i, n)
            optim.zero_grad()
            loss = loss_function(x, theta)
            loss.backward()
            if grad_clip_value is not None:
                torch.nn.utils.clip_grad_norm_(model.parameters(), grad_clip_value)
            optim.step()
            if scheduler is not None:
                scheduler.step()
            n += 1
            epoch_loss += loss.detach().cpu()
        with torch.no_grad():
            loss_function.eval()
            train_loss.append(epoch_loss / n)
            if validation_loader is not None:
                n = 0
                loss_val = torch.zeros(1, device=device)
                for x_val, theta_val in validation_loader:
                    x_val, theta_val = x_val.to(device), theta_val.to(device)
                    if train_in_unconstrained_space:
                        theta_val = output_transform.inv(theta_val)
                    loss_val += loss_function(x_val, theta_val)
                    n += 1
                validation_loss.append(loss_val.cpu() / n)
        if checker is not None:
            if checker(i, validation_loss[i] if validation_loader is not None else train_loss[i]):
                break
    if train_in_unconstrained_space:
        model.output_transform = output_transform
    test_loss = None
    if test_loader is not None:
        with torch.no_grad():
            test_loss, n = 0.0, 0.0
            loss_function.eval()
            for x, theta in test_loader:
                test_loss += loss_function(x.to(device), theta.to(device)).cpu()
                n += 1
            test_loss /= n
    return model, defense_model, train_loss, (validation_loss if validation_loader is not None else None), test_loss
