"""ORACLE (test infrastructure): golden vectors for the reference RUNNERS on either side of the hot path.
Run in the BUILD container only (needs /root/reference):

    python -m oracle.gen_golden_runs        # writes tests/golden/runs_small.pt and tests/golden/ref_configs.json

1. The hydra yaml files the seam reads (model=maf_pyro, defense=fisher_trace / l2pgdTrades_rKL, eval_rob/attack=l2pgd,
   eval_rob/metric=rKL, eval_rob=untargeted, train=fKL) are copied VALUE BY VALUE into tests/golden/ref_configs.json, so the
   CPU seam test can resolve classes the way rbibm_script.py:597-636,753-764 does without /root/reference.
2. The UNMODIFIED ``rbibm.runs.run_rob_eval.run_rob_eval`` and ``rbibm.runs.run_train.run_train`` are executed over the
   unmodified ``rbi`` / ``rbibm.metric`` classes (third-party stubs underneath) with every base-noise draw and every PGD
   starting point recorded; the restatements in oracle/rbibm_runs_restated.py, run over the SAME classes and the recorded
   noise, must return bit-identical results.  Inputs, noise and the reference's outputs are stored for the GPU test that
   drives the restated runners over ``rabi_b200`` classes (tests/test_gpu_runs.py).
"""
from __future__ import annotations

import json
import os
import sys

import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import rbi_oracle as O  # noqa: E402
from oracle import rbibm_runs_restated as R  # noqa: E402
from oracle.gen_golden import _bitwise, _perturb_weights  # noqa: E402
from oracle.gen_golden_next import _SpyDelta0  # noqa: E402
from oracle.thirdparty import install_stubs  # noqa: E402

CFG_ROOT = "/root/reference/src/rbibm/config"
CFG_FILES = {
    "model": "model/maf_pyro.yaml",
    "defense_fisher_trace": "defense/fisher_trace.yaml",
    "defense_l2pgdTrades_rKL": "defense/l2pgdTrades_rKL.yaml",
    "attack_l2pgd": "eval_rob/attack/l2pgd.yaml",
    "metric_rKL": "eval_rob/metric/rKL.yaml",
    "eval_rob": "eval_rob/untargeted.yaml",
    "train": "train/fKL.yaml",
}
SHAPE = dict(dx=20, D=4, hidden=[32, 32], K=3, N_train=128, N_val=64, N_test=64, batch=64, N_rob=96)
MODEL_KW = dict(log_scale_min_clip=-7, log_scale_max_clip=5, stable=False)


def read_configs():
    out = {}
    for key, rel in CFG_FILES.items():
        with open(os.path.join(CFG_ROOT, rel)) as f:
            out[key] = yaml.safe_load(f)
        out[key]["__file__"] = "src/rbibm/config/" + rel
    return out


def make_data(seed=5):
    g = torch.Generator().manual_seed(seed)
    s = SHAPE
    low, high = -2.0 * torch.ones(s["D"]), 3.0 * torch.ones(s["D"])

    def draw(n):
        theta = low + (high - low) * torch.rand(n, s["D"], generator=g)
        x = torch.cat([theta.repeat(1, s["dx"] // s["D"])], 1) * torch.linspace(0.5, 2.0, s["dx"]) \
            + 0.3 * torch.randn(n, s["dx"], generator=g)
        return x, theta

    return dict(train=draw(s["N_train"]), val=draw(s["N_val"]), test=draw(s["N_test"]), rob=draw(s["N_rob"]), low=low, high=high)


def loaders(data):
    from torch.utils.data import DataLoader, TensorDataset

    mk = lambda k: DataLoader(TensorDataset(*data[k]), batch_size=SHAPE["batch"], shuffle=False)
    return mk("train"), mk("val"), mk("test")


def fresh_reference_model(data, seed=77):
    import rbi.models as RM
    from torch.distributions import biject_to, constraints

    torch.manual_seed(seed)
    tf = biject_to(constraints.independent(constraints.interval(data["low"], data["high"]), 1))
    m = RM.InverseAffineAutoregressiveModel(SHAPE["dx"], SHAPE["D"], hidden_dims=SHAPE["hidden"], num_transforms=SHAPE["K"],
                                            shuffle=False, output_transform=tf, **MODEL_KW)
    _perturb_weights(m, 3)
    return m


def state_of(m):
    return {k: (v.bool() if k.endswith(".mask") else v.detach().clone()) for k, v in m.state_dict().items()}


def main():
    cfgs = read_configs()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    with open(os.path.join(ROOT, "tests", "golden", "ref_configs.json"), "w") as f:
        json.dump(cfgs, f, indent=1, sort_keys=True)

    install_stubs.install_rbibm()
    import rbi.attacks.advertorch_attack as RAA
    import rbi.defenses as RD
    import rbi.loss as RL
    import rbibm.metric.robustness_metric as RMet
    from rbi.models.module import ZScoreLayer as RefZ
    from rbibm.runs.run_rob_eval import run_rob_eval as ref_run_rob_eval
    from rbibm.runs.run_train import run_train as ref_run_train

    data = make_data()
    out = {"shape": dict(SHAPE), "model_kw": dict(MODEL_KW), "low": data["low"], "high": data["high"],
           **{f"{k}_x": data[k][0] for k in ("train", "val", "test", "rob")},
           **{f"{k}_theta": data[k][1] for k in ("train", "val", "test", "rob")}}

    # ------------------------------------------------------------------ run_train, NLL only and under FIMTraceRegularizer(ema)
    train_hp = dict(cfgs["train"]["params"])
    train_hp.pop("loss_fn_hyper_parameters")
    train_hp.update(optimizer=torch.optim.Adam, lr_scheduler=None, max_epochs=3, min_epochs=1, patience=1,
                    initialize_as_prior=False)
    train_hp["lr"] = float(train_hp["lr"])
    out["train_hp"] = {k: v for k, v in train_hp.items() if k not in ("optimizer",)}
    for tag, defense_cls, defense_hp in (("nll", None, {}),
                                         ("fim", RD.FIMTraceRegularizer, dict(cfgs["defense_fisher_trace"]["params"]))):
        m0 = fresh_reference_model(data)
        out[f"train_{tag}_init"] = state_of(m0)
        tr, va, te = loaders(data)
        rec = []
        import copy

        m_ref = copy.deepcopy(m0)
        with O.recorded_base_noise(rec):
            m_ref, d_ref, tl, vl, tel = ref_run_train(None, m_ref, RL.NLLLoss, tr, va, te, defense=defense_cls,
                                                      defense_hyper_parameters=dict(defense_hp), device="cpu", verbose=False,
                                                      loss_fn_hyper_parameters={}, **train_hp)
        m_mine = copy.deepcopy(m0)
        hp2 = {k: v for k, v in train_hp.items() if k not in ("initialize_as_prior", "initialize_as_prior_rounds")}
        with O.injected_base_noise(list(rec)):
            m_mine, d_mine, tl2, vl2, tel2 = R.run_train(m_mine, RL.NLLLoss, RefZ, tr, va, te, defense=defense_cls,
                                                         defense_hyper_parameters=dict(defense_hp), device="cpu", **hp2)
        _bitwise(torch.stack(tl), torch.stack(tl2), f"run_train[{tag}]: train losses")
        _bitwise(torch.cat(vl), torch.cat(vl2), f"run_train[{tag}]: validation losses")
        _bitwise(tel.reshape(-1), tel2.reshape(-1), f"run_train[{tag}]: test loss")
        for (k, a), (_, b) in zip(m_ref.state_dict().items(), m_mine.state_dict().items()):
            _bitwise(a.float(), b.float(), f"run_train[{tag}]: final {k}")
        assert (m_ref.output_transform is not None) and isinstance(m_ref.embedding_net[0], RefZ)
        # the base noise the product's estimators will be fed: only draws with a sample dimension (the D normals each
        # ``model(x)`` burns -- utils/distributions.py:457 -- are not MC samples)
        mc = [z for z in rec if z.dim() == 3]
        out[f"train_{tag}"] = dict(train_loss=torch.stack(tl), val_loss=torch.cat(vl), test_loss=tel, final=state_of(m_ref),
                                   z=mc, defense_hp=dict(defense_hp), epochs=len(tl))
        print(f"run_train[{tag}]: {len(tl)} epochs, train {[round(float(t), 4) for t in tl]}, val {[round(float(t), 4) for t in vl]}, "
              f"test {float(tel):.4f}, MC draws {len(mc)}")

    # ------------------------------------------------------------------ run_rob_eval on the NLL-trained model
    m_eval = fresh_reference_model(data)
    m_eval.embedding_net = torch.nn.Sequential(RefZ(mean=data["train"][0].mean(0), std=data["train"][0].std(0).clamp(min=0.05)),
                                               m_eval.embedding_net)
    m_eval.load_state_dict({k: (v.float() if v.dtype == torch.bool else v) for k, v in out["train_nll"]["final"].items()})
    out["rob_model"] = state_of(m_eval)
    xs, thetas = data["rob"]
    from torch.utils.data import DataLoader, TensorDataset

    rob_loader = DataLoader(TensorDataset(xs, thetas), batch_size=40, shuffle=False)
    attack_cfg, metric_cfg, er = cfgs["attack_l2pgd"], cfgs["metric_rKL"], cfgs["eval_rob"]
    attack_params = dict(attack_cfg["params"])
    attack_params["nb_iter"] = 10          # the yaml's 200 iterations, shortened for the fixture (stated in the test)
    kw = dict(attack_attemps=er["attack_attemps"], attack_mc_budget=attack_cfg["attack_mc_budget"], eval_mc_budget=32,
              eps=er["eps"], attack_hyper_parameters=attack_params, metric_hyper_parameters=dict(metric_cfg["params"]),
              num_adversarial_examples=7, batch_size=32, eps_relative_to_std=er["eps_relative_to_std"],
              clip_min_max_automatic=er["clip_min_max_automatic"])
    rec = []
    torch.manual_seed(31)
    with _SpyDelta0() as spy, O.recorded_base_noise(rec):
        (name, value, additional, x_uns, th_uns, x_adv, _c2st, eps_abs) = ref_run_rob_eval(
            None, m_eval, RMet.ReverseKLRobMetric, RAA.L2PGDAttack, test_loader=rob_loader, attack_loss_fn=RL.ReverseKLLoss,
            targeted=False, target_strategy=None, device="cpu", verbose=False, **kw)
    d0 = list(spy.d0)

    class _Replay:   # serve the recorded PGD starting points to the restated run
        def __enter__(self):
            self._orig = O.adv.rand_init_delta
            q = list(d0)

            def fake(delta, x, ord, eps, clip_min, clip_max):
                delta.data = q.pop(0).clone()
                return delta.data

            O.adv.rand_init_delta = fake
            return self

        def __exit__(self, *exc):
            O.adv.rand_init_delta = self._orig
            return False

    with _Replay(), O.injected_base_noise(list(rec)):
        mine = R.run_rob_eval(m_eval, RMet.ReverseKLRobMetric, RAA.L2PGDAttack, xs, thetas, attack_loss_fn=RL.ReverseKLLoss,
                              device="cpu", **kw)
    assert mine["metric_name"] == name and mine["eps_abs"] == eps_abs
    _bitwise(torch.as_tensor(value).reshape(-1), torch.as_tensor(mine["value"]).reshape(-1), "run_rob_eval: value")
    _bitwise(x_adv, mine["x_adversarial"], "run_rob_eval: adversarial examples")
    _bitwise(x_uns, mine["x_unsecure"], "run_rob_eval: x_unsecure")
    _bitwise(th_uns, mine["theta_unsecure"], "run_rob_eval: theta_unsecure")
    nb = attack_params["nb_iter"]
    chunks = [c.shape[0] for c in torch.split(xs, kw["batch_size"])]
    assert len(d0) == len(chunks) and len([z for z in rec if z.dim() == 3]) == len(chunks) * (nb + 1)
    mc = [z for z in rec if z.dim() == 3]
    out["rob"] = dict(kw=kw, name=name, value=value, additional=additional, x_unsecure=x_uns, theta_unsecure=th_uns,
                      x_adversarial=x_adv, eps_abs=eps_abs, attack_params=mine["attack_params"], x_idx=mine["x_idx"],
                      delta0=torch.cat(d0, 0),
                      z_attack=[torch.cat([mc[c * (nb + 1) + t] for c in range(len(chunks))], 1) for t in range(nb)],
                      z_eval=torch.cat([mc[c * (nb + 1) + nb] for c in range(len(chunks))], 1))
    print(f"run_rob_eval: {name} value {value}, eps_abs {eps_abs:.4f}, attack params {mine['attack_params']}")
    torch.save(out, os.path.join(ROOT, "tests", "golden", "runs_small.pt"))
    print("wrote tests/golden/runs_small.pt", os.path.getsize(os.path.join(ROOT, "tests", "golden", "runs_small.pt")) // 1024, "KiB")


if __name__ == "__main__":
    main()
