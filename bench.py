#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: attacked (x, MC-sample) MAF posterior evals/sec.

One "step" = one pass of the robustness hot path over one shard of synthetic observations, exactly the work
``rbibm ... eval_rob/attack=l2pgd eval_rob/metric=rKL`` does per epsilon: for every observation row an
L2-PGD attack (nb_iter=200 iterations x S=5 MC samples, reverse-KL loss) followed by the robustness metric
KL(q(.|x~) || q(.|x)) with S=256 samples.  Workload = BASELINE config 2 (lotka_volterra shapes: dx=C=100, D=4,
hidden [100,100], K=3, z-score embedding), 10 000 observations per GPU => 12.56 M (x, MC-sample) pairs per step per
GPU.  Data are synthetic (x ~ N(0, I), seed 0; default nn.Linear init perturbed with seeded noise) because the
simulators need packages that are not installed and there is no network.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched by the driver under torchrun, one rank per GPU; observations are sharded by rank (weak scaling,
no data-path collective); the only exchange is the final all_gather of the per-observation losses (NCCL).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]
    "lotka_volterra_l2pgd_rkl": dict(dx=100, D=4, hidden=[100, 100], K=3, rows=10000, chunk=1000, nb_iter=200,
                                     S_att=5, S_eval=256),
    # not bench lines of their own (parity-test shapes); selectable for the record with --workload
    "gaussian_linear_l2pgd_rkl": dict(dx=10, D=10, hidden=[100, 100], K=3, rows=10000, chunk=1000, nb_iter=200,
                                      S_att=5, S_eval=256),
    # BASELINE configs[4], ctx-space variant: h ~ N(0, I_100) given, 12 500 rows per GPU
    "pyloric_ctx_l2pgd_rkl": dict(dx=100, D=31, hidden=[200, 200, 200], K=3, rows=12500, chunk=500, nb_iter=200,
                                  S_att=5, S_eval=256),
    # x-space variant: raw 2400-sample traces through a stand-in 1-D CNN embedding (torch) -> 100 features;
    # the PGD gradient chains through the embedding by autograd, the flow runs on the kernels
    "pyloric_x_l2pgd_rkl": dict(dx=2400, D=31, hidden=[200, 200, 200], K=3, rows=12500, chunk=500, nb_iter=200,
                                S_att=5, S_eval=256, embed="cnn1d"),
}
METRIC = "attacked (x, MC-sample) MAF posterior evals/sec"
UNIT = "pairs/s"


def pairs_per_row(w):
    return w["nb_iter"] * w["S_att"] + w["S_eval"]


# ------------------------------------------------------------------------------------------------ model / data
def make_embedding(w):
    """Stand-in for sbi's CNNEmbedding (absent): 1-D CNN, dx samples -> 100 summary features."""
    if w.get("embed") != "cnn1d":
        return None
    nn = torch.nn
    dx = w["dx"]
    return nn.Sequential(nn.Unflatten(1, (1, dx)), nn.Conv1d(1, 8, 9, stride=4, padding=4), nn.ReLU(),
                         nn.Conv1d(8, 16, 9, stride=4, padding=4), nn.ReLU(), nn.Flatten(),
                         nn.Linear(16 * (dx // 16), 100))


def make_weights(w, seed=0):
    """state_dict of the MAF (reference layout) + z-score stats + observations; CPU tensors."""
    from rabi_b200.models import InverseAffineAutoregressiveModel

    torch.manual_seed(seed)
    m = InverseAffineAutoregressiveModel(w["dx"], w["D"], hidden_dims=w["hidden"], num_transforms=w["K"], shuffle=False,
                                         log_scale_min_clip=-7, log_scale_max_clip=5, stable=False,
                                         embedding_net=make_embedding(w))
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():  # non-trivial means / log-scales (plain init gives a nearly Gaussian posterior)
        for p in m.parameters():
            p.add_(0.25 * torch.randn(p.shape, generator=g) * (p.abs().mean() + 0.05))
    return m


def make_rows(w, rank, n_rows):
    g = torch.Generator().manual_seed(1000 + rank)
    return torch.randn(n_rows, w["dx"], generator=g)


def algorithmic_flops(model, w):
    """FLOPs per attack / eval pair as SURVEY 8d defines them: masked MACs only, sampling = one pass-equivalent.
    X = context projection (shared by the S samples of a row), R = rest (per sample)."""
    X = R = 0
    for k in range(w["K"]):
        layers = getattr(model, f"t{k}").nn.layers
        C = model.context_dim
        m0 = layers[0].mask
        X += int(m0[:, :C].sum())
        R += int(m0[:, C:].sum()) + sum(int(l.mask.sum()) for l in layers[1:])
    att = 2.0 * (4 * R + 2 * X / w["S_att"])
    ev = 2.0 * (2 * R + 2 * X / w["S_eval"])
    pair_kernel_att = 2.0 * 4 * R  # what one launch of the per-pair fwd+bwd kernel covers (X is in the projection kernels)
    return dict(attack_pair=att, eval_pair=ev, pair_kernel_attack=pair_kernel_att, X=X / w["K"], R=R / w["K"])


def bench_config(wname, w, world, flops):
    """The ``config`` object of the bench line; the reference arm prints the SAME object (same workload, same sizes)."""
    return {"workload": wname, "rows_per_gpu": w["rows"], "chunk": w["chunk"], "nb_iter": w["nb_iter"],
            "S_att": w["S_att"], "S_eval": w["S_eval"], "dx": w["dx"], "D": w["D"], "hidden": w["hidden"], "K": w["K"],
            "pairs_per_step": w["rows"] * pairs_per_row(w) * world * max(1, int(w.get("eps_sweep", 1))), "eps_values_batched": max(1, int(w.get("eps_sweep", 1))),
            "parallelism": f"rows sharded over {world} GPU(s), no data-path collective",
            "l2": "per-step inputs (fresh base noise, 160 MB) exceed the 126 MB L2",
            "flop_per_attack_pair": flops["attack_pair"], "flop_per_eval_pair": flops["eval_pair"]}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [s.strip() for s in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_reference_sample(w, model_state, threads, it_sample=10, seed=0):
    """Times the oracle (op-for-op restatement of the reference's Pyro/advertorch path) on one 1000-row chunk:
    ``it_sample`` PGD iterations + the S_eval metric, and extrapolates to the chunk's full nb_iter (iterations are
    identical in cost; chunks are independent).  Returns pairs/s and a description."""
    from oracle import rbi_oracle as O

    torch.set_num_threads(threads)
    ref = O.OracleMAF(w["dx"], w["D"], hidden_dims=w["hidden"], num_transforms=w["K"], shuffle=False,
                      embedding_net=torch.nn.Sequential(O.ZScoreLayer(torch.zeros(w["dx"]), torch.ones(w["dx"])),
                                                        make_embedding(w) or torch.nn.Identity()),
                      log_scale_min_clip=-7, log_scale_max_clip=5, stable=False)
    sd = {k: v for k, v in model_state.items()}
    ref.load_state_dict(sd, strict=True)
    ref.eval()
    x = make_rows(w, 0, w["chunk"])
    eps = float(0.5 * x.std(1).mean())
    torch.manual_seed(seed)
    t0 = time.perf_counter()
    xa = O.l2pgd_perturb(ref, x, O.ReverseKLLoss(mc_samples=w["S_att"]), eps, it_sample, eps / 10, float(x.min()),
                         float(x.max()))
    t_att = time.perf_counter() - t0
    t0 = time.perf_counter()
    O.rkl_rob_metric(ref, x, xa, w["S_eval"])
    t_eval = time.perf_counter() - t0
    t_chunk = t_att / it_sample * w["nb_iter"] + t_eval
    pairs = w["chunk"] * pairs_per_row(w)
    sample = (f"oracle (PyTorch restatement of the reference path) on one {w['chunk']}-row chunk: {it_sample} of "
              f"{w['nb_iter']} PGD iterations ({t_att:.2f} s) + S={w['S_eval']} metric ({t_eval:.2f} s)"
              + ("" if it_sample == w["nb_iter"] else f", extrapolated linearly to {w['nb_iter']} iterations"))
    return pairs / t_chunk, sample, t_att + t_eval


def run_reference_arm(args, w, wname):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if os.environ.get("OMP_NUM_THREADS") == "1" and not os.environ.get("RABI_REF_CHILD"):
        # torchrun exports OMP_NUM_THREADS=1, which pins the CPU arm to one thread no matter what
        # torch.set_num_threads says later: re-run it in a clean environment so it can use all host threads
        env = {k: v for k, v in os.environ.items()
               if k not in ("OMP_NUM_THREADS", "MKL_NUM_THREADS") and not k.startswith(("TORCHELASTIC", "GROUP_", "ROLE_"))}
        env["RABI_REF_CHILD"] = "1"
        for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE"):
            env.pop(k, None)
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--gpus", str(args.gpus),
                              "--steps", str(args.steps), "--warmup", str(args.warmup), "--workload", wname],
                             env=env, capture_output=True, text=True)
        sys.stderr.write(out.stderr[-2000:])
        emit(out.stdout.strip().splitlines()[-1] if out.stdout.strip() else
             json.dumps({"impl": "reference", "unavailable": "reference arm subprocess failed"}))
        return
    from rabi_b200.models import ZScoreLayer

    model = make_weights(w)
    model.embedding_net = torch.nn.Sequential(ZScoreLayer(torch.zeros(w["dx"]), torch.ones(w["dx"])), model.embedding_net)
    threads = os.cpu_count() or 1
    flops = algorithmic_flops(model, w)
    # One timed step = ONE WHOLE chunk of the workload through the reference path: all nb_iter PGD iterations + the S_eval
    # metric (lotka_volterra: ~7 s of CPU work) -- a measurement, not an extrapolation.  The warm-up steps only have to spin
    # up the thread pool and the allocator, so they use 3 iterations.  Conditioners whose whole chunk takes minutes (pyloric
    # shapes: ~80 s) time a bounded number of iterations per step and say so.
    full = len(w["hidden"]) == 2
    it_timed = w["nb_iter"] if full else 20
    vals, secs = [], []
    for i in range(args.warmup + args.steps):
        timed_step = i >= args.warmup
        v, sample, dt = cpu_reference_sample(w, model.state_dict(), threads, it_sample=it_timed if timed_step else 3, seed=i)
        if timed_step:
            vals.append(v)
            secs.append(dt)
    pairs_chunk = w["chunk"] * pairs_per_row(w)
    value = pairs_chunk * len(secs) / sum(secs) if full else sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(wname, w, 1, flops),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "pairs_per_timed_step": pairs_chunk if full else w["chunk"] * (it_timed * w["S_att"] + w["S_eval"]),
                         "extrapolated": not full},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference (pyro/advertorch/sbi) cannot be installed here; this arm times the oracle port of its "
                "CPU path on all host threads; a timed step is one whole chunk of the config's rows (the CPU walks the "
                "rows chunk by chunk at the same cost per chunk), value = pairs of the timed chunks / their wall time",
    }
    emit(json.dumps(line))



# ------------------------------------------------------------------------------------------------ SURVEY 8f rows
def run_next_rows(args):
    """Measurements for the SURVEY 8f "next" rows at lotka_volterra shapes on ONE GPU, each beside the oracle port of the
    reference path timed on the host cores (bounded samples).  Prints its own JSON document (not the bench line)."""
    from oracle import rbi_oracle as O

    from rabi_b200 import _lib
    from rabi_b200.attacks import L2PGDAttack
    from rabi_b200.defenses import L2PGDAdversarialTraining, L2PGDTrades, L2UniformNoiseTraining
    from rabi_b200.loss import ForwardKLLoss, NLLLoss
    from rabi_b200.metric import ExpectedCoverageMetric, ForwardKLRobMetric, NegativeLogLikelihoodMetric
    from rabi_b200.models import ZScoreLayer

    w = dict(WORKLOADS["lotka_volterra_l2pgd_rkl"])
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    _lib.build()
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    model = make_weights(w)
    state = state_cpu_with_zscore({k: v.clone() for k, v in model.state_dict().items()}, w)
    model.embedding_net = torch.nn.Sequential(ZScoreLayer(torch.zeros(w["dx"]), torch.ones(w["dx"])), model.embedding_net)
    ref = O.OracleMAF(w["dx"], w["D"], hidden_dims=w["hidden"], num_transforms=w["K"], shuffle=False,
                      embedding_net=torch.nn.Sequential(O.ZScoreLayer(torch.zeros(w["dx"]), torch.ones(w["dx"])),
                                                        torch.nn.Identity()),
                      log_scale_min_clip=-7, log_scale_max_clip=5, stable=False)
    ref.load_state_dict(state)
    ref.eval()
    model = model.to(dev).eval()
    x = make_rows(w, 0, w["rows"])
    xd = x.to(dev)
    eps = float(0.5 * x.std(1).mean())
    cmin, cmax = float(x.min()), float(x.max())
    out = {"shapes": "lotka_volterra", "host_threads": threads, "rows": w["rows"]}

    def gpu_ms(fn, n=3, warm=1):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def cpu_s(fn):
        t0 = time.perf_counter()
        fn()
        return time.perf_counter() - t0

    # f1: forward-KL attack + ForwardKLRobMetric
    atk = L2PGDAttack(model, loss_fn=ForwardKLLoss(mc_samples=w["S_att"]), eps=eps, nb_iter=w["nb_iter"],
                      eps_iter=eps / 10, clip_min=cmin, clip_max=cmax)
    met = ForwardKLRobMetric(model, atk, mc_samples=w["S_eval"], attack_attemps=1, device=dev, batch_size=w["chunk"],
                             reduction=None)

    def f1():
        met._cached_x = met._cached_xs_tilde = met._cached_losses = None
        met._eval(xd)

    ms = gpu_ms(f1, 2)
    pairs = w["rows"] * pairs_per_row(w)
    xc = x[: w["chunk"]]
    it = 5
    t_att = cpu_s(lambda: O.l2pgd_perturb(ref, xc, O.ForwardKLLoss(mc_samples=w["S_att"]), eps, it, eps / 10, cmin, cmax))
    t_ev = cpu_s(lambda: O.ForwardKLLoss(mc_samples=w["S_eval"], reduction=None)(ref(xc + 0.01), ref(xc)))
    cpu_rate = w["chunk"] * pairs_per_row(w) / (t_att / it * w["nb_iter"] + t_ev)
    out["f1_fkl_attack_and_metric"] = {"gpu_pairs_per_s": pairs / (ms * 1e-3), "gpu_ms": ms, "cpu_pairs_per_s": cpu_rate,
                                       "cpu_sample": f"{w['chunk']} rows, {it} of {w['nb_iter']} iterations + metric"}
    # f2: approximation metrics on 10 000 rows (coverage: 1000 posterior samples per row)
    theta = torch.randn(w["rows"], w["D"])
    mc = 1000
    cov = ExpectedCoverageMetric(model, mc_samples=mc, device=dev, batch_size=w["chunk"])
    ms = gpu_ms(lambda: cov.eval(x, theta))
    n_cpu = 200
    t_c = cpu_s(lambda: O.expected_coverage(ref, x[:n_cpu], theta[:n_cpu], mc, torch.linspace(0, 1, 100)))
    out["f2_expected_coverage"] = {"gpu_samples_per_s": w["rows"] * mc / (ms * 1e-3), "gpu_ms": ms,
                                   "cpu_samples_per_s": n_cpu * mc / t_c, "cpu_sample": f"{n_cpu} rows x {mc} samples"}
    nllm = NegativeLogLikelihoodMetric(model, reduction="mean", device=dev)
    ms = gpu_ms(lambda: nllm.eval(x, theta))
    t_c = cpu_s(lambda: O.nll_metric(ref, x, theta))
    out["f2_nll_metric"] = {"gpu_rows_per_s": w["rows"] / (ms * 1e-3), "gpu_ms": ms, "cpu_rows_per_s": w["rows"] / t_c}
    # f3: training steps, batch 512 (config/defense/l2pgdAdvTrain.yaml, l2noiseAdvTrain.yaml, l2pgdTrades.yaml)
    B = 512
    xb, tb = xd[:B], theta[:B].to(dev)
    model.train()
    ref.train()
    std = float(x.std(1).mean())
    for name, make, cpu_fn in (
        ("f3_l2pgd_adversarial_training",
         lambda lf: L2PGDAdversarialTraining(model, lf, eps=0.5 * std, eps_iter=0.1 * std, nb_iter=20),
         lambda: O.train_loss_with_pre_regularizer(ref, x[:B], theta[:B], lambda xi, ti: (O.l2pgd_perturb(
             ref, xi, O.ForwardKLLoss(mc_samples=1), 0.5 * std, 20, 0.1 * std).detach(), ti))[0].backward()),
        ("f3_l2_uniform_noise_training",
         lambda lf: L2UniformNoiseTraining(model, lf, eps=1.0 * std),
         lambda: O.train_loss_with_pre_regularizer(ref, x[:B], theta[:B], lambda xi, ti: (
             O.l2_uniform_noise_perturb(xi, 1.0 * std), ti))[0].backward()),
        ("f3_l2pgd_trades_fkl",
         lambda lf: L2PGDTrades(model, lf, eps=0.5 * std, eps_iter=0.1 * std, nb_iter=20, beta=0.1),
         lambda: O.train_loss_with_regularizer(ref, x[:B], theta[:B], lambda xi, qi, ti: O.trades_regularizer(
             ref, xi, 0.1, 0.5 * std, 0.1 * std, 20, O.ForwardKLLoss, 1)[0]).backward()),
    ):
        lf = NLLLoss(model).train()
        d = make(lf)
        d.activate()

        def step():
            model.zero_grad(set_to_none=True)
            lf(xb, tb).backward()

        ms = gpu_ms(step, 10, warm=6)  # the first calls of a batch shape run eagerly and then capture CUDA graphs
        d.deactivate()
        ref.zero_grad()
        cpu_fn()
        t_c = cpu_s(cpu_fn)
        out[name] = {"gpu_ms_per_step": ms, "cpu_ms_per_step": 1e3 * t_c, "batch": B}
    emit(json.dumps(out))

# ------------------------------------------------------------------------------------------------ GPU arm
_RESULT_FD = None


def emit(text):
    """Write a result line to the process' original stdout (see main)."""
    if _RESULT_FD is None:
        print(text, flush=True)
    else:
        os.write(_RESULT_FD, (text + "\n").encode())


def tensor_fraction(model, w):
    """Share of the per-pair kernel's masked MACs that run as tcgen05.mma products: the hidden -> hidden products (and, on
    the three-layer path, the hidden -> output product); layer 0 (K = D) and the own-degree blocks stay FMA code."""
    tc = tot = 0
    for k in range(w["K"]):
        layers = getattr(model, f"t{k}").nn.layers
        C = model.context_dim
        tot += int(layers[0].mask[:, C:].sum()) + sum(int(l.mask.sum()) for l in layers[1:])
        tc += sum(int(l.mask.sum()) for l in layers[1:-1])
        if len(w["hidden"]) == 3:
            tc += int(layers[-1].mask.sum())
    return tc / max(tot, 1)


def gpu_eager_sample(w, model_state, dev, it_sample=5):
    """The oracle (op-for-op restatement of the reference's Pyro / advertorch path) run EAGERLY on the same B200 -- what
    ``device: cuda`` gave the reference's authors: one chunk, ``it_sample`` PGD iterations + the S_eval metric, extrapolated."""
    from oracle import rbi_oracle as O

    ref = O.OracleMAF(w["dx"], w["D"], hidden_dims=w["hidden"], num_transforms=w["K"], shuffle=False,
                      embedding_net=torch.nn.Sequential(O.ZScoreLayer(torch.zeros(w["dx"]), torch.ones(w["dx"])),
                                                        make_embedding(w) or torch.nn.Identity()),
                      log_scale_min_clip=-7, log_scale_max_clip=5, stable=False)
    ref.load_state_dict(dict(model_state), strict=True)
    ref = ref.to(dev).eval()
    x = make_rows(w, 0, w["chunk"]).to(dev)
    eps = float(0.5 * x.std(1).mean())
    cmin, cmax = float(x.min()), float(x.max())
    loss = O.ReverseKLLoss(mc_samples=w["S_att"])
    O.l2pgd_perturb(ref, x, loss, eps, 1, eps / 10, cmin, cmax)   # warm-up (cuBLAS handles, allocator)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    xa = O.l2pgd_perturb(ref, x, loss, eps, it_sample, eps / 10, cmin, cmax)
    torch.cuda.synchronize()
    t_att = time.perf_counter() - t0
    t0 = time.perf_counter()
    O.rkl_rob_metric(ref, x, xa, w["S_eval"])
    torch.cuda.synchronize()
    t_eval = time.perf_counter() - t0
    t_chunk = t_att / it_sample * w["nb_iter"] + t_eval
    return {"value": w["chunk"] * pairs_per_row(w) / t_chunk, "unit": UNIT, "kind": "port",
            "sample": f"oracle in eager PyTorch on cuda:{dev.index}: one {w['chunk']}-row chunk, {it_sample} of {w['nb_iter']} PGD "
                      f"iterations ({t_att:.2f} s) + S={w['S_eval']} metric ({t_eval:.2f} s), extrapolated linearly"}


def measure(wname, w, args, steps, warmup, dev, world, rank, local, scaling, with_baselines):
    """One bench record for workload ``wname`` (rows PER GPU already set in ``w``)."""
    import torch.distributed as dist

    from rabi_b200 import _lib
    from rabi_b200.attacks import L2PGDAttack
    from rabi_b200.loss import ReverseKLLoss
    from rabi_b200.metric import ReverseKLRobMetric
    from rabi_b200.models import ZScoreLayer

    model = make_weights(w)
    model.embedding_net = torch.nn.Sequential(ZScoreLayer(torch.zeros(w["dx"]), torch.ones(w["dx"])), model.embedding_net)
    state_cpu = {k: v.clone() for k, v in model.state_dict().items()}
    flops = algorithmic_flops(model, w)
    tfrac = tensor_fraction(model, w)
    model = model.to(dev).eval()
    x_host = make_rows(w, rank, w["rows"]).pin_memory()
    x_dev = x_host.to(dev)
    eps = float(0.5 * x_host.std(1).mean())
    cmin, cmax = float(x_host.min()), float(x_host.max())

    attack = L2PGDAttack(model, loss_fn=ReverseKLLoss(mc_samples=w["S_att"]), eps=eps, nb_iter=w["nb_iter"],
                         eps_iter=eps / 10, clip_min=cmin, clip_max=cmax, targeted=False)
    metric = ReverseKLRobMetric(model, attack, mc_samples=w["S_eval"], attack_attemps=1, device=dev,
                                batch_size=w["chunk"], reduction=None)
    eng = model.engine()

    n_eps = max(1, int(getattr(args, "eps_sweep", 1)))
    sweep = [0.1, 0.2, 0.3, 0.5, 1.0, 2.0][:n_eps] if n_eps > 1 else None   # config/experiment/eval_l2_pyloric.yaml:13-16

    def run(xd):
        """one pass of the hot path over this rank's rows: attack + metric (at every radius of the sweep, batched, if asked)"""
        metric._cached_x = metric._cached_xs_tilde = metric._cached_losses = None
        if sweep is None:
            return metric._eval(xd)
        return torch.stack([v for v in metric.eval_sweep(xd, sweep).values()]).reshape(-1)

    def step_device():
        losses = run(x_dev)
        if world > 1:  # the path's one exchange step: gather the per-observation losses for the quantile summary
            out = [torch.empty_like(losses) for _ in range(world)]
            dist.all_gather(out, losses)
            losses = torch.cat(out)
        return losses

    loss_host = torch.empty(w["rows"] * world * n_eps, dtype=torch.float32).pin_memory()
    xadv_host = torch.empty_like(x_host).pin_memory()

    def step_e2e():
        xd = x_host.to(dev, non_blocking=True)
        losses = run(xd)
        if world > 1:
            out = [torch.empty_like(losses) for _ in range(world)]
            dist.all_gather(out, losses)
            losses = torch.cat(out)
        loss_host.copy_(losses, non_blocking=True)
        if sweep is None:
            xadv_host.copy_(metric._cached_xs_tilde, non_blocking=True)
        torch.cuda.synchronize()
        return loss_host

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(n):
            fn()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    fma_peak = float(_lib.lib().rabi_fma_peak_tflops(local, 5))
    for _ in range(warmup):
        step_device()
    barrier()
    launches0 = eng.launches
    with ClockSampler(local) as clocks:   # `value` and `e2e` are both timed with the per-launch event records OFF
        ms = timed(step_device, steps)
    launches = eng.launches - launches0
    step_e2e()
    ms_e2e = timed(step_e2e, steps)
    # separate pass for the dominant kernel's average launch duration (CUDA events around every launch on its stream)
    # (attacks that run as two row lanes on two streams -- rabi_b200/engine.py: pgd_run -- are profiled as ONE lane: a kernel
    #  timed while the other lane's kernel shares the GPU would not be a launch duration)
    n_prof = min(steps, 3)
    lanes_env = os.environ.get("RABI_PGD_LANES")
    os.environ["RABI_PGD_LANES"] = "1"
    try:
        eng.profile(True)
        ms_prof = timed(step_device, n_prof)
        kern_ms, kern_n = eng.profile_read()
        eng.profile(False)
    finally:
        if lanes_env is None:
            os.environ.pop("RABI_PGD_LANES", None)
        else:
            os.environ["RABI_PGD_LANES"] = lanes_env

    pairs_step = w["rows"] * pairs_per_row(w) * world * n_eps
    value = pairs_step * steps / (ms * 1e-3)
    e2e = pairs_step * steps / (ms_e2e * 1e-3)

    # roofline of the dominant kernel: per-pair rKL forward + reverse (one launch per PGD iteration)
    pairs_launch = w["rows"] * w["S_att"] * n_eps
    # the persistent attack kernel covers all nb_iter iterations (and the context projection / back-projection products,
    # 2 x 2 K X FLOP per row and iteration) in ONE launch; the per-iteration path launches the per-pair kernel nb_iter times
    iters_per_launch = max(1, round(w["nb_iter"] * n_prof / max(kern_n, 1)))
    flop_launch = pairs_launch * flops["pair_kernel_attack"] * iters_per_launch
    if iters_per_launch > 1:
        flop_launch += iters_per_launch * w["rows"] * n_eps * 4.0 * w["K"] * flops["X"]
    achieved = flop_launch / (kern_ms / max(kern_n, 1) * 1e-3) / 1e12 if kern_n else None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    three = len(w["hidden"]) == 3
    kernel_name = "k_tcd<GRAD>" if three else "k_tc<GRAD>"
    traffic = traffic_src = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed `ncu --set full` capture of this kernel
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[kernel_name]
        if t["pairs_per_launch"] == pairs_launch and t.get("pgd_iterations_per_launch", 1) == iters_per_launch:
            traffic, traffic_src = t["dram_bytes_per_launch"], f"ncu capture at commit {t['commit']} ({t['file']})"
    except Exception:
        pass
    # ceiling: the tensor-eligible share of the MACs at the 3xTF32 rate of the tensor pipe, the rest at the fp32 FMA rate
    bf16 = peaks.get("bf16_tflops_sustained") or 1386.3
    tc_peak = bf16 / 2.0 / 3.0      # TF32 runs at half the bf16 rate; one fp32-accurate product = 3 TF32 MMAs
    ceiling = 1.0 / (tfrac / tc_peak + (1.0 - tfrac) / fma_peak) if fma_peak > 0 else None
    roofline = {
        "bound": "tensor", "kernel": f"{kernel_name} (per-pair rKL forward + reverse on tcgen05.mma kind::tf32 (3xTF32) + FMA, " + (
            "one launch per PGD iteration)" if iters_per_launch == 1 else
            f"persistent: ONE launch for all {iters_per_launch} PGD iterations incl. projection, back-projection and update)"),
        "achieved": achieved, "peak": ceiling, "unit": "TFLOP/s",
        "frac": (achieved / ceiling) if (achieved and ceiling) else None, "traffic": traffic, "traffic_source": traffic_src,
        # per launch: base noise of every iteration; per-iteration path: + A_p, A_q in, gA out; persistent kernel: A_q + x, delta once
        "algorithmic_bytes_per_launch": int(4 * n_eps * (iters_per_launch * w["rows"] * w["S_att"] * w["D"] + (
            3 * w["K"] * w["hidden"][0] * w["rows"] if iters_per_launch == 1 else
            w["K"] * w["hidden"][0] * w["rows"] + 3 * w["rows"] * w["dx"]))),
        "peak_source": f"{tfrac:.3f} of the kernel's algorithmic MACs are tensor-core products: ceiling = 1 / (share / (bf16_tflops_sustained "
                       f"{bf16} [MEASURED_PEAKS.json] / 2 [TF32] / 3 [3xTF32]) + (1 - share) / fp32 FMA peak {fma_peak:.1f} "
                       "[rabi_fma_peak_tflops, register-only FFMA microbenchmark run in this process])",
        "tensor_share_of_macs": tfrac, "fp32_fma_peak": fma_peak,
        "frac_of_fp32_fma_peak": (achieved / fma_peak) if (achieved and fma_peak > 0) else None,
        "flop_per_launch": flop_launch, "pgd_iterations_per_launch": iters_per_launch,
        "kernel_ms_avg": kern_ms / max(kern_n, 1), "kernel_launches_timed": kern_n,
        "kernel_share_of_step": kern_ms / ms_prof if ms_prof else None,
        "whole_step_tflops": (w["rows"] * (w["nb_iter"] * w["S_att"] * flops["attack_pair"] + w["S_eval"] * flops["eval_pair"])
                              * steps / (ms * 1e-3) / 1e12),
        "hbm_gbs_measured_peak": peaks.get("hbm_gbs"),
    }
    cpu = gpu_eager = None
    if rank == 0 and world == 1 and with_baselines:
        v, sample, _ = cpu_reference_sample(w, state_cpu, os.cpu_count() or 1, it_sample=30 if not three else 10)
        cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": sample}
        try:
            gpu_eager = gpu_eager_sample(w, state_cpu, dev)
        except Exception as e:  # the comparator must never take the bench line down
            gpu_eager = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": bench_config(wname, w, world, flops),
        "clocks": clocks.summary(),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(x_host.numel() * 4),
                "d2h_bytes_per_step": int(loss_host.numel() * 4 + xadv_host.numel() * 4), "ms_per_step": ms_e2e / steps},
        "gpu_launches": int(launches),
        # 2: the attack's rows ran as two halves on two streams (engine.pgd_run: fills the ragged last round of a launch per
        # iteration); the roofline's kernel time was taken in a separate single-lane pass
        "attack_row_lanes": int(eng._lanes(w["S_att"], w["rows"] * n_eps)),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "gpu_eager_baseline": gpu_eager,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="rabi_b200", choices=["rabi_b200", "reference"])
    ap.add_argument("--workload", default="lotka_volterra_l2pgd_rkl", choices=list(WORKLOADS))
    ap.add_argument("--rows", type=int, default=None, help="observations per GPU (default: the workload's)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the workload's rows on EVERY GPU; strong: the workload's rows in total, split over the GPUs")
    ap.add_argument("--eps-sweep", type=int, default=1,
                    help="attack every row at this many radii of the reference's eps sweep in ONE batched call (rows x eps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config5", action="store_true",
                    help="skip the second record (BASELINE config 5: pyloric ctx-space) carried under 'also' in the same JSON line")
    ap.add_argument("--next-rows", action="store_true",
                    help="measure the SURVEY 8f rows (fKL eval, coverage / NLL metrics, defenses) instead of the bench line")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: whatever libraries print while the bench runs (e.g. NCCL's version banner under
    # torchrun) is sent to stderr, and the result lines below are written to the saved descriptor
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.next_rows:
        return run_next_rows(args)
    w = dict(WORKLOADS[args.workload])
    if args.rows:
        w["rows"] = args.rows
    if args.impl == "reference":
        return run_reference_arm(args, w, args.workload)
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist

    from rabi_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl rabi_b200) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.build()
    if args.scaling == "strong":
        w["rows"] = w["rows"] // world
    w["eps_sweep"] = args.eps_sweep
    line = measure(args.workload, w, args, args.steps, args.warmup, dev, world, rank, local, args.scaling,
                   with_baselines=not args.no_cpu_baseline)
    if args.workload == "lotka_volterra_l2pgd_rkl" and not args.no_config5:
        # BASELINE config 5 (pyloric shapes, ctx-space, 100 000 observations over 8 GPUs = 12 500 per GPU) in front of the driver:
        # a second, shorter measurement in the same run, carried under "also" (stdout keeps exactly one JSON line)
        w5 = dict(WORKLOADS["pyloric_ctx_l2pgd_rkl"])
        if args.scaling == "strong":
            w5["rows"] = (8 * w5["rows"]) // world
        rec = measure("pyloric_ctx_l2pgd_rkl", w5, args, 2, 1, dev, world, rank, local, args.scaling,
                      with_baselines=not args.no_cpu_baseline)
        line["also"] = {"pyloric_ctx_l2pgd_rkl": rec}
    if rank == 0:
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def state_cpu_with_zscore(state, w):
    s = dict(state)
    s["embedding_net.0.mean"] = torch.zeros(w["dx"])
    s["embedding_net.0.std"] = torch.ones(w["dx"])
    return s


if __name__ == "__main__":
    main()
