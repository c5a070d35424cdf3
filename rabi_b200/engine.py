"""Thin Python wrapper around one ``rabi_maf_t`` handle: device buffers, streams and error handling.

PyTorch is used here only for device memory and the current CUDA stream; all arithmetic happens in
``librabi_b200.so``.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import RabiError


def _ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), (t.device, t.dtype, t.is_contiguous())
    return ctypes.c_void_p(t.data_ptr())


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


class MafEngine:
    """Owns the packed weight image of one MAF (K masked-MADE conditioners) on one CUDA device."""

    def __init__(self, K: int, D: int, C: int, hidden: Sequence[int], log_scale_min: float, log_scale_max: float,
                 device: torch.device):
        self.L = _lib.lib()
        self.K, self.D, self.C, self.hidden = int(K), int(D), int(C), [int(h) for h in hidden]
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RabiError("rabi_b200 runs on CUDA devices only (no CPU fallback); move the model with .to('cuda')")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        self.h = ctypes.c_void_p()
        arr = (ctypes.c_int * len(self.hidden))(*self.hidden)
        rc = self.L.rabi_maf_create(ctypes.byref(self.h), idx, self.K, self.D, self.C, len(self.hidden), arr,
                                    float(log_scale_min), float(log_scale_max))
        if rc != 0:
            msg = self.L.rabi_last_error(None).decode()
            self.h = None
            raise ValueError(msg) if "Hidden dimension" in msg else RabiError(msg)
        self._versions = None
        self.epoch = 0   # bumped whenever a layer is re-packed: projections computed before are stale
        self._ctor = (K, D, C, list(hidden), log_scale_min, log_scale_max)
        self._structure = None
        self._params = None
        self._lane = None          # second handle + stream for attacks whose launches leave a ragged last round (pgd_run)
        self._row_scales = None
        self._col_mask = None

    def invalidate(self):
        self._versions = None
        if self._lane is not None:
            self._lane[0].invalidate()

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.rabi_maf_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _check(self, rc: int):
        if rc != 0:
            raise RabiError(self.L.rabi_last_error(self.h).decode())

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def launches(self) -> int:
        n = int(self.L.rabi_launch_count(self.h))
        return n + (self._lane[0].launches if self._lane is not None else 0)

    def profile(self, enable: bool):
        self._check(self.L.rabi_profile(self.h, int(enable)))

    def profile_read(self):
        """-> (summed duration in ms of the timed dominant-kernel launches, number of launches); synchronises."""
        ms, n = ctypes.c_double(), ctypes.c_ulonglong()
        self._check(self.L.rabi_profile_read(self.h, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, int(n.value)

    # ------------------------------------------------------------------ structure / weights
    def set_structure(self, perms: List[torch.Tensor], masks: List[List[torch.Tensor]],
                      post_perms: Optional[List[Optional[torch.Tensor]]] = None):
        for k in range(self.K):
            pp = None if post_perms is None else post_perms[k]
            if pp is not None:
                pp = pp.detach().to("cpu", torch.int64).contiguous()
                self._check(self.L.rabi_maf_set_post_permutation(
                    self.h, k, ctypes.cast(pp.data_ptr(), ctypes.POINTER(ctypes.c_int64))))
            else:
                self._check(self.L.rabi_maf_set_post_permutation(self.h, k, None))
            p = perms[k].detach().to("cpu", torch.int64).contiguous()
            self._check(self.L.rabi_maf_set_permutation(self.h, k, ctypes.cast(p.data_ptr(), ctypes.POINTER(ctypes.c_int64))))
            for l, m in enumerate(masks[k]):
                m = m.detach().to("cpu", torch.float32).contiguous()
                self._check(self.L.rabi_maf_set_mask(self.h, k, l, ctypes.cast(m.data_ptr(), ctypes.POINTER(ctypes.c_float)),
                                                     m.shape[0], m.shape[1]))
        self._check(self.L.rabi_maf_finalize_structure(self.h))
        self._structure = (perms, masks, post_perms)
        self._lane = None
        self._versions = None
        self._has_post = post_perms is not None and any(p is not None for p in post_perms)

    def set_weights(self, params: List[List[Tuple[torch.Tensor, torch.Tensor]]]):
        """params[k][layer] = (weight, bias) CUDA fp32 tensors; re-packs only what changed since the last call."""
        vers = [[(w.data_ptr(), w._version, b.data_ptr(), b._version) for (w, b) in pk] for pk in params]
        self._params = params
        if self._versions == vers:
            return
        self.epoch += 1
        with torch.cuda.device(self.device):
            for k, pk in enumerate(params):
                for l, (w, b) in enumerate(pk):
                    if self._versions is not None and self._versions[k][l] == vers[k][l]:
                        continue
                    w_, b_ = _f32c(w.detach()), _f32c(b.detach())
                    self._check(self.L.rabi_maf_set_weights(self.h, k, l, _ptr(w_), _ptr(b_), self._stream()))
        self._versions = vers

    # ------------------------------------------------------------------ kernels
    def ctx_project(self, x: torch.Tensor, zs_mean: Optional[torch.Tensor] = None,
                    zs_std: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [B, C] -> A [K, H0, B]."""
        x = _f32c(x)
        B = x.shape[0]
        A = torch.empty((self.K, self.hidden[0], B), device=self.device, dtype=torch.float32)
        if B:
            with torch.cuda.device(self.device):
                self._check(self.L.rabi_maf_ctx_project(self.h, _ptr(x), _ptr(zs_mean), _ptr(zs_std), B, _ptr(A),
                                                        self._stream()))
        return A

    def ctx_project_bwd(self, gA: torch.Tensor, zs_std: Optional[torch.Tensor] = None) -> torch.Tensor:
        B = gA.shape[-1]
        gx = torch.empty((B, self.C), device=self.device, dtype=torch.float32)
        if B:
            with torch.cuda.device(self.device):
                self._check(self.L.rabi_maf_ctx_project_bwd(self.h, _ptr(_f32c(gA)), _ptr(zs_std), B, _ptr(gx),
                                                            self._stream()))
        return gx

    def log_prob(self, A: torch.Tensor, theta: torch.Tensor) -> torch.Tensor:
        """theta [S, B, D] -> log q [S, B]."""
        theta = _f32c(theta)
        S, B, _ = theta.shape
        out = torch.empty((S, B), device=self.device, dtype=torch.float32)
        if S * B:
            with torch.cuda.device(self.device):
                self._check(self.L.rabi_maf_log_prob_fwd(self.h, _ptr(A), _ptr(theta), S, B, _ptr(out), self._stream()))
        return out

    def _grad_buffers(self):
        """[[(g_weight, g_bias) per layer] per transform] as views of ONE block laid out [W(k,l) | b(k,l)] in (k, l) order: the
        library then clears all of them with a single memset."""
        dims = [self.C + self.D] + list(self.hidden)
        outs = list(self.hidden) + [2 * self.D]
        total = self.K * sum(o * d + o for o, d in zip(outs, dims))
        flat = torch.empty((total,), device=self.device, dtype=torch.float32)
        grads, off = [], 0
        for _ in range(self.K):
            gk = []
            for o, d in zip(outs, dims):
                w = flat[off: off + o * d].view(o, d)
                off += o * d
                b = flat[off: off + o]
                off += o
                gk.append((w, b))
            grads.append(gk)
        return grads

    def fused_training_ok(self) -> bool:
        """rabi_maf_log_prob_bwd serves two-hidden-layer conditioners (the tensor-core path's shapes) without Permute layers."""
        import os
        if os.environ.get("RABI_FUSED_TRAIN", "1") == "0":   # A/B switch: op-by-op differentiable path everywhere
            return False
        return len(self.hidden) == 2 and self.D <= 12 and max(self.hidden) <= 104 and not getattr(self, "_has_post", False)

    def log_prob_bwd(self, ctx_rows: torch.Tensor, A: torch.Tensor, theta: torch.Tensor, gout: torch.Tensor,
                     want_weights: bool = True):
        """theta [S, B, D], ctx_rows [B, C] (after z-score / embedding), A [K, H0, B], gout [S, B] = dL/dlogq ->
        (g_theta [S, B, D], g_A [K, H0, B], [[(g_weight, g_bias) per layer] per transform] | None); reference layouts."""
        theta, gout, ctx_rows = _f32c(theta), _f32c(gout), _f32c(ctx_rows)
        S, B, D = theta.shape
        g_theta = torch.empty_like(theta)
        g_A = torch.empty((self.K, self.hidden[0], B), device=self.device, dtype=torch.float32)
        grads, wp, bp = None, None, None
        if want_weights:
            grads = self._grad_buffers()
            flat = [g for gk in grads for g in gk]
            wp = (ctypes.c_void_p * len(flat))(*[g[0].data_ptr() for g in flat])
            bp = (ctypes.c_void_p * len(flat))(*[g[1].data_ptr() for g in flat])
        with torch.cuda.device(self.device):
            self._check(self.L.rabi_maf_log_prob_bwd(self.h, _ptr(ctx_rows), _ptr(A), _ptr(theta), _ptr(gout), S, B, None,
                                                     _ptr(g_theta), _ptr(g_A), wp, bp, self._stream()))
        return g_theta, g_A, grads

    def fisher_trace(self, ctx_rows: torch.Tensor, A: torch.Tensor, z: torch.Tensor, zs_std: Optional[torch.Tensor] = None):
        """z [S, B, D] -> (trace [B], score [S, B, C], theta [S, B, D]): the MC score estimate of tr F(x) on the kernels."""
        z, ctx_rows = _f32c(z), _f32c(ctx_rows)
        S, B, D = z.shape
        trace = torch.empty((B,), device=self.device, dtype=torch.float32)
        score = torch.empty((S, B, self.C), device=self.device, dtype=torch.float32)
        theta = torch.empty_like(z)
        if S * B:
            with torch.cuda.device(self.device):
                self._check(self.L.rabi_fisher_trace_fwd(self.h, _ptr(ctx_rows), _ptr(A), _ptr(z), _ptr(zs_std), S, B, _ptr(trace),
                                                         _ptr(score), _ptr(theta), self._stream()))
        return trace, score, theta

    def fisher_trace_bwd(self, ctx_rows, A, z, zs_std, score, w_row, masked_weights, out_biases):
        """Weight gradient of sum_b w_row[b] tr F(x_b) right after ``fisher_trace`` (same ctx_rows / A / z): the closed form on the
        kernels (k_fim_dual + k_wgrad + the sampling-chain reverse).  masked_weights[k] = (W0*m0, W1*m1, Wo*mo) in the reference
        layout, out_biases[k] = output-layer bias.  -> [[(g_weight, g_bias) per layer] per transform]."""
        z, ctx_rows, score, w_row = _f32c(z), _f32c(ctx_rows), _f32c(score), _f32c(w_row)
        S, B, _ = z.shape
        grads = self._grad_buffers()
        flat = [g for gk in grads for g in gk]
        wp = (ctypes.c_void_p * len(flat))(*[g[0].data_ptr() for g in flat])
        bp = (ctypes.c_void_p * len(flat))(*[g[1].data_ptr() for g in flat])
        keep = [_f32c(w) for wk in masked_weights for w in wk] + [_f32c(b) for b in out_biases]
        n = 3 * self.K
        wm = (ctypes.c_void_p * n)(*[t.data_ptr() for t in keep[:n]])
        bo = (ctypes.c_void_p * self.K)(*[t.data_ptr() for t in keep[n:]])
        with torch.cuda.device(self.device):
            self._check(self.L.rabi_fisher_trace_bwd(self.h, _ptr(ctx_rows), _ptr(A), _ptr(z), _ptr(zs_std), S, B, _ptr(score),
                                                     _ptr(w_row), wm, bo, wp, bp, self._stream()))
        return grads

    def train_emissions(self, S: int, B: int) -> dict:
        """Per-pair vectors of the last training-side kernel call (same S, B), as [K, units, P] views of one copied buffer."""
        K, H0, H1, D, P = self.K, self.hidden[0], self.hidden[1], self.D, S * B
        buf = torch.empty((K * P * (2 * H0 + 2 * H1 + 3 * D),), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            self._check(self.L.rabi_train_emissions(self.h, S, B, _ptr(buf), self._stream()))
        out, off = {}, 0
        for name, n in (("h0", H0), ("a0", H0), ("h1", H1), ("a1", H1), ("o", 2 * D), ("y", D)):
            out[name] = buf[off: off + K * n * P].view(K, n, P)
            off += K * n * P
        return out

    def rsample_bwd(self, ctx_rows: torch.Tensor, A: torch.Tensor, z: torch.Tensor, g_theta: torch.Tensor,
                    g_logq: torch.Tensor, want_weights: bool = True):
        """Reverse of theta, log q = rsample(z): upstream g_theta [S, B, D], g_logq [S, B] ->
        (g_A [K, H0, B], [[(g_weight, g_bias) per layer] per transform] | None)."""
        z, g_theta, g_logq, ctx_rows = _f32c(z), _f32c(g_theta), _f32c(g_logq), _f32c(ctx_rows)
        S, B, D = z.shape
        g_A = torch.empty((self.K, self.hidden[0], B), device=self.device, dtype=torch.float32)
        grads, wp, bp = None, None, None
        if want_weights:
            grads = self._grad_buffers()
            flat = [g for gk in grads for g in gk]
            wp = (ctypes.c_void_p * len(flat))(*[g[0].data_ptr() for g in flat])
            bp = (ctypes.c_void_p * len(flat))(*[g[1].data_ptr() for g in flat])
        with torch.cuda.device(self.device):
            self._check(self.L.rabi_maf_rsample_bwd(self.h, _ptr(ctx_rows), _ptr(A), _ptr(z), _ptr(g_theta), _ptr(g_logq), S, B,
                                                    None, None, _ptr(g_A), wp, bp, self._stream()))
        return g_A, grads

    def rsample(self, A: torch.Tensor, z: torch.Tensor, want_logq: bool = True):
        z = _f32c(z)
        S, B, _ = z.shape
        theta = torch.empty_like(z)
        logq = torch.empty((S, B), device=self.device, dtype=torch.float32) if want_logq else None
        if S * B:
            with torch.cuda.device(self.device):
                self._check(self.L.rabi_maf_rsample_fwd(self.h, _ptr(A), _ptr(z), S, B, _ptr(theta), _ptr(logq),
                                                        self._stream()))
        return theta, logq

    def rkl(self, A_p: torch.Tensor, A_q: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
        z = _f32c(z)
        S, B, _ = z.shape
        out = torch.empty((B,), device=self.device, dtype=torch.float32)
        if S * B:
            with torch.cuda.device(self.device):
                self._check(self.L.rabi_rkl_fwd(self.h, _ptr(A_p), _ptr(A_q), _ptr(z), S, B, _ptr(out), self._stream()))
        return out

    def rkl_input_grad(self, x, zs_mean, zs_std, A_q, z, inv_B, want_kl=True, forward_kl=False):
        """-> (d[inv_B * sum_b KL(q(.|x_b) || target_b)]/dx  [B, C], kl [B] | None); ``forward_kl``: KL(target || q(.|x))."""
        z = _f32c(z)
        S, B, _ = z.shape
        gx = torch.empty((B, self.C), device=self.device, dtype=torch.float32)
        kl = torch.empty((B,), device=self.device, dtype=torch.float32) if want_kl else None
        with torch.cuda.device(self.device):
            fn = self.L.rabi_fkl_input_grad if forward_kl else self.L.rabi_rkl_input_grad
            self._check(fn(self.h, _ptr(_f32c(x)), _ptr(zs_mean), _ptr(zs_std), _ptr(A_q), _ptr(z),
                                                   S, B, float(inv_B), _ptr(gx), _ptr(kl), self._stream()))
        return gx, kl

    def set_row_scales(self, eps_rows=None, eps_iter_rows=None):
        """Per-row attack radius / step size for batched eps-sweeps ([B] fp32 device tensors the caller keeps alive until it
        clears them with ``set_row_scales()``); while set they replace the scalar eps / eps_iter of every PGD call."""
        if (eps_rows is None) != (eps_iter_rows is None):
            raise ValueError("pass both per-row arrays or neither")
        self._row_scales = (None if eps_rows is None else (_f32c(eps_rows), _f32c(eps_iter_rows)))
        a, b = self._row_scales if self._row_scales else (None, None)
        self._check(self.L.rabi_pgd_set_row_scales(self.h, _ptr(a), _ptr(b)))

    def set_column_mask(self, mask=None):
        """Masked-feature attacks: [C] fp32 device tensor (1 = attacked, 0 = frozen), kept alive here until cleared."""
        self._col_mask = None if mask is None else _f32c(mask)
        self._check(self.L.rabi_pgd_set_column_mask(self.h, _ptr(self._col_mask)))

    def pgd_update(self, x, delta, grad, eps, eps_iter, clip_min, clip_max, norm_floor=1e-6):
        """In-place L2 step / box clamp / projection of ``delta`` given d loss / d delta."""
        B, dx = x.shape
        with torch.cuda.device(self.device):
            self._check(self.L.rabi_pgd_update(self.h, _ptr(x), _ptr(delta), _ptr(_f32c(grad)), B, dx, float(eps),
                                               float(eps_iter), float(clip_min), float(clip_max), float(norm_floor),
                                               self._stream()))

    def pgd_step(self, x, delta, zs_mean, zs_std, A_q, z, eps, eps_iter, clip_min, clip_max, inv_B,
                 norm_floor=1e-6, want_kl=False):
        z = _f32c(z)
        S, B, _ = z.shape
        kl = torch.empty((B,), device=self.device, dtype=torch.float32) if want_kl else None
        with torch.cuda.device(self.device):
            self._check(self.L.rabi_rkl_pgd_step(self.h, _ptr(x), _ptr(delta), _ptr(zs_mean), _ptr(zs_std), _ptr(A_q),
                                                 _ptr(z), S, B, float(eps), float(eps_iter), float(clip_min),
                                                 float(clip_max), float(inv_B), float(norm_floor), _ptr(kl),
                                                 self._stream()))
        return kl

    # ---- two lanes: the three-hidden-layer path runs ONE tile of 128 pairs per SM and one launch per PGD iteration, so an
    # attack of T tiles costs ceil(T / SMs) tile times per iteration (pyloric, 12 500 rows x 5: 489 tiles = 3.3 -> 4 rounds).
    # Rows are independent: two halves of the rows, each on its own handle (own scratch) and stream, fill each other's ragged
    # rounds -- iteration i + 1 of one half starts on the SMs the other half's last round leaves idle.
    def _lanes(self, S: int, B: int) -> int:
        import os
        want = int(os.environ.get("RABI_PGD_LANES", "0"))   # 0: decide here
        if want:
            return max(1, min(2, want))
        if len(self.hidden) != 3 or B < 8 or torch.cuda.is_current_stream_capturing():
            return 1
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        tiles = (S * B + 127) // 128
        rounds = -(-tiles // sms)
        return 2 if (rounds >= 2 and tiles <= 0.9 * rounds * sms) else 1

    def _lane_engine(self):
        if self._lane is None:
            eng = MafEngine(*self._ctor, self.device)
            eng.set_structure(*self._structure)
            self._lane = (eng, torch.cuda.Stream(device=self.device))
        eng, stream = self._lane
        eng.set_weights(self._params)
        return eng, stream

    def pgd_run(self, x, delta, zs_mean, zs_std, A_q, z_all, eps, eps_iter, clip_min, clip_max, inv_B,
                norm_floor=1e-6, forward_kl=False):
        """z_all [nb_iter, S, B, D]; updates delta in place and returns x_adv = clamp(x + delta).
        A_q None: untargeted (projection of the clean x is computed inside)."""
        z_all = _f32c(z_all)
        nb_iter, S, B, _ = z_all.shape
        x_adv = torch.empty_like(x)
        if nb_iter > 0 and x.is_contiguous() and delta.is_contiguous() and self._structure is not None \
                and self._params is not None and self._lanes(S, B) == 2:
            other, side = self._lane_engine()
            B1 = ((B + 1) // 2 + 3) & ~3   # (row offsets stay 16-byte aligned for any feature count)
            main = torch.cuda.current_stream(self.device)
            # the halves of every [.., B, ..] operand as contiguous copies, made on the calling stream before the fork
            parts = []
            for lo, hi in ((0, B1), (B1, B)):
                parts.append(dict(
                    z=z_all[:, :, lo:hi].contiguous(),
                    A=None if A_q is None else A_q[:, :, lo:hi].contiguous(),
                    rows=None if self._row_scales is None else tuple(t[lo:hi].contiguous() for t in self._row_scales)))
            other.set_column_mask(self._col_mask)
            other.set_row_scales(*(parts[1]["rows"] or (None, None)))
            saved_rows = self._row_scales
            if saved_rows is not None:
                self.set_row_scales(*parts[0]["rows"])
            side.wait_stream(main)
            try:
                with torch.cuda.stream(side):
                    other._pgd_run_into(x[B1:], delta[B1:], zs_mean, zs_std, parts[1]["A"], parts[1]["z"], eps, eps_iter, clip_min,
                                        clip_max, inv_B, norm_floor, forward_kl, x_adv[B1:])
                self._pgd_run_into(x[:B1], delta[:B1], zs_mean, zs_std, parts[0]["A"], parts[0]["z"], eps, eps_iter, clip_min,
                                   clip_max, inv_B, norm_floor, forward_kl, x_adv[:B1])
            finally:
                main.wait_stream(side)
                other.set_row_scales()
                other.set_column_mask()
                if saved_rows is not None:
                    self.set_row_scales(*saved_rows)
            return x_adv
        self._pgd_run_into(x, delta, zs_mean, zs_std, A_q, z_all, eps, eps_iter, clip_min, clip_max, inv_B, norm_floor,
                           forward_kl, x_adv)
        return x_adv

    def _pgd_run_into(self, x, delta, zs_mean, zs_std, A_q, z_all, eps, eps_iter, clip_min, clip_max, inv_B, norm_floor,
                      forward_kl, x_adv):
        nb_iter, S, B, _ = z_all.shape
        with torch.cuda.device(self.device):
            fn = self.L.rabi_fkl_pgd_run if forward_kl else self.L.rabi_rkl_pgd_run
            self._check(fn(self.h, _ptr(x), _ptr(delta), _ptr(zs_mean), _ptr(zs_std), _ptr(A_q),
                                                _ptr(z_all), nb_iter, S, B, float(eps), float(eps_iter), float(clip_min),
                                                float(clip_max), float(inv_B), float(norm_floor), _ptr(x_adv),
                                                self._stream()))
