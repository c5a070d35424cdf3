"""CUDA-graph replay of the training-side segments.

The training-side math (rabi_b200/autograd.py) is a few hundred small launches per step -- `rabi_sgemm` calls with
elementwise glue -- i.e. launch-bound at batch 512.  ``GraphedParamFunction`` captures such a segment ONCE per input
shape into two CUDA graphs,

    forward :  out  = f(*inputs)                               (no autograd)
    backward:  grads = d <out, gout> / d parameters            (forward recomputed inside the graph)

and exposes it as one autograd node whose inputs are the module's parameters: the caller's ``loss.backward()`` replays
the backward graph and hands the parameter gradients to autograd, so the reference's training loop
(``loss = loss_fn(x, theta); loss.backward(); optimizer.step()``, rbibm/rbibm/runs/run_train.py) runs unchanged.
The graphs differentiate fresh leaves that alias the live parameter storage, so in-place optimizer updates are seen.
Only parameter gradients are produced (inputs are data); callers fall back to the eager path when an input requires
grad, when base noise is injected (parity tests) or during the first ``warmup`` calls of a shape.
"""
from __future__ import annotations

from typing import Callable, Dict, Sequence

import torch
from torch import Tensor
from torch.nn.utils import stateless


class _Replay(torch.autograd.Function):
    @staticmethod
    def forward(ctx, st, n_in, *args):
        inputs = args[:n_in]
        ctx.st = st
        # the call remembers its inputs (by reference + version counter) and the staging generation: when no other forward of
        # this signature ran in between, its backward finds them still staged; otherwise it stages them again, so several
        # forwards may precede their backwards (gradient accumulation, several pre-loss regularisers)
        ctx.inputs = [t.detach() for t in inputs]
        ctx.versions = [t._version for t in inputs]
        for dst, src in zip(st["in"], ctx.inputs):
            dst.copy_(src)
        st["gen"] += 1
        ctx.gen = st["gen"]
        st["fwd"].replay()
        return st["out"].clone()

    @staticmethod
    def backward(ctx, gout):
        st = ctx.st
        if st["gen"] != ctx.gen:
            for t, v in zip(ctx.inputs, ctx.versions):
                if t._version != v:
                    raise RuntimeError("rabi_b200: an input of a graphed training segment was modified in place before its backward")
            for dst, src in zip(st["in"], ctx.inputs):
                dst.copy_(src)
            st["gen"] += 1   # (the staged inputs now belong to nobody's forward)
        st["gout"].copy_(gout)
        st["bwd"].replay()
        # ONE copy of the flat gradient block: autograd may adopt the returned tensors as .grad, and the static buffer is
        # rewritten by the next replay; the per-parameter gradients are views of the copy
        flat = st["flat"].clone()
        return (None, None) + (None,) * len(ctx.inputs) + tuple(
            flat[o: o + n].view(shape) for o, n, shape in st["slices"])


class GraphedParamFunction:
    def __init__(self, module: torch.nn.Module, fn: Callable[..., Tensor], warmup: int = 2):
        self.module = module
        self.fn = fn
        self.warmup = warmup
        self._seen: Dict[tuple, int] = {}
        self._graphs: Dict[tuple, dict] = {}
        self._failed: set = set()
        self._last = None   # (input objects, key, parameter list) of the latest ready(): __call__ follows with the same inputs

    def _key(self, inputs: Sequence[Tensor]):
        last = self._last
        if last is not None and len(last[0]) == len(inputs) and all(a is b for a, b in zip(last[0], inputs)):
            return last[1]
        params = list(self.module.parameters())   # one walk of the module tree per step
        key = (tuple((tuple(t.shape), t.dtype, t.device) for t in inputs), tuple(id(p) for p in params))
        self._last = (tuple(inputs), key, params)
        return key

    def ready(self, inputs: Sequence[Tensor]) -> bool:
        """True once this input signature has been seen ``warmup`` times (the eager calls warm every lazy init up)."""
        self._last = None
        key = self._key(inputs)
        if key in self._failed:
            return False
        n = self._seen.get(key, 0)
        self._seen[key] = n + 1
        return n >= self.warmup

    def _capture(self, key, inputs: Sequence[Tensor]) -> dict:
        dev = inputs[0].device
        names = [n for n, _ in self.module.named_parameters()]
        st = {"in": [t.detach().clone() for t in inputs]}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            fwd = torch.cuda.CUDAGraph()
            # The device weight image is re-packed when the HOST sees new parameter versions -- a check no replay runs.  Each
            # graph therefore captures its own re-pack (module.invalidate() forces one inside the capture): replays read the
            # live parameter storage whatever happened to the weights since the capture.
            invalidate = getattr(self.module, "invalidate", lambda: None)
            with torch.cuda.graph(fwd, stream=side):
                invalidate()
                # under enable_grad on purpose: the differentiable path (rabi_sgemm + glue) has no host-side state,
                # whereas the inference path re-packs the weight image only when the host sees a new parameter version
                with torch.enable_grad():
                    st["out"] = self.fn(*st["in"]).detach()
            st["gout"] = torch.zeros_like(st["out"])
            bwd = torch.cuda.CUDAGraph()
            with torch.cuda.graph(bwd, stream=side):
                leaves = {n: p.detach().requires_grad_() for n, p in self.module.named_parameters()}
                invalidate()
                with stateless._reparametrize_module(self.module, leaves), torch.enable_grad():
                    out = self.fn(*st["in"])
                grads = torch.autograd.grad(out, [leaves[n] for n in names], grad_outputs=st["gout"], allow_unused=True)
                grads = [g if g is not None else torch.zeros_like(leaves[n]) for g, n in zip(grads, names)]
                st["flat"] = torch.cat([g.reshape(-1) for g in grads])   # (one launch inside the graph)
        st["slices"], off = [], 0
        for g_ in grads:
            st["slices"].append((off, g_.numel(), tuple(g_.shape)))
            off += g_.numel()
        st["gen"] = 0
        torch.cuda.current_stream(dev).wait_stream(side)
        st["fwd"], st["bwd"] = fwd, bwd
        self._graphs[key] = st
        return st

    def __call__(self, *inputs: Tensor) -> Tensor:
        key = self._key(inputs)
        st = self._graphs.get(key)
        if st is None:
            try:
                st = self._capture(key, inputs)
            except RuntimeError as e:
                # a capture can be refused for reasons outside this module (e.g. the driver loading a kernel image lazily on its
                # first launch, which is not permitted while a stream captures): this signature then stays on the eager path
                import warnings

                warnings.warn(f"rabi_b200: CUDA-graph capture failed ({str(e).splitlines()[0][:120]}); running this segment eagerly")
                self._failed.add(key)
                torch.cuda.synchronize()
                return self.fn(*inputs)
        params = self._last[2]
        self._last = None   # (do not keep the batch alive)
        return _Replay.apply(st, len(inputs), *inputs, *params)
