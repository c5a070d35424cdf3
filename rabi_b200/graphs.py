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
        # every call keeps its own copy of the inputs: its backward re-stages them, so several forwards of one signature
        # may precede their backwards (gradient accumulation, several pre-loss regularisers)
        ctx.inputs = [t.detach().clone() for t in inputs]
        for dst, src in zip(st["in"], ctx.inputs):
            dst.copy_(src)
        st["fwd"].replay()
        return st["out"].clone()

    @staticmethod
    def backward(ctx, gout):
        st = ctx.st
        for dst, src in zip(st["in"], ctx.inputs):
            dst.copy_(src)
        st["gout"].copy_(gout)
        st["bwd"].replay()
        # clones: autograd may adopt the returned tensors as .grad, and the static buffers are rewritten next step
        return (None, None) + (None,) * len(ctx.inputs) + tuple(g.clone() for g in st["grads"])


class GraphedParamFunction:
    def __init__(self, module: torch.nn.Module, fn: Callable[..., Tensor], warmup: int = 2):
        self.module = module
        self.fn = fn
        self.warmup = warmup
        self._seen: Dict[tuple, int] = {}
        self._graphs: Dict[tuple, dict] = {}
        self._failed: set = set()

    def _key(self, inputs: Sequence[Tensor]):
        return (tuple((tuple(t.shape), t.dtype, t.device) for t in inputs),
                tuple(id(p) for p in self.module.parameters()))

    def ready(self, inputs: Sequence[Tensor]) -> bool:
        """True once this input signature has been seen ``warmup`` times (the eager calls warm every lazy init up)."""
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
            with torch.cuda.graph(fwd, stream=side):
                # under enable_grad on purpose: the differentiable path (rabi_sgemm + glue) has no host-side state,
                # whereas the inference path re-packs the weight image only when the host sees a new parameter version
                with torch.enable_grad():
                    st["out"] = self.fn(*st["in"]).detach()
            st["gout"] = torch.zeros_like(st["out"])
            bwd = torch.cuda.CUDAGraph()
            with torch.cuda.graph(bwd, stream=side):
                leaves = {n: p.detach().requires_grad_() for n, p in self.module.named_parameters()}
                with stateless._reparametrize_module(self.module, leaves), torch.enable_grad():
                    out = self.fn(*st["in"])
                grads = torch.autograd.grad(out, [leaves[n] for n in names], grad_outputs=st["gout"], allow_unused=True)
                st["grads"] = [g if g is not None else torch.zeros_like(leaves[n]) for g, n in zip(grads, names)]
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
        return _Replay.apply(st, len(inputs), *inputs, *self.module.parameters())
