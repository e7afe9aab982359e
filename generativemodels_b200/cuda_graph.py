"""CUDA-graph replay of a network forward for the launch-bound configurations (2-D latent UNets: ~200 kernels of a
few microseconds each per step, where the reference runs at 60-160 it/s regardless of model size — BASELINE.md §1).

    unet = graphed(unet)                       # same call signature; captures once per input-shape signature
    inferer.sample(input_noise=z, diffusion_model=unet, scheduler=scheduler)

The wrapped module is captured with ``torch.cuda.graph`` on its first call for a given (shapes, dtypes, optional-args)
signature: inputs are copied into static buffers, the graph is replayed, and a copy of the static output tensor is returned.  A captured graph bakes in the
addresses of the packed weights, so the cache is dropped whenever a parameter or buffer of the module is replaced or
modified in place (``load_state_dict``, ``.to()``, an optimiser step): the next call re-captures.  Everything the forward does
is capture-safe by construction: allocations come from the graph's private pool, kernels are enqueued on the current
(capturing) stream through the C-ABI, tensor maps are encoded on the host with fixed addresses.
"""
from __future__ import annotations

import contextlib
import gc

import torch
import torch.nn as nn


@contextlib.contextmanager
def capture(graph: "torch.cuda.CUDAGraph"):
    """``torch.cuda.graph(graph)`` with the cyclic garbage collector held off for the duration of the capture.

    A collection that happens to run in the middle of a capture can finalise an OLDER CUDAGraph (e.g. the decode-step
    graph of a transformer cache that has just gone out of scope); destroying it releases its private memory pool
    (cudaFree), which is not allowed while a stream is capturing and invalidates the capture in progress
    ("operation failed due to a previous error during capture" — seen intermittently in the round-2 GPU suite).
    ``torch.cuda.graph`` collects on entry; this additionally keeps the collector from starting on its own until the
    capture has ended.  Reference-count frees of tensors are unaffected (they only return blocks to the allocator)."""
    gc.collect()
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        with torch.cuda.graph(graph):
            yield
    finally:
        if was_enabled:
            gc.enable()


class GraphedModule(nn.Module):
    def __init__(self, module: nn.Module, warmup: int = 2) -> None:
        super().__init__()
        self.module = module
        self._warmup = warmup
        self._entries: dict = {}
        self._weights_sig = None
        self._tensors = None
        inner_apply = module._apply

        def _apply_and_invalidate(fn, *a, **k):       # .to() / .cuda() / .half() replace the parameter tensors
            self._tensors = None
            return inner_apply(fn, *a, **k)
        module._apply = _apply_and_invalidate

    def _current_weights_sig(self):
        # walking the module tree costs ~0.6 ms for a UNet, the cached tensor list ~45 us; in-place updates
        # (load_state_dict, optimiser steps) bump ``_version``, wholesale replacement goes through ``_apply`` above
        if self._tensors is None:
            self._tensors = list(self.module.parameters()) + list(self.module.buffers())
        return tuple((t.data_ptr(), t._version) for t in self._tensors)

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("module"), name)

    @staticmethod
    def _sig(args, kwargs):
        def one(v):
            if torch.is_tensor(v):
                return ("T", tuple(v.shape), v.dtype, v.device)
            if v is None or isinstance(v, (bool, int, float, str)):
                return ("C", v)                       # baked into the captured graph: part of the signature
            if isinstance(v, (list, tuple)):
                return ("L", type(v).__name__, tuple(one(e) for e in v))
            raise TypeError("graphed modules take tensors, None, scalars and (nested) lists / tuples of them")
        return tuple(one(a) for a in args), tuple((k, one(v)) for k, v in sorted(kwargs.items()))

    @staticmethod
    def _clone(v):
        if torch.is_tensor(v):
            return v.clone()
        if isinstance(v, (list, tuple)):
            return type(v)(GraphedModule._clone(e) for e in v)
        return v

    @staticmethod
    def _copy_into(dst, src):
        if torch.is_tensor(dst):
            dst.copy_(src, non_blocking=True)
        elif isinstance(dst, (list, tuple)):
            for d, s_ in zip(dst, src):
                GraphedModule._copy_into(d, s_)

    @torch.no_grad()
    def forward(self, *args, **kwargs):
        try:
            key = self._sig(args, kwargs)
        except TypeError:
            return self.module(*args, **kwargs)          # e.g. channels-last residual handles: run eagerly
        sig = self._current_weights_sig()
        if sig != self._weights_sig:            # weights changed since the graphs were captured: they are stale
            self._entries.clear()
            self._weights_sig = sig
        entry = self._entries.get(key)
        if entry is None:
            static_args = [self._clone(a) for a in args]
            static_kwargs = {k: self._clone(v) for k, v in kwargs.items()}
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(self._warmup):           # weight packing, attribute setup, allocator warm-up
                    self.module(*static_args, **static_kwargs)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with capture(graph):
                out = self.module(*static_args, **static_kwargs)
            entry = self._entries[key] = (graph, static_args, static_kwargs, out)
        graph, static_args, static_kwargs, out = entry
        for dst, src in zip(static_args, args):
            self._copy_into(dst, src)
        for k, dst in static_kwargs.items():
            self._copy_into(dst, kwargs[k])
        graph.replay()
        # fresh tensors per call (small stream-ordered copies): callers such as PNDMScheduler keep references to past
        # outputs, which shared static buffers would silently overwrite on the next replay
        return self._clone(out)


def graphed(module: nn.Module, warmup: int = 2) -> GraphedModule:
    """Wrap ``module`` so that each distinct input signature is captured into a CUDA graph on first use."""
    return GraphedModule(module, warmup)
