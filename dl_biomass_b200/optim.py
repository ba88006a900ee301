"""Flat parameter arena + fused Adam: the optimiser of the reference's training loop on one launch.

/root/reference/main.py:84 builds ``torch.optim.Adam(model.parameters(), lr, weight_decay)`` and :172 steps it: a
multi-tensor launch list over ~40 small tensors plus a host-side step counter.  Here every trainable parameter of the
model lives in ONE flat fp32 buffer (``ParamArena``), grouped into buckets in backward order (head + SA3, SA2, SA1 --
the order in which backward finishes them, so a bucket can be all-reduced while the next one is still being
computed, see parallel.py); gradients, ``exp_avg`` and ``exp_avg_sq`` are parallel buffers.  The backward kernels of
libb2pn write their weight gradients straight into the gradient buffer (``param._b2pn_grad``), the data-parallel
all-reduce runs over slices of it, and ``FlatAdam.step()`` is a single ``b2pn_adam_step`` launch (csrc/optim.cu) whose
step counter is a device scalar -- the whole training step, optimiser included, replays from one CUDA graph.
There is no CPU path.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

DEFAULT_BUCKETS = (("mlp.", "sa3_module."), ("sa2_module.",), ("sa1_module.",))
ALIGN = 64  # floats: every parameter (and bucket) starts on a 256-byte boundary


def _round_up(x: int, a: int) -> int:
    return (x + a - 1) // a * a


class ParamArena:
    """All trainable parameters of ``module`` re-homed into one flat buffer (``p.data`` becomes a view of it).

    Build it AFTER the module is on its device and in its final dtype (``module.to(...)`` afterwards would detach the
    parameters from the arena).  ``load_state_dict`` / in-place updates keep working: they write through the views."""

    def __init__(self, module: torch.nn.Module, buckets: Sequence[Sequence[str]] = DEFAULT_BUCKETS):
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        if not named:
            raise ValueError("module has no trainable parameters")
        dev, dt = named[0][1].device, named[0][1].dtype
        if dt != torch.float32 or any(p.device != dev or p.dtype != dt for _, p in named):
            raise ValueError("ParamArena needs all parameters in float32 on one device")
        assign: List[List[Tuple[str, torch.nn.Parameter]]] = [[] for _ in buckets]
        rest: List[Tuple[str, torch.nn.Parameter]] = []
        for n, p in named:
            for i, prefixes in enumerate(buckets):
                if any(n.startswith(pre) for pre in prefixes):
                    assign[i].append((n, p))
                    break
            else:
                rest.append((n, p))
        if rest:
            assign.append(rest)
        groups = [g for g in assign if g]
        self.module = module
        self.device = dev
        self.names: List[str] = []
        self.params: List[torch.nn.Parameter] = []
        self.offsets: List[int] = []
        self.bucket_ranges: List[Tuple[int, int]] = []
        self.bucket_params: List[List[torch.nn.Parameter]] = []
        off = 0
        for g in groups:
            start = off
            for n, p in g:
                self.names.append(n)
                self.params.append(p)
                self.offsets.append(off)
                off = _round_up(off + p.numel(), ALIGN)
            self.bucket_ranges.append((start, off))
            self.bucket_params.append([p for _, p in g])
        self.numel = off
        self.flat_params = torch.zeros(off, dtype=dt, device=dev)
        self.flat_grads = torch.zeros(off, dtype=dt, device=dev)
        self.param_views: Dict[torch.nn.Parameter, torch.Tensor] = {}
        self.grad_views: Dict[torch.nn.Parameter, torch.Tensor] = {}
        self.bucket_of: Dict[torch.nn.Parameter, int] = {}
        with torch.no_grad():
            for bi, ps in enumerate(self.bucket_params):
                for p in ps:
                    self.bucket_of[p] = bi
            for p, o in zip(self.params, self.offsets):
                pv = self.flat_params[o:o + p.numel()].view_as(p)
                pv.copy_(p.data)
                p.data = pv
                gv = self.flat_grads[o:o + p.numel()].view_as(p)
                self.param_views[p] = pv
                self.grad_views[p] = gv
                # where libb2pn's backward kernels put this parameter's gradient (sa.py / head.py look it up)
                p._b2pn_grad = gv
                p.grad = None
        module._b2pn_arena = self

    @staticmethod
    def of(module: torch.nn.Module) -> Optional["ParamArena"]:
        return getattr(module, "_b2pn_arena", None)

    def bucket_grads(self, i: int) -> torch.Tensor:
        s, e = self.bucket_ranges[i]
        return self.flat_grads[s:e]

    def bucket_flat_params(self, i: int) -> torch.Tensor:
        s, e = self.bucket_ranges[i]
        return self.flat_params[s:e]

    def intact(self) -> bool:
        """True while every parameter still lives in the arena (``module.to()`` / ``p.data = ...`` would break that)."""
        return all(p.data_ptr() == v.data_ptr() for p, v in self.param_views.items())

    def fold(self, p: torch.nn.Parameter) -> None:
        """Make the arena hold ``p``'s gradient of this step.  Free when the backward kernel wrote it there (the fused
        paths do); a gradient autograd produced elsewhere (ATen ops) is copied in; a missing one counts as zero."""
        gv = self.grad_views[p]
        g = p.grad
        if g is None:
            gv.zero_()
        elif g.data_ptr() != gv.data_ptr():
            gv.copy_(g)
            p.grad = gv

    def collect(self) -> None:
        for p in self.params:
            self.fold(p)

    def release_grads(self) -> None:
        """``zero_grad(set_to_none=True)``: the next backward overwrites the arena, nothing needs clearing."""
        for p in self.params:
            p.grad = None


class FlatAdam:
    """``torch.optim.Adam(params, lr, betas, eps, weight_decay)`` (L2-in-gradient form, no amsgrad) of
    /root/reference/main.py:84 over a ``ParamArena``: one ``b2pn_adam_step`` launch per ``step()``.

    ``param_groups`` mirrors torch's list-of-dicts far enough for schedulers that only touch ``lr`` (one group).
    ``grad_scale`` multiplies the gradient first (1/world_size after a SUM all-reduce; 1 after an AVG one)."""

    def __init__(self, module_or_arena, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, buckets: Sequence[Sequence[str]] = DEFAULT_BUCKETS):
        if isinstance(module_or_arena, ParamArena):
            arena = module_or_arena
        else:
            arena = ParamArena.of(module_or_arena) or ParamArena(module_or_arena, buckets)
        if arena.device.type != "cuda":
            raise RuntimeError("FlatAdam runs on a B200 only: there is no CPU fallback")
        self.arena = arena
        self.param_groups = [{"params": arena.params, "lr": float(lr), "betas": (float(betas[0]), float(betas[1])),
                              "eps": float(eps), "weight_decay": float(weight_decay)}]
        self.exp_avg = torch.zeros_like(arena.flat_params)
        self.exp_avg_sq = torch.zeros_like(arena.flat_params)
        self.state_dev = torch.zeros(2, dtype=torch.int64, device=arena.device)  # [steps taken, ticket]
        self.grad_scale = 1.0

    # -- torch.optim.Optimizer surface the training loop uses ----------------------------------------------
    def zero_grad(self, set_to_none: bool = True) -> None:
        if set_to_none:
            self.arena.release_grads()
        else:
            self.arena.flat_grads.zero_()
            for p in self.arena.params:
                p.grad = self.arena.grad_views[p]

    def step(self, collected: bool = False) -> None:
        from . import _lib
        a = self.arena
        if not collected:
            a.collect()
        g = self.param_groups[0]
        with torch.cuda.device(a.device):
            rc = _lib.lib().b2pn_adam_step(a.flat_params.data_ptr(), a.flat_grads.data_ptr(), self.exp_avg.data_ptr(),
                                           self.exp_avg_sq.data_ptr(), a.numel, g["lr"], g["betas"][0], g["betas"][1],
                                           g["eps"], g["weight_decay"], float(self.grad_scale),
                                           self.state_dev.data_ptr(), torch.cuda.current_stream(a.device).cuda_stream)
        _lib.check(rc, "b2pn_adam_step")

    @property
    def steps_taken(self) -> int:
        return int(self.state_dev[0].item())

    def state_dict(self) -> dict:
        a = self.arena
        per = {}
        for n, p, o in zip(a.names, a.params, a.offsets):
            per[n] = {"exp_avg": self.exp_avg[o:o + p.numel()].view_as(p).clone(),
                      "exp_avg_sq": self.exp_avg_sq[o:o + p.numel()].view_as(p).clone()}
        g = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        return {"step": self.steps_taken, "state": per, "param_group": g}

    def load_state_dict(self, sd: dict) -> None:
        a = self.arena
        with torch.no_grad():
            for n, p, o in zip(a.names, a.params, a.offsets):
                st = sd["state"][n]
                self.exp_avg[o:o + p.numel()].view_as(p).copy_(st["exp_avg"])
                self.exp_avg_sq[o:o + p.numel()].view_as(p).copy_(st["exp_avg_sq"])
            self.state_dev.zero_()
            self.state_dev[0] = int(sd["step"])
        self.param_groups[0].update(sd.get("param_group", {}))

    # snapshot / restore used by the graph-capturing steppers so that warm-up iterations leave no trace
    def _snapshot(self):
        return self.exp_avg.clone(), self.exp_avg_sq.clone(), self.state_dev.clone()

    def _restore(self, snap) -> None:
        self.exp_avg.copy_(snap[0])
        self.exp_avg_sq.copy_(snap[1])
        self.state_dev.copy_(snap[2])


def grad_dst(p: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """Arena slice the backward kernels should write ``p``'s gradient into, or None (allocate a fresh tensor).  The
    autograd Functions keep this (long-lived) view on their ctx and hand ``fresh_alias(dst)`` back to autograd."""
    d = getattr(p, "_b2pn_grad", None) if p is not None else None
    if d is None or d.device != p.device or d.shape != p.shape:
        return None
    return d


def fresh_alias(dst: torch.Tensor) -> torch.Tensor:
    """A NEW tensor object over ``dst``'s memory, to be created inside ``backward`` and returned from it: autograd's
    AccumulateGrad takes a returned gradient over as ``p.grad`` without copying only when nobody else references that
    tensor object -- an alias that was created in forward and kept on the ctx gets CLONED (one device-to-device memcpy
    per parameter and step, plus the copy back into the arena: 80 memcpy nodes in the step graph, measured round 2)."""
    return dst.view(dst.shape)


def iter_trainable(params: Iterable[torch.nn.Parameter]) -> List[torch.nn.Parameter]:
    return [p for p in params if p.requires_grad]
