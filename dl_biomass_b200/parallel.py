"""Data-parallel training: one process per GPU, bucketed gradient all-reduce overlapped with backward.

Replaces ``torch_geometric.nn.DataParallel`` of /root/reference/main.py:140 (single process, one Python
thread per GPU, full parameter broadcast + gradient reduce-to-root every step).
Here every rank owns a replica and its own shard of tree clouds; the only collective per step is an
all-reduce (average) of the gradients, issued bucket by bucket on a side stream as soon as backward has
produced a bucket (head + SA3 first, then SA2, then SA1 = backward order), so NCCL runs over NVLink while
the remaining backward kernels execute.  BatchNorm statistics stay per rank, as under DataParallel.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

DEFAULT_BUCKETS = (("mlp.", "sa3_module."), ("sa2_module.",), ("sa1_module.",))


def shard_clouds(num_clouds: int, rank: int, world_size: int) -> range:
    """Contiguous shard of cloud indices for ``rank`` (clouds are the independent units of the path)."""
    per = (num_clouds + world_size - 1) // world_size
    return range(min(rank * per, num_clouds), min((rank + 1) * per, num_clouds))


class GradReducer:
    """Flat gradient buckets + async all-reduce.  Backend-agnostic: NCCL on GPUs, gloo in the CPU tests."""

    def __init__(self, module: torch.nn.Module, buckets: Sequence[Sequence[str]] = DEFAULT_BUCKETS,
                 process_group=None, broadcast_params: bool = True, overlap: bool = True):
        if not dist.is_initialized():
            raise RuntimeError("GradReducer needs torch.distributed to be initialised")
        self.group = process_group
        # overlap=True: a bucket is reduced as soon as backward has filled it (NCCL runs beside the remaining
        # backward kernels).  overlap=False: all buckets are reduced in finish(), after backward -- used when another
        # stream already shares the GPU with backward (train.PipelinedTrainStep): NCCL's CTAs would otherwise queue
        # behind the persistent kernels and the sampling kernels and stall both ranks.
        self.overlap = overlap
        # inline=True (only with overlap=False): the all-reduce is issued on the CURRENT stream, no side stream and no
        # events -- the form that can be captured into the step's CUDA graph.
        self.inline = False
        self.world_size = dist.get_world_size(process_group)
        self.module = module
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        assign: List[List[torch.nn.Parameter]] = [[] for _ in buckets]
        rest: List[torch.nn.Parameter] = []
        for n, p in named:
            for i, prefixes in enumerate(buckets):
                if any(n.startswith(pre) for pre in prefixes):
                    assign[i].append(p)
                    break
            else:
                rest.append(p)
        if rest:
            assign.append(rest)
        self.bucket_params = [b for b in assign if b]
        dev = named[0][1].device
        self.device = dev
        self.flat: List[torch.Tensor] = []
        self.views: Dict[torch.nn.Parameter, torch.Tensor] = {}
        self.bucket_of: Dict[torch.nn.Parameter, int] = {}
        for i, ps in enumerate(self.bucket_params):
            flat = torch.zeros(sum(p.numel() for p in ps), dtype=ps[0].dtype, device=dev)
            off = 0
            for p in ps:
                self.views[p] = flat[off:off + p.numel()].view_as(p)
                self.bucket_of[p] = i
                off += p.numel()
            self.flat.append(flat)
        self.pending = [0] * len(self.flat)
        self.works: List[Optional[object]] = [None] * len(self.flat)
        self.comm_stream = torch.cuda.Stream(dev) if dev.type == "cuda" else None
        self.done_events: List[Optional[torch.cuda.Event]] = [None] * len(self.flat)
        for p in self.views:
            p.register_post_accumulate_grad_hook(self._hook)
        if broadcast_params:  # replicas start identical (DataParallel re-broadcast every step; once is enough)
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, src=0, group=process_group)

    def prepare(self) -> None:
        """Point every .grad at its slice of a zeroed flat bucket so backward accumulates in place."""
        for i, flat in enumerate(self.flat):
            flat.zero_()
            self.pending[i] = len(self.bucket_params[i])
            self.works[i] = None
        for p, v in self.views.items():
            p.grad = v

    def _hook(self, p: torch.nn.Parameter) -> None:
        i = self.bucket_of[p]
        if p.grad is not self.views[p]:  # autograd replaced the tensor: fold it back into the bucket
            self.views[p].copy_(p.grad)
            p.grad = self.views[p]
        self.pending[i] -= 1
        if self.pending[i] == 0 and self.overlap:
            self._launch(i)

    def _launch(self, i: int) -> None:
        flat = self.flat[i]
        if self.inline and self.comm_stream is not None:
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
            self.works[i] = True
            self.done_events[i] = None
            return
        if self.comm_stream is not None:
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ready)
                dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
                ev = torch.cuda.Event()
                ev.record(self.comm_stream)
                self.done_events[i] = ev
            self.works[i] = True
        else:
            self.works[i] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self) -> None:
        """Block the compute stream (not the host) until every bucket is reduced."""
        for i in range(len(self.flat)):
            if self.works[i] is None:  # a bucket none of whose parameters received a gradient
                self._launch(i)
            if self.comm_stream is not None:
                if self.done_events[i] is not None:
                    torch.cuda.current_stream(self.device).wait_event(self.done_events[i])
            else:
                self.works[i].wait()
                self.flat[i].div_(self.world_size)

    def wire_bytes_per_step(self) -> int:
        n = sum(f.numel() * f.element_size() for f in self.flat)
        return int(2 * (self.world_size - 1) / self.world_size * n)
