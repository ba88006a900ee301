"""Data-parallel training: one process per GPU, bucketed gradient all-reduce overlapped with backward.

Replaces ``torch_geometric.nn.DataParallel`` of /root/reference/main.py:140 (single process, one Python
thread per GPU, full parameter broadcast + gradient reduce-to-root every step).
Here every rank owns a replica and its own shard of tree clouds; the only collective per step is an
all-reduce (average) of the gradients, issued bucket by bucket as soon as backward has produced a bucket
(head + SA3 first, then SA2, then SA1 = backward order).  The buckets are slices of the flat gradient buffer of
``optim.ParamArena``: libb2pn's backward kernels write into it directly, so a bucket is ready the moment the last
kernel of its level has been enqueued -- no flattening copies.  BatchNorm statistics stay per rank, as under
DataParallel.

Three ways to run a step, all through the same two calls ``prepare()`` ... backward ... ``finish()``:
  * eager: the post-accumulate hooks launch a bucket's all-reduce on a side stream while backward continues;
  * whole step captured in a CUDA graph (train.PipelinedTrainStep / GraphedTrainStep with ``capture_collective``):
    the same host code runs once, at capture time, and the collectives become nodes of the graph;
  * forward/backward captured, collective eager (``split``): ``prepare()`` and the hooks only ran at capture time, so
    ``finish()`` finds no bucket launched and launches ALL of them, every call (round-1 bug: it used to launch them
    only on the first call after capture).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from .optim import DEFAULT_BUCKETS, ParamArena


def shard_clouds(num_clouds: int, rank: int, world_size: int) -> range:
    """Contiguous shard of cloud indices for ``rank`` (clouds are the independent units of the path)."""
    per = (num_clouds + world_size - 1) // world_size
    return range(min(rank * per, num_clouds), min((rank + 1) * per, num_clouds))


def shard_by_points(sizes: Sequence[int], world_size: int) -> List[range]:
    """Contiguous chunks of clouds balanced by POINT count -- the rule of ``DataParallel.scatter`` in the reference's
    wrapper (cumulative node count, device = floor(G * midpoint / total)); used to shard an evaluation
    set of unequal clouds across ranks (/root/reference/testing_model.py:56-64)."""
    total = float(sum(sizes))
    out: List[List[int]] = [[] for _ in range(world_size)]
    cum = 0.0
    for i, n in enumerate(sizes):
        mid = cum + 0.5 * n
        dev = min(world_size - 1, int(world_size * mid / total)) if total > 0 else 0
        out[dev].append(i)
        cum += n
    ranges = []
    nxt = 0
    for chunk in out:
        ranges.append(range(nxt, nxt + len(chunk)))
        nxt += len(chunk)
    return ranges


class GradReducer:
    """Bucketed gradient all-reduce over the flat gradient buffer.  Backend-agnostic: NCCL on GPUs, gloo in the CPU
    tests."""

    def __init__(self, module: torch.nn.Module, buckets: Sequence[Sequence[str]] = DEFAULT_BUCKETS,
                 process_group=None, broadcast_params: bool = True, overlap: bool = True):
        if not dist.is_initialized():
            raise RuntimeError("GradReducer needs torch.distributed to be initialised")
        self.group = process_group
        # overlap=True: a bucket is reduced as soon as backward has filled it (the collective runs beside the remaining
        # backward kernels, on a side stream).  overlap=False: all buckets are reduced in finish(), after backward.
        self.overlap = overlap
        # inline=True (only with overlap=False): the all-reduce is issued on the CURRENT stream, no side stream/events
        self.inline = False
        self.world_size = dist.get_world_size(process_group)
        self.module = module
        self.arena = ParamArena.of(module) or ParamArena(module, buckets)
        a = self.arena
        self.bucket_params = a.bucket_params
        self.device = a.device
        self.flat: List[torch.Tensor] = [a.bucket_grads(i) for i in range(len(a.bucket_ranges))]
        self.pending = [len(ps) for ps in self.bucket_params]
        self.works: List[Optional[object]] = [None] * len(self.flat)
        self.comm_stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self.done_events: List[Optional[torch.cuda.Event]] = [None] * len(self.flat)
        self.allreduce_calls = 0   # collectives issued from the host so far (eager launches and capture-time ones)
        self.avg_in_collective = self.comm_stream is not None  # NCCL averages in the collective; gloo sums
        for p in a.params:
            p.register_post_accumulate_grad_hook(self._hook)
        if broadcast_params:  # replicas start identical (DataParallel re-broadcast every step; once is enough)
            dist.broadcast(a.flat_params, src=0, group=process_group)
            for t in module.buffers():
                dist.broadcast(t.data, src=0, group=process_group)

    def prepare(self) -> None:
        """Start of a step: forget last step's gradients (the backward kernels overwrite the arena) and re-arm."""
        self.arena.release_grads()
        self._rearm()

    def _rearm(self) -> None:
        for i in range(len(self.flat)):
            self.pending[i] = len(self.bucket_params[i])
            self.works[i] = None
            self.done_events[i] = None

    def _hook(self, p: torch.nn.Parameter) -> None:
        self.arena.fold(p)  # free when the kernel wrote into the arena; else one copy
        i = self.arena.bucket_of[p]
        self.pending[i] -= 1
        if self.pending[i] == 0 and self.overlap and self.works[i] is None:
            self._launch(i)

    def _launch(self, i: int) -> None:
        flat = self.flat[i]
        self.allreduce_calls += 1
        if self.comm_stream is None:  # gloo (CPU tests)
            self.works[i] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            return
        if self.inline or not self.overlap:
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
            self.works[i] = True
            self.done_events[i] = None
            return
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ready)
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
            ev = torch.cuda.Event()
            ev.record(self.comm_stream)
            self.done_events[i] = ev
        self.works[i] = True

    def finish(self) -> None:
        """Block the compute stream (not the host) until every bucket is reduced, then re-arm: a bucket nobody launched
        since the last ``finish()`` -- because its hooks did not all fire, or because forward/backward were a graph
        replay and no host code ran at all -- is launched here, on EVERY call."""
        for i in range(len(self.flat)):
            if self.works[i] is None:
                for p in self.bucket_params[i]:
                    self.arena.fold(p)
                self._launch(i)
            if self.comm_stream is not None:
                if self.done_events[i] is not None:
                    torch.cuda.current_stream(self.device).wait_event(self.done_events[i])
            else:
                self.works[i].wait()
                self.flat[i].div_(self.world_size)
        self._rearm()

    def wire_bytes_per_step(self) -> int:
        n = sum(f.numel() * f.element_size() for f in self.flat)
        return int(2 * (self.world_size - 1) / self.world_size * n)

    def replicas_identical(self) -> float:
        """max |param - rank 0's param| over the whole arena, reduced (MAX) over ranks: 0.0 iff the replicas are
        bit-identical -- the invariant data-parallel training must keep (bench.py asserts it after the timed loop)."""
        ref = self.arena.flat_params.clone()
        dist.broadcast(ref, src=0, group=self.group)
        d = (self.arena.flat_params - ref).abs().max().reshape(1)
        dist.all_reduce(d, op=dist.ReduceOp.MAX, group=self.group)
        return float(d.item())
