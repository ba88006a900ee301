"""Loss and optimiser of the reference's training loop (/root/reference/main.py:146-176), host side."""
from __future__ import annotations

import torch

# bark, branch, foliage, wood shares of total biomass: /root/reference/main.py:163-166
LOSS_WEIGHTS = (1.0 / 11.0, 1.0 / 12.0, 1.0 / 5.0, 1.0 / 72.0)
ADAM_LR = 0.00179966410046844          # main.py:38
ADAM_WEIGHT_DECAY = 8.0250963438986e-05  # main.py:39


_LOSS_W: dict = {}


def _loss_weights(dtype, device) -> torch.Tensor:
    key = (dtype, str(device))
    w = _LOSS_W.get(key)
    if w is None:
        w = torch.tensor(LOSS_WEIGHTS, dtype=dtype, device=device)
        _LOSS_W[key] = w
    return w


class _WeightedMSE(torch.autograd.Function):
    """Loss value and its gradient from one kernel (include/b2pn.h, b2pn_weighted_mse)."""

    @staticmethod
    def forward(ctx, outs, y, w):
        from . import _lib
        outs = outs.contiguous()
        loss = torch.empty((), dtype=torch.float32, device=outs.device)
        grad = torch.empty_like(outs) if ctx.needs_input_grad[0] else None
        with torch.cuda.device(outs.device):
            rc = _lib.lib().b2pn_weighted_mse(outs.data_ptr(), y.data_ptr(), w.data_ptr(), outs.size(0), outs.size(1),
                                              loss.data_ptr(), None if grad is None else grad.data_ptr(),
                                              torch.cuda.current_stream(outs.device).cuda_stream)
        _lib.check(rc, "b2pn_weighted_mse")
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None


def weighted_mse_loss(outs: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """sum_c w_c * mse(y[:, c], outs[:, c])  (main.py:154-169); ``y`` may come flat ([4B]) as PyG collates it.
    Value and gradient come from one libb2pn launch, in fp32; there is no CPU path."""
    if not outs.is_cuda:
        raise RuntimeError("weighted_mse_loss runs on a B200 only: there is no CPU fallback")
    outs = outs.to(torch.float32)
    y = y.reshape(outs.size(0), 4).to(device=outs.device, dtype=torch.float32).contiguous()
    return _WeightedMSE.apply(outs, y, _loss_weights(torch.float32, outs.device))


def loss_and_grad(outs: torch.Tensor, y: torch.Tensor):
    """Loss value and d loss / d outs from ONE libb2pn launch, outside autograd: the training step feeds the gradient to
    ``outs.backward(grad)`` directly instead of letting autograd build ones_like(loss) and scale by it."""
    from . import _lib
    if not outs.is_cuda:
        raise RuntimeError("the training loss runs on a B200 only: there is no CPU fallback")
    o = outs.detach().to(torch.float32).contiguous()
    y = y.reshape(o.size(0), 4).to(device=o.device, dtype=torch.float32).contiguous()
    w = _loss_weights(torch.float32, o.device)
    loss = torch.empty((), dtype=torch.float32, device=o.device)
    grad = torch.empty_like(o)
    with torch.cuda.device(o.device):
        rc = _lib.lib().b2pn_weighted_mse(o.data_ptr(), y.data_ptr(), w.data_ptr(), o.size(0), o.size(1), loss.data_ptr(),
                                          grad.data_ptr(), torch.cuda.current_stream(o.device).cuda_stream)
    _lib.check(rc, "b2pn_weighted_mse")
    return loss, grad


def make_optimizer(params, lr: float = ADAM_LR, weight_decay: float = ADAM_WEIGHT_DECAY,
                   capturable: bool = False):
    """The optimiser of main.py:84, ``Adam(model.parameters(), lr, weight_decay)`` (L2-in-gradient form).

    Hand it the MODULE (already on its GPU) and it returns ``optim.FlatAdam``: every parameter re-homed into one flat
    arena, one libb2pn launch per step, the step counter on the device (replayable from a CUDA graph).  Handed an
    iterable of parameters it returns ``torch.optim.Adam`` (ATen's fused kernel) as the reference does;
    ``capturable=True`` then keeps its step counter on the device."""
    if isinstance(params, torch.nn.Module):
        from .optim import FlatAdam
        return FlatAdam(params, lr=lr, weight_decay=weight_decay)
    params = list(params)
    fused = len(params) > 0 and params[0].is_cuda
    return torch.optim.Adam(params, lr=lr, weight_decay=weight_decay, fused=fused, capturable=capturable and fused)


def forward_backward(model, optimizer, batch, reducer=None, sampling=None, after_grouping=None,
                     before_level1_backward=None) -> torch.Tensor:
    """zero_grad + forward + loss + backward of main.py:150-171 (everything of a step that involves no collective)."""
    optimizer.zero_grad(set_to_none=True)
    if reducer is not None:
        reducer.prepare()
    if sampling is None and after_grouping is None and before_level1_backward is None:
        outs = model(batch)
    else:
        outs = model(batch, sampling=sampling, after_grouping=after_grouping,
                     before_level1_backward=before_level1_backward)
    loss, grad = loss_and_grad(outs, batch.y)
    outs.backward(grad)
    return loss


def reduce_and_update(optimizer, reducer=None) -> None:
    """gradient all-reduce (data parallel) + optimizer.step() of main.py:172."""
    if reducer is not None:
        reducer.finish()
    optimizer.step()


def train_step(model, optimizer, batch, reducer=None, sampling=None, after_grouping=None,
               before_level1_backward=None) -> torch.Tensor:
    """One iteration of main.py:150-172.  ``reducer`` is the data-parallel gradient reducer (parallel.py);
    ``sampling`` an optional ``Net.sample(batch)`` computed ahead of time."""
    loss = forward_backward(model, optimizer, batch, reducer, sampling, after_grouping, before_level1_backward)
    reduce_and_update(optimizer, reducer)
    return loss


class _TrainingState:
    """Everything a training step mutates besides the gradients: parameters, BatchNorm buffers, optimiser moments and
    step count, and the device-side random-number counters (dropout, FPS start).  The graph-capturing steppers run real
    warm-up steps before capture (NCCL communicators, lazily-built kernels attributes and the caching allocator all want
    that); they snapshot this first and restore it afterwards IN PLACE (the graphs hold the pointers), so building a
    stepper leaves the model exactly as it found it: the reference trains on every batch once per epoch
    (/root/reference/main.py:150-172) and so does this."""

    def __init__(self, model, optimizer):
        self.model, self.optimizer = model, optimizer
        self.tensors = [t for t in list(model.parameters()) + list(model.buffers())]
        self.saved = [t.detach().clone() for t in self.tensors]
        self.counters = []
        for m in model.modules():
            for name in ("_head_rng_counter", "_fps_rng_state"):
                t = getattr(m, name, None)
                if torch.is_tensor(t):
                    self.counters.append((t, t.clone()))
        if hasattr(optimizer, "_snapshot"):
            self.opt = optimizer._snapshot()
        else:
            self.opt = [(t, t.clone()) for st in optimizer.state.values() for t in st.values() if torch.is_tensor(t)]
            self.opt_keys = set(optimizer.state.keys())

    def restore(self) -> None:
        with torch.no_grad():
            for t, s in zip(self.tensors, self.saved):
                t.copy_(s)
            for t, s in self.counters:
                t.copy_(s)
            if hasattr(self.optimizer, "_restore"):
                self.optimizer._restore(self.opt)
            else:
                seen = {id(t) for t, _ in self.opt}
                for t, s in self.opt:
                    t.copy_(s)
                # state tensors created lazily by the warm-up steps: back to their initial values, in place
                for st in self.optimizer.state.values():
                    for t in st.values():
                        if torch.is_tensor(t) and id(t) not in seen:
                            t.zero_()
            # counters created lazily during warm-up start from zero
            for m in self.model.modules():
                for name in ("_head_rng_counter", "_fps_rng_state"):
                    t = getattr(m, name, None)
                    if torch.is_tensor(t) and all(t is not c for c, _ in self.counters):
                        t.zero_()


class GraphedTrainStep:
    """``train_step`` captured once into a CUDA graph and replayed: one graph launch per iteration instead of
    ~150 kernel launches.  Valid for batches with the same cloud sizes as the example batch -- the reference trains on
    fixed-size resampled clouds (/root/reference/main.py:55-57: 7 168 points per plot) -- anything else falls back to the
    eager step.  The optimiser must be ``make_optimizer(model)`` (or torch's Adam with ``capturable=True``).
    Construction leaves model, optimiser and random-number state untouched (``_TrainingState``)."""

    def __init__(self, model, optimizer, example_batch, reducer=None, warmup: int = 3):
        from . import _lib
        self.model, self.optimizer, self.reducer = model, optimizer, reducer
        dev = example_batch.pos.device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTrainStep needs the batch on a B200")
        self.sizes = tuple(example_batch.cloud_sizes)
        self.static = example_batch.to(dev)  # private copies of pos / x / y / batch: the graph reads these
        for k in ("pos", "x", "y", "batch"):
            v = getattr(example_batch, k, None)
            setattr(self.static, k, None if v is None else v.clone())
        state = _TrainingState(model, optimizer)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                train_step(model, optimizer, self.static, reducer)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        state.restore()
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        lib = _lib.lib()
        l0 = lib.b2pn_launch_count()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.loss = train_step(model, optimizer, self.static, reducer)
        self.launches_per_replay = int(lib.b2pn_launch_count() - l0)

    def matches(self, batch) -> bool:
        return tuple(getattr(batch, "cloud_sizes", ())) == self.sizes

    def __call__(self, batch) -> torch.Tensor:
        if not self.matches(batch):
            return train_step(self.model, self.optimizer, batch, self.reducer)
        st = self.static
        st.pos.copy_(batch.pos, non_blocking=True)
        if st.x is not None:
            st.x.copy_(batch.x, non_blocking=True)
        if st.y is not None:
            st.y.copy_(batch.y, non_blocking=True)
        self.graph.replay()
        return self.loss


class PipelinedTrainStep:
    """Training step that overlaps everything that depends on the point positions only -- farthest-point sampling,
    ball query, row compaction, the gathered level-1 message inputs (``Net.sample``) -- of the NEXT batch with the
    training of the current one.

    FPS is a chain of ~2 500 dependent arg-max iterations per batch: it keeps one SM per cloud busy for ~1.4 ms
    (12 of 148 SMs at the reference's batch size) while the rest of the GPU idles.  ``step(next_batch)`` therefore
    runs ``Net.sample(next_batch)`` on a second stream while the current batch goes through forward / loss /
    backward / all-reduce / Adam, and returns the loss of the current batch; the persistent tensor-core kernels are told
    to leave one SM per cloud free (``b2pn_sa_args::sm_limit`` through ``sa.set_sm_limit``, a per-thread / per-call
    option; ``cap=False`` disables it).  ``join``: where the training stream waits for the branch -- "end" (after the
    step; measured best, round 1) or "backward" (inside the backward pass, before the level-2 backward, lifting the SM
    cap from there on).  With ``graph=True`` both branches are captured in CUDA graphs (fork / join inside the graph; two
    graphs over a double buffer, replayed alternately, so that no hand-over copy sits on the training stream) and every
    call is a single replay; the batches must then keep the cloud sizes of the example batch (``graph=False`` takes
    ragged batches, e.g. after the reference's point-removal / duplication augmentation).
    ``uncap_level1_backward``: with the join at the end, launch the last quarter of the step (the level-1 backward, by
    then the branch is past its farthest-point sampling) on every SM again.  ``grouping=False`` leaves ball query and row
    compaction inside forward; ``aux=True`` (default; 2.18 -> 2.09 ms/step, round 2) runs the level-1 grouping on a third
    stream beside the level-2 sampling.

    Data parallel (``reducer``): with ``capture_collective=True`` (default) the bucketed all-reduce and the optimiser are
    part of the graph (``overlap_collective``: buckets go out on a side stream as backward finishes them; else after
    backward on the training stream); with ``capture_collective=False`` the graph holds forward + backward and the
    all-reduce + optimiser follow every replay eagerly.

    Building the stepper does not train: the warm-up iterations it needs run on a snapshot that is restored before
    capture.  The pipeline holds one batch in flight; ``flush()`` trains it at the end of an epoch.

        stepper = PipelinedTrainStep(model, opt, first_batch)       # also samples first_batch
        for nxt in loader:                                          # loader yields the batches after the first
            loss = stepper.step(nxt)                                # trains on the batch submitted before
        loss = stepper.flush()                                      # trains on the last batch
    """

    def __init__(self, model, optimizer, first_batch, reducer=None, graph: bool = True, warmup: int = 2,
                 join: str = "end", cap: bool = True, grouping: bool = True, aux: bool = True,
                 uncap_level1_backward: bool = True, capture_collective: bool = True,
                 overlap_collective: bool = False, side_priority: int = 0):
        from . import _lib, sa
        if join not in ("backward", "end"):
            raise ValueError("join must be 'backward' or 'end'")
        self.join_at, self.grouping = join, bool(grouping)
        self.uncap_l1 = bool(uncap_level1_backward)
        self.model, self.optimizer, self.reducer = model, optimizer, reducer
        self.dev = first_batch.pos.device
        if self.dev.type != "cuda":
            raise RuntimeError("PipelinedTrainStep needs the batch on a B200")
        self.lib = _lib.lib()
        self._sa = sa
        self.side = torch.cuda.Stream(self.dev, priority=int(side_priority))   # < 0: the branch's kernels are scheduled first
        self.aux = torch.cuda.Stream(self.dev) if (aux and grouping) else None
        self.sizes = tuple(first_batch.cloud_sizes)
        sms = torch.cuda.get_device_properties(self.dev).multi_processor_count
        ncl = len(self.sizes)
        self.sm_limit = sms - ncl if (cap and ncl * 4 <= sms) else 0
        # data parallel + graph, split mode: forward/backward are captured, the all-reduce and the optimiser follow the
        # replay eagerly
        self.split = reducer is not None and bool(graph) and not capture_collective
        if reducer is not None:
            reducer.overlap = bool(overlap_collective) and not self.split
            reducer.inline = not reducer.overlap  # on the current stream, no side stream needed
        self.graph = None
        self.launches_per_step = 0
        self.allreduce_per_step = 0
        self.pending = True  # a batch is in flight (submitted, not trained on yet)
        if graph:
            self._capture(first_batch, warmup)
        else:
            self.cur = first_batch
            self.cur_sampling = model.sample(first_batch, grouping=self.grouping, aux_stream=self.aux)

    def close(self) -> None:
        self._sa.set_sm_limit(0)

    def release_graphs(self) -> None:
        """Destroy the captured graphs.  REQUIRED before ``torch.distributed.destroy_process_group()`` when the
        collectives were captured: NCCL keeps a communicator alive (and ``ncclCommDestroy`` waits) while a CUDA graph still
        holds one of its persistent plans -- measured in round 2: the 2-GPU bench printed its line and then never
        exited."""
        import gc
        torch.cuda.synchronize(self.dev)
        self.graphs, self.graph, self.losses = None, None, None
        self.pending = False
        gc.collect()
        torch.cuda.synchronize(self.dev)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    # ---- eager -------------------------------------------------------------------------------------------------
    def _eager_step(self, cur, cur_sampling, nxt, publish=None, sample_next=True):
        main = torch.cuda.current_stream(self.dev)
        nxt_sampling = None
        if sample_next:
            self.side.wait_stream(main)                   # nxt's tensors (an H2D copy, say) are ordered before this
            with torch.cuda.stream(self.side):
                nxt_sampling = self.model.sample(nxt, grouping=self.grouping, aux_stream=self.aux)
                if publish is not None:                   # graph mode: into the other half of the double buffer
                    for d, s in zip(publish.tensors(), nxt_sampling.tensors()):
                        d.copy_(s)
        # The sampling branch (~1.8 ms) must not fight the persistent tcgen05 kernels for SMs: while it may be in
        # flight they are launched with a capped grid.  join="end": the whole step runs capped and waits for the
        # branch after the optimiser (the HBM-bound kernels lose little on 136 of 148 SMs; 2.59 vs 2.93 ms/step).
        # join="backward": the wait sits in the backward pass right before the level-2 backward and lifts the cap.
        # the options holder of THIS thread: forward remembers it, backward (autograd's thread) reads it when it runs,
        # so the closures below change the cap for the kernels launched after them, whichever thread launches them
        opts = self._sa.current_options()
        opts.sm_limit = self.sm_limit if sample_next else 0

        def set_cap(n):
            opts.sm_limit = n

        def join():
            if sample_next:
                main.wait_stream(self.side)
            set_cap(0)

        step = forward_backward if self.split else train_step
        early = self.join_at == "backward"
        uncap = (lambda: set_cap(0)) if (self.uncap_l1 and not early) else None
        loss = step(self.model, self.optimizer, cur, self.reducer, sampling=cur_sampling,
                    after_grouping=join if early else None, before_level1_backward=uncap)
        if not early:
            join()
        return loss, nxt_sampling

    # ---- graph ---------------------------------------------------------------------------------------------------
    @staticmethod
    def _clone_batch(b):
        out = b.to(b.pos.device)
        for k in ("pos", "x", "y", "batch"):
            v = getattr(b, k, None)
            setattr(out, k, None if v is None else v.clone())
        return out

    @staticmethod
    def _copy_batch(dst, src):
        dst.pos.copy_(src.pos, non_blocking=True)
        if dst.x is not None:
            dst.x.copy_(src.x, non_blocking=True)
        if dst.y is not None:
            dst.y.copy_(src.y, non_blocking=True)

    def _capture(self, first_batch, warmup):
        # Double buffer: graph i trains on bufs[i] with samp[i] while its sampling branch works on bufs[1 - i] and
        # leaves the result in samp[1 - i]; step() alternates between the two graphs, so nothing is copied on the
        # training stream (the publishing copies ride on the sampling stream, which has slack).
        self.bufs = [self._clone_batch(first_batch), self._clone_batch(first_batch)]
        state = _TrainingState(self.model, self.optimizer)   # warm-up must leave no trace (see _TrainingState)
        samp = self.model.sample(self.bufs[0], grouping=self.grouping)
        self.samp = [samp.clone(), samp.clone()]

        def body(i):
            loss, _ = self._eager_step(self.bufs[i], self.samp[i], self.bufs[1 - i], publish=self.samp[1 - i])
            return loss

        warm = torch.cuda.Stream(self.dev)
        warm.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(warm):
            for _ in range(max(warmup, 1)):
                for i in (0, 1):
                    body(i)
                    if self.split:
                        reduce_and_update(self.optimizer, self.reducer)
        torch.cuda.current_stream(self.dev).wait_stream(warm)
        torch.cuda.synchronize(self.dev)
        state.restore()
        # the sampling of the first batch again, now with the restored random-number counters: samp[0] is what the
        # first step() trains with
        samp = self.model.sample(self.bufs[0], grouping=self.grouping)
        for d, s in zip(self.samp[0].tensors(), samp.tensors()):
            d.copy_(s)
        torch.cuda.synchronize(self.dev)
        self.graphs, self.losses = [], []
        for i in (0, 1):
            g = torch.cuda.CUDAGraph()
            l0 = self.lib.b2pn_launch_count()
            c0 = self.reducer.allreduce_calls if self.reducer is not None else 0
            with torch.cuda.graph(g, pool=None if i == 0 else self.graphs[0].pool(), capture_error_mode="thread_local"):
                self.losses.append(body(i))
            self.launches_per_step = int(self.lib.b2pn_launch_count() - l0)
            if self.reducer is not None:
                self.allreduce_per_step = len(self.reducer.flat) if self.split else self.reducer.allreduce_calls - c0
            self.graphs.append(g)
        self.graph = self.graphs[0]
        self.parity = 0

    def step(self, next_batch) -> torch.Tensor:
        """Train on the batch submitted by the previous call (or the constructor) while sampling ``next_batch``."""
        if self.graph is not None and tuple(getattr(next_batch, "cloud_sizes", ())) != self.sizes:
            raise ValueError("PipelinedTrainStep(graph=True): the batch layout (cloud sizes) must stay fixed; "
                             "use graph=False for ragged batches")
        if not self.pending:
            raise RuntimeError("the pipeline was flushed: build a new stepper (or call restart(batch)) to go on")
        if self.graph is not None:
            i = self.parity
            self._copy_batch(self.bufs[1 - i], next_batch)
            self.graphs[i].replay()
            if self.split:
                reduce_and_update(self.optimizer, self.reducer)
            self.parity = 1 - i
            return self.losses[i]
        l0 = self.lib.b2pn_launch_count()
        c0 = self.reducer.allreduce_calls if self.reducer is not None else 0
        nb = next_batch if next_batch.pos.is_cuda else next_batch.to(self.dev, non_blocking=True)
        loss, nxt_sampling = self._eager_step(self.cur, self.cur_sampling, nb)
        main = torch.cuda.current_stream(self.dev)
        for t in nxt_sampling.tensors():
            t.record_stream(main)  # allocated on the side stream, consumed on this one in the next call
        self.cur, self.cur_sampling = nb, nxt_sampling
        self.launches_per_step = int(self.lib.b2pn_launch_count() - l0)
        if self.reducer is not None:
            self.allreduce_per_step = self.reducer.allreduce_calls - c0
        return loss

    def flush(self) -> torch.Tensor:
        """Train on the batch still in flight WITHOUT submitting a new one (end of an epoch): every batch handed to the
        stepper is trained on exactly once.  Graph mode replays the step once more (its sampling branch re-samples the
        stale contents of the other buffer, which nobody reads); eager mode runs the plain step."""
        if not self.pending:
            raise RuntimeError("nothing in flight")
        self.pending = False
        if self.graph is not None:
            i = self.parity
            self.graphs[i].replay()
            if self.split:
                reduce_and_update(self.optimizer, self.reducer)
            self.parity = 1 - i
            return self.losses[i]
        loss, _ = self._eager_step(self.cur, self.cur_sampling, None, sample_next=False)
        if self.split:
            reduce_and_update(self.optimizer, self.reducer)
        return loss

    def restart(self, batch) -> None:
        """Submit ``batch`` as the first one of a new epoch after ``flush()``."""
        if self.pending:
            raise RuntimeError("a batch is still in flight: flush() first")
        if self.graph is not None:
            if tuple(getattr(batch, "cloud_sizes", ())) != self.sizes:
                raise ValueError("the batch layout (cloud sizes) must stay fixed")
            i = self.parity
            self._copy_batch(self.bufs[i], batch)
            samp = self.model.sample(self.bufs[i], grouping=self.grouping)
            for d, s in zip(self.samp[i].tensors(), samp.tensors()):
                d.copy_(s)
        else:
            self.cur = batch if batch.pos.is_cuda else batch.to(self.dev, non_blocking=True)
            self.cur_sampling = self.model.sample(self.cur, grouping=self.grouping, aux_stream=self.aux)
        self.pending = True
