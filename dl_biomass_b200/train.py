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


def weighted_mse_loss(outs: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """sum_c w_c * mse(y[:, c], outs[:, c])  (main.py:154-169); ``y`` may come flat ([4B]) as PyG collates it."""
    y = y.reshape(outs.size(0), 4).to(outs.dtype)
    return (((outs - y) ** 2).mean(0) * _loss_weights(outs.dtype, outs.device)).sum()


def make_optimizer(params, lr: float = ADAM_LR, weight_decay: float = ADAM_WEIGHT_DECAY,
                   capturable: bool = False) -> torch.optim.Optimizer:
    """torch.optim.Adam(model.parameters(), lr, weight_decay) of main.py:84 (L2-in-gradient form).
    ``capturable=True`` keeps the step counter on the device so the update can live in a CUDA graph."""
    params = list(params)
    fused = len(params) > 0 and params[0].is_cuda
    return torch.optim.Adam(params, lr=lr, weight_decay=weight_decay, fused=fused, capturable=capturable and fused)


def train_step(model, optimizer, batch, reducer=None) -> torch.Tensor:
    """One iteration of main.py:150-172.  ``reducer`` is the data-parallel gradient reducer (parallel.py)."""
    optimizer.zero_grad(set_to_none=reducer is None)
    if reducer is not None:
        reducer.prepare()
    outs = model(batch)
    loss = weighted_mse_loss(outs, batch.y)
    loss.backward()
    if reducer is not None:
        reducer.finish()
    optimizer.step()
    return loss.detach()


class GraphedTrainStep:
    """``train_step`` captured once into a CUDA graph and replayed: one graph launch per iteration instead of
    ~150 kernel launches (the regression head alone is ~60 tiny ATen kernels).  Valid for batches with the same
    cloud sizes as the example batch -- the reference trains on fixed-size resampled clouds
    (/root/reference/main.py:55-57: 7 168 points per plot) -- anything else falls back to the eager step.
    The optimiser must be built with ``make_optimizer(..., capturable=True)``."""

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
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                train_step(model, optimizer, self.static, reducer)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        lib = _lib.lib()
        l0 = lib.b2pn_launch_count()
        with torch.cuda.graph(self.graph):
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
