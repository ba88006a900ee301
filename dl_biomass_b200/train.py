"""Loss and optimiser of the reference's training loop (/root/reference/main.py:146-176), host side."""
from __future__ import annotations

import torch

# bark, branch, foliage, wood shares of total biomass: /root/reference/main.py:163-166
LOSS_WEIGHTS = (1.0 / 11.0, 1.0 / 12.0, 1.0 / 5.0, 1.0 / 72.0)
ADAM_LR = 0.00179966410046844          # main.py:38
ADAM_WEIGHT_DECAY = 8.0250963438986e-05  # main.py:39


def weighted_mse_loss(outs: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """sum_c w_c * mse(y[:, c], outs[:, c])  (main.py:154-169); ``y`` may come flat ([4B]) as PyG collates it."""
    y = y.reshape(outs.size(0), 4).to(outs.dtype)
    w = torch.tensor(LOSS_WEIGHTS, dtype=outs.dtype, device=outs.device)
    return (((outs - y) ** 2).mean(0) * w).sum()


def make_optimizer(params, lr: float = ADAM_LR, weight_decay: float = ADAM_WEIGHT_DECAY) -> torch.optim.Optimizer:
    """torch.optim.Adam(model.parameters(), lr, weight_decay) of main.py:84 (L2-in-gradient form)."""
    params = list(params)
    fused = len(params) > 0 and params[0].is_cuda
    return torch.optim.Adam(params, lr=lr, weight_decay=weight_decay, fused=fused)


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
