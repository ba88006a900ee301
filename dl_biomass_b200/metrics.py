"""Evaluation metrics of the reference: R2, RMSE and MAPE per biomass component and for their
sum, as /root/reference/testing_model.py:72-98 computes them with scikit-learn -- here as tensor ops on whatever device
the predictions live on, so an evaluation loop needs no ``.to('cpu')`` per batch (testing_model.py:64-65)."""
from __future__ import annotations

from typing import Dict, Sequence

import torch

COMPONENTS = ("bark_btphr", "branch_btphr", "foliage_btphr", "wood_btphr")   # column order of data.y (main.py:163-166)


def _r2(obs: torch.Tensor, pred: torch.Tensor) -> torch.Tensor:
    # sklearn.metrics.r2_score: 1 - SS_res / SS_tot (single output, uniform weights)
    ss_res = ((obs - pred) ** 2).sum()
    ss_tot = ((obs - obs.mean()) ** 2).sum()
    return 1.0 - ss_res / ss_tot


def _rmse(obs: torch.Tensor, pred: torch.Tensor) -> torch.Tensor:
    return ((obs - pred) ** 2).mean().sqrt()


def _mape(obs: torch.Tensor, pred: torch.Tensor) -> torch.Tensor:
    # sklearn.metrics.mean_absolute_percentage_error: |obs - pred| / max(|obs|, eps), eps = float64 machine epsilon
    eps = torch.finfo(torch.float64).eps
    return ((obs - pred).abs() / obs.abs().clamp_min(eps)).mean()


def regression_metrics(obs: torch.Tensor, pred: torch.Tensor) -> Dict[str, Dict[str, float]]:
    """``obs`` / ``pred``: [n_plots, 4] (bark, branch, foliage, wood).  Returns {component: {r2, rmse, mape}} for the four
    components and ``tree_btphr`` (their sum, testing_model.py:78-79), unrounded, computed in float64."""
    obs, pred = obs.reshape(-1, 4).double(), pred.reshape(-1, 4).double()
    cols = {name: (obs[:, i], pred[:, i]) for i, name in enumerate(COMPONENTS)}
    cols["tree_btphr"] = (obs.sum(1), pred.sum(1))
    stacked = torch.stack([torch.stack([_r2(o, p), _rmse(o, p), _mape(o, p)]) for o, p in cols.values()]).cpu()
    return {name: {"r2": float(stacked[i, 0]), "rmse": float(stacked[i, 1]), "mape": float(stacked[i, 2])}
            for i, name in enumerate(cols)}


@torch.no_grad()
def evaluate(model: torch.nn.Module, batches, return_predictions: bool = False):
    """The evaluation loop of /root/reference/testing_model.py:60-98: ``model.eval()``, predictions for every batch,
    observed values from ``batch.y``, the metric table at the end.  The reference pushes the whole test set through
    the model as ONE batch and moves predictions to the CPU per batch; here ``batches`` may be any iterable of batches
    (e.g. chunks of 256 clouds) and everything stays on the device until the final table: one device->host read in
    total.  The model's training flag is restored afterwards.  Returns the ``regression_metrics`` table (and, with
    ``return_predictions``, the stacked ``(obs, pred)`` tensors as a second value)."""
    was_training = model.training
    model.eval()
    obs, pred = [], []
    try:
        for b in batches:
            out = model(b)
            pred.append(out.reshape(-1, 4))
            obs.append(b.y.reshape(-1, 4).to(out.device))
    finally:
        model.train(was_training)
    if not pred:
        raise ValueError("evaluate needs at least one batch")
    o, p = torch.cat(obs, 0), torch.cat(pred, 0)
    table = regression_metrics(o, p)
    return (table, (o, p)) if return_predictions else table


@torch.no_grad()
def evaluate_distributed(model: torch.nn.Module, clouds: Sequence, device, batch_size: int = 256, process_group=None,
                         return_predictions: bool = False):
    """The same evaluation over several GPUs, one process per GPU (the reference evaluates through its DataParallel
    wrapper, /root/reference/testing_model.py:30-37,56-64, which scatters the ONE giant batch over the devices by point
    count and gathers the ``[B, 4]`` outputs on device 0 -- SURVEY.md A.8, rows C3 / 8(e)).  ``clouds``: the whole test set
    as a list of per-cloud ``Data`` (what ``DataListLoader`` yields), identical on every rank.  Rank r takes the r-th
    contiguous chunk balanced by point count (``parallel.shard_by_points``), runs it in batches of ``batch_size`` clouds,
    and ONE all-gather of the ``[n, 8]`` (observed | predicted) rows gives every rank the full table, in input order."""
    import torch.distributed as dist
    from .data import Batch
    from .parallel import shard_by_points
    if not dist.is_initialized():
        raise RuntimeError("evaluate_distributed needs torch.distributed to be initialised")
    world, rank = dist.get_world_size(process_group), dist.get_rank(process_group)
    parts = shard_by_points([int(d.pos.size(0)) for d in clouds], world)
    mine = [clouds[i] for i in parts[rank]]
    was_training = model.training
    model.eval()
    rows = []
    try:
        for i in range(0, len(mine), batch_size):
            b = Batch.from_data_list(mine[i:i + batch_size]).to(device)
            out = model(b).reshape(-1, 4).to(torch.float32)
            rows.append(torch.cat([b.y.reshape(-1, 4).to(out.device, torch.float32), out], 1))
    finally:
        model.train(was_training)
    counts = [len(r) for r in parts]
    dev = torch.device(device)
    local = torch.zeros(max(counts + [1]), 8, dtype=torch.float32, device=dev)
    if rows:
        local[:counts[rank]] = torch.cat(rows, 0)
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local, group=process_group)
    full = torch.cat([g[:c] for g, c in zip(gathered, counts)], 0)
    if full.size(0) == 0:
        raise ValueError("evaluate_distributed needs at least one cloud")
    o, p = full[:, :4].contiguous(), full[:, 4:].contiguous()
    table = regression_metrics(o, p)
    return (table, (o, p)) if return_predictions else table
