"""Evaluation metrics of the reference: R2, RMSE and MAPE per biomass component and for their
sum, as /root/reference/testing_model.py:72-98 computes them with scikit-learn -- here as tensor ops on whatever device
the predictions live on, so an evaluation loop needs no ``.to('cpu')`` per batch (testing_model.py:64-65)."""
from __future__ import annotations

from typing import Dict

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
