"""Minimal ``Data`` / ``Batch`` stand-ins and the synthetic tree-cloud generator.

torch_geometric is not installable here, so the module contract of
/root/reference/pointnet2_regressor.py:52-53 (``data.x``, ``data.pos``, ``data.batch``) is
duck-typed: a real PyG ``Batch`` works unchanged, and these two tiny classes give tests and
``bench.py`` something with the same attributes (``x``, ``pos``, ``batch``, ``ptr``, ``y``) that
``torch_geometric.data.Batch.from_data_list`` would build.

Synthetic clouds follow SURVEY.md §8(d): seeded Gaussian clouds sigma=(4,4,8) m centred like
/root/reference/pointcloud_dataloader.py:108, intensity in [0,20) like
/root/reference/pointcloud_dataloader.py:42-44.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch


class Data:
    """One tree cloud: ``x [n,F]`` features, ``pos [n,3]``, ``y [4]`` biomass components."""

    def __init__(self, x: Optional[torch.Tensor] = None, pos: Optional[torch.Tensor] = None,
                 y: Optional[torch.Tensor] = None, **kwargs):
        self.x, self.pos, self.y = x, pos, y
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def num_nodes(self) -> int:
        return int(self.pos.size(0))

    def to(self, device, non_blocking: bool = False) -> "Data":
        out = self.__class__.__new__(self.__class__)
        for k, v in self.__dict__.items():
            setattr(out, k, v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v)
        return out

    def pin_memory(self) -> "Data":
        out = self.__class__.__new__(self.__class__)
        for k, v in self.__dict__.items():
            setattr(out, k, v.pin_memory() if torch.is_tensor(v) else v)
        return out


class Batch(Data):
    """Concatenation of clouds with ``batch`` (sorted cloud id per point) and ``ptr`` (offsets)."""

    @classmethod
    def from_data_list(cls, data_list: Sequence[Data]) -> "Batch":
        if len(data_list) == 0:
            raise ValueError("from_data_list needs at least one cloud")
        sizes = torch.tensor([d.num_nodes for d in data_list], dtype=torch.int64)
        ptr = torch.zeros(len(data_list) + 1, dtype=torch.int64)
        ptr[1:] = torch.cumsum(sizes, 0)
        out = cls()
        out.pos = torch.cat([d.pos for d in data_list], 0)
        out.x = None if data_list[0].x is None else torch.cat([d.x for d in data_list], 0)
        out.y = None if data_list[0].y is None else torch.cat([d.y.reshape(-1) for d in data_list], 0)
        out.batch = torch.repeat_interleave(torch.arange(len(data_list), dtype=torch.int64), sizes)
        out.ptr = ptr
        out.cloud_sizes = sizes.tolist()  # host copy: lets the model size its launches without a device read
        out.num_graphs = len(data_list)
        return out


_SIGMA = (4.0, 4.0, 8.0)
_Y_SCALE = (5.0, 6.0, 3.0, 40.0)


def synthetic_cloud(seed: int, num_points: int, num_features: int = 1, ragged: bool = False) -> Data:
    """Seeded synthetic tree cloud (CPU tensors; identical bits wherever it is regenerated)."""
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    n = int(num_points)
    if ragged:
        u = torch.rand(1, generator=g).item() * 0.2 + 0.9
        n = max(1, int(round(num_points * u)))
    pos = torch.randn(n, 3, generator=g) * torch.tensor(_SIGMA)
    pos = pos - pos.mean(0, keepdim=True)
    x = torch.rand(n, num_features, generator=g) * 20.0 if num_features > 0 else None
    y = torch.rand(4, generator=g) * torch.tensor(_Y_SCALE)
    return Data(x=x, pos=pos.contiguous(), y=y)


def synthetic_clouds(base_seed: int, num_clouds: int, num_points: int, num_features: int = 1,
                     ragged: bool = False) -> List[Data]:
    return [synthetic_cloud(base_seed + c, num_points, num_features, ragged) for c in range(num_clouds)]


def ptr_from_batch(batch: torch.Tensor, num_clouds: Optional[int] = None) -> torch.Tensor:
    """``ptr`` from a sorted ``batch`` vector (what PyG's fps wrapper derives)."""
    if num_clouds is None:
        num_clouds = int(batch.max().item()) + 1 if batch.numel() else 0
    counts = torch.bincount(batch, minlength=num_clouds)
    ptr = torch.zeros(num_clouds + 1, dtype=torch.int64, device=batch.device)
    ptr[1:] = torch.cumsum(counts, 0)
    return ptr
