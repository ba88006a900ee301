"""Resident cloud cache + on-device training augmentation.

The reference re-reads one LAS file and the biomass CSV per sample and augments it in numpy on the loader's CPU
workers (/root/reference/augmentation.py:258-307, /root/reference/pointcloud_dataloader.py:168-204).  Here the whole
(resampled) training set sits in HBM once -- 7 168 points x 16 B per plot: a thousand plots are 115 MB -- and a batch is
one kernel launch (libb2pn's augment.cu): ``point_removal -> random_noise -> rotate_points`` of
augmentation.py:54-122 for every cloud of the batch, written straight into the concatenated (pos, x, batch) layout
``Net.forward`` takes, together with the host-side ``ptr`` / ``cloud_sizes`` (the per-cloud sizes are drawn on the
host, so no device->host read is needed to size the model's launches).

There is no CPU path: the cache lives on a B200.
"""
from __future__ import annotations

import ctypes
import math
import random
from typing import List, Optional, Sequence

import torch

from . import _lib
from .data import Batch, Data

MIN_POINTS = 100   # augmentation.py:305-306: samples with fewer points are dropped from the batch


def draw_scalars(rng: random.Random, n: int):
    """Per-sample scalar draws with the reference's ranges: n_keep = randint(round(0.9 n), n)
    (augmentation.py:79), sd ~ U(0.01, 0.025) (:94), added iff U(0,1) >= 0.5 (:97), n_dup = randint(0,
    round(0.1 n_keep)) (:115), angle ~ U(-180, 180) degrees (:55).  Returns (n_keep, n_dup, signed sd, angle)."""
    n_keep = rng.randint(round(n * 0.9), n)
    sd = rng.uniform(0.01, 0.025)
    add = rng.uniform(0.0, 1.0) >= 0.5
    n_dup = rng.randint(0, round(n_keep * 0.1))
    angle = rng.uniform(-180.0, 180.0)
    return n_keep, n_dup, (sd if add else -sd), angle


class CloudCache:
    """All clouds of a dataset, concatenated, resident on the GPU: ``pos [N,3]``, ``x [N,F]`` (or None), ``y [C,4]``."""

    def __init__(self, clouds: Sequence[Data], device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("CloudCache lives on a B200: there is no CPU fallback")
        if len(clouds) == 0:
            raise ValueError("CloudCache needs at least one cloud")
        self.device = device
        self.sizes: List[int] = [int(c.pos.shape[0]) for c in clouds]
        self.offsets: List[int] = [0]
        for n in self.sizes:
            self.offsets.append(self.offsets[-1] + n)
        self.pos = torch.cat([c.pos.to(torch.float32) for c in clouds], 0).contiguous().to(device)
        has_x = clouds[0].x is not None
        self.x = torch.cat([c.x.to(torch.float32) for c in clouds], 0).contiguous().to(device) if has_x else None
        self.num_features = int(self.x.shape[1]) if has_x else 0
        has_y = getattr(clouds[0], "y", None) is not None
        self.y = torch.stack([c.y.reshape(-1).to(torch.float32) for c in clouds], 0).to(device) if has_y else None
        mx = int(_lib.lib().b2pn_augment_max_points())
        if max(self.sizes) > mx:
            raise NotImplementedError(f"on-device augmentation handles clouds of up to {mx} points")

    def __len__(self) -> int:
        return len(self.sizes)

    def plan(self, cloud_ids: Sequence[int], rng: random.Random, epoch: int = 0, augment: bool = True):
        """Host side of a batch: the scalar draws and the output layout.  Returns a list of per-cloud records
        (cloud id, n_src, n_keep, n_dup, signed sd, angle in degrees, uid); clouds that would end up with fewer than
        ``MIN_POINTS`` points are dropped like the reference's collate drops ``None`` samples."""
        recs = []
        for cid in cloud_ids:
            n = self.sizes[cid]
            if augment:
                n_keep, n_dup, sd, angle = draw_scalars(rng, n)
            else:
                n_keep, n_dup, sd, angle = n, 0, 0.0, 0.0
            if n_keep + n_dup < MIN_POINTS:
                continue
            # uid = the per-point random stream of THIS sample: fresh 64 bits per record.  The reference puts every cloud
            # into an epoch 1 + num_augs times, each copy augmented independently (/root/reference/main.py:100-112); a uid
            # derived from (epoch, cloud id) alone would give all those copies the same permutation and the same noise.
            uid = ((rng.getrandbits(64) ^ (int(epoch) * 0x9E3779B97F4A7C15)) & 0xFFFFFFFFFFFFFFFF) if augment \
                else (int(epoch) * len(self.sizes) + int(cid))
            recs.append((int(cid), n, n_keep, n_dup, sd, angle, uid))
        return recs

    def batch(self, cloud_ids: Sequence[int], rng: Optional[random.Random] = None, seed: int = 0, epoch: int = 0,
              augment: bool = True, plan=None, return_source: bool = False) -> Batch:
        """An (augmented) training batch of the given clouds, assembled on the device in one launch per 64 clouds.
        ``augment=False`` copies the clouds unchanged except for their point ORDER, which is still shuffled (use
        ``Batch.from_data_list`` for evaluation batches that must keep the stored order)."""
        recs = plan if plan is not None else self.plan(cloud_ids, rng or random.Random((int(seed) << 24) ^ int(epoch)), epoch, augment)
        if len(recs) == 0:
            raise ValueError("no cloud of this batch has enough points")
        B = len(recs)
        arr = (_lib.AugmentCloud * B)()
        sizes, off = [], 0
        for i, (cid, n, n_keep, n_dup, sd, angle, uid) in enumerate(recs):
            a = math.radians(angle)
            arr[i].src_off, arr[i].out_off, arr[i].uid = self.offsets[cid], off, uid
            arr[i].n_src, arr[i].n_keep, arr[i].n_dup = n, n_keep, n_dup
            arr[i].noise_sd, arr[i].cos_a, arr[i].sin_a = sd, math.cos(a), math.sin(a)
            sizes.append(n_keep + n_dup)
            off += n_keep + n_dup
        dev, F = self.device, self.num_features
        out = Batch()
        out.pos = torch.empty(off, 3, dtype=torch.float32, device=dev)
        out.x = torch.empty(off, F, dtype=torch.float32, device=dev) if F > 0 else None
        out.batch = torch.empty(off, dtype=torch.int64, device=dev)
        src = torch.empty(off, dtype=torch.int32, device=dev) if return_source else None
        with torch.cuda.device(dev):
            rc = _lib.lib().b2pn_augment_batch(self.pos.data_ptr(), None if self.x is None else self.x.data_ptr(), F,
                                               arr, B, seed & 0xFFFFFFFFFFFFFFFF, out.pos.data_ptr(),
                                               None if out.x is None else out.x.data_ptr(), out.batch.data_ptr(),
                                               None if src is None else src.data_ptr(),
                                               torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "b2pn_augment_batch")
        ids = [r[0] for r in recs]
        out.y = None if self.y is None else self.y[torch.tensor(ids, device=dev)].reshape(-1)
        ptr = torch.zeros(B + 1, dtype=torch.int64)
        ptr[1:] = torch.cumsum(torch.tensor(sizes, dtype=torch.int64), 0)
        out.ptr = ptr
        out.cloud_sizes = sizes
        out.num_graphs = B
        out.cloud_ids = ids
        if return_source:
            out.source_index = src
        return out
