"""Offline resampler: the B200 drop-in for the reference's numpy ``farthest_point_sampling(coords, k)``
(/root/reference/downsampling_point_clouds.py:55-92), used there to cut every raw lidar plot down to the 7 168 points
the network trains on (:153).  Same name, same arguments, same result bit for bit (float64 arithmetic on the raw,
un-centred coordinates); ``farthest_point_sampling_batch`` resamples many plots in one launch -- one CTA per plot, which is
where the parallelism of a dataset of thousands of plots is."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from . import _lib


def farthest_point_sampling_batch(clouds: Sequence, k: int, device="cuda") -> List[np.ndarray]:
    """``[farthest_point_sampling(c, k) for c in clouds]`` in one kernel launch.  Every cloud needs >= k points
    (the reference samples with replacement below that, downsampling_point_clouds.py:155-156)."""
    lib = _lib.lib()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("b2pn resampler runs on a B200 only: there is no CPU fallback")
    arrs = [np.ascontiguousarray(np.asarray(c, dtype=np.float64)[:, :3]) for c in clouds]
    sizes = [a.shape[0] for a in arrs]
    if any(n < k for n in sizes):
        raise ValueError("farthest_point_sampling needs at least k points per cloud")
    if not arrs:
        return []
    ptr = torch.zeros(len(arrs) + 1, dtype=torch.int64)
    ptr[1:] = torch.cumsum(torch.tensor(sizes, dtype=torch.int64), 0)
    out_ptr = torch.arange(len(arrs) + 1, dtype=torch.int64) * int(k)
    pos = torch.from_numpy(np.concatenate(arrs, 0)).to(dev)
    ptr_d, out_ptr_d = ptr.to(dev), out_ptr.to(dev)
    out = torch.empty(len(arrs) * int(k), dtype=torch.int64, device=dev)
    dist = torch.empty(pos.shape[0], dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.b2pn_fps_f64(pos.data_ptr(), ptr_d.data_ptr(), out_ptr_d.data_ptr(), None, len(arrs), out.data_ptr(),
                              dist.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "b2pn_fps_f64")
    idx = out.cpu().numpy().reshape(len(arrs), int(k))
    return [idx[i] - int(ptr[i]) for i in range(len(arrs))]


def farthest_point_sampling(coords, k: int) -> np.ndarray:
    """Indices of ``k`` farthest-point samples of ``coords`` ([N, >=3] array-like, float64), starting at point 0 --
    the signature and result of the reference function."""
    return farthest_point_sampling_batch([coords], k)[0]
