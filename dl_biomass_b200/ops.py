"""torch.library surface over the C ABI (include/b2pn.h): grouping ops.

``torch.ops.b2pn.fps`` and ``torch.ops.b2pn.ball_query`` stand where
``torch.ops.torch_cluster.fps`` / ``torch.ops.torch_cluster.radius`` stand in the reference
(/root/reference/pointnet2_regressor.py:13-15 via torch_geometric).  They only enqueue kernels
on the current stream of the tensors' device; all sizes come in as host integers, so there is no
device->host synchronisation anywhere in the forward pass.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib

_DEF = torch.library.Library("b2pn", "DEF")
_DEF.define("fps(Tensor pos, Tensor ptr, Tensor out_ptr, Tensor? start, int max_n, int num_out, int seed=0, "
            "Tensor? rng_state=None, int cluster=0, int threads=0) -> (Tensor, Tensor, Tensor)")
_DEF.define("ball_query(Tensor src, Tensor qry, Tensor src_ptr, Tensor qry_ptr, int max_src, int max_qry, "
            "float r, int K, str mode='auto') -> (Tensor, Tensor)")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _require_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("b2pn ops run on a B200 only: got a CPU tensor and there is no CPU fallback")


def _fps_cuda(pos, ptr, out_ptr, start, max_n, num_out, seed=0, rng_state=None, cluster=0, threads=0):
    _require_cuda(pos, ptr, out_ptr, start, rng_state)
    if pos.dtype != torch.float32 or pos.dim() != 2 or pos.size(1) != 3:
        raise ValueError("fps: pos must be [N,3] float32")
    pos = pos.contiguous()
    B = ptr.numel() - 1
    idx = torch.empty(num_out, dtype=torch.int64, device=pos.device)
    pos_out = torch.empty(num_out, 3, dtype=torch.float32, device=pos.device)
    batch_out = torch.empty(num_out, dtype=torch.int64, device=pos.device)
    if rng_state is not None and (rng_state.dtype != torch.int64 or rng_state.numel() < 2):
        raise ValueError("fps: rng_state must be an int64 tensor of 2 elements")
    opts = _lib.FpsOptions(int(cluster), int(threads), int(seed) & 0xffffffffffffffff, _ptr(rng_state))
    with torch.cuda.device(pos.device):
        rc = _lib.lib().b2pn_fps_f32(pos.data_ptr(), ptr.data_ptr(), out_ptr.data_ptr(), _ptr(start), B, max_n,
                                     idx.data_ptr(), pos_out.data_ptr(), batch_out.data_ptr(), ctypes.byref(opts),
                                     _stream(pos))
    _lib.check(rc, "b2pn_fps_f32")
    return idx, pos_out, batch_out


GRID_MIN_SOURCES = 4096   # below this the plain scan with early exit wins (level 2: most sources are neighbours)
GRID_MAX_SOURCES = 32768  # above this the clouds are so dense that K hits come early in the scan (100k-point trees)


def _ball_query_cuda(src, qry, src_ptr, qry_ptr, max_src, max_qry, r, K, mode="auto"):
    """``mode``: "auto" picks the kernel by cloud size, "scan" forces the brute-force kernel (tests compare both);
    a per-call argument -- there is no process-wide switch."""
    _require_cuda(src, qry, src_ptr, qry_ptr)
    if mode not in ("auto", "scan"):
        raise ValueError("mode must be 'auto' or 'scan'")
    if src.dtype != torch.float32 or qry.dtype != torch.float32:
        raise ValueError("ball_query: positions must be float32")
    src, qry = src.contiguous(), qry.contiguous()
    B = src_ptr.numel() - 1
    M = qry.size(0)
    nbr = torch.empty(M, K, dtype=torch.int32, device=src.device)
    cnt = torch.empty(M, dtype=torch.int32, device=src.device)
    lib = _lib.lib()
    if GRID_MIN_SOURCES <= max_src <= GRID_MAX_SOURCES and mode != "scan":
        # large clouds: uniform-grid kernel (same result, two orders of magnitude fewer distance tests)
        ws = torch.empty(int(lib.b2pn_ball_query_workspace_bytes(B, src.size(0))), dtype=torch.uint8, device=src.device)
        with torch.cuda.device(src.device):
            rc = lib.b2pn_ball_query_grid_f32(src.data_ptr(), qry.data_ptr(), src_ptr.data_ptr(), qry_ptr.data_ptr(), B,
                                              src.size(0), max_qry, float(r), K, nbr.data_ptr(), cnt.data_ptr(),
                                              ws.data_ptr(), ws.numel(), _stream(src))
        _lib.check(rc, "b2pn_ball_query_grid_f32")
        return nbr, cnt
    with torch.cuda.device(src.device):
        rc = _lib.lib().b2pn_ball_query_f32(src.data_ptr(), qry.data_ptr(), src_ptr.data_ptr(), qry_ptr.data_ptr(),
                                            B, max_src, max_qry, float(r), K, nbr.data_ptr(), cnt.data_ptr(),
                                            _stream(src))
    _lib.check(rc, "b2pn_ball_query_f32")
    return nbr, cnt


_IMPL = torch.library.Library("b2pn", "IMPL")
_IMPL.impl("fps", _fps_cuda, "CUDA")
_IMPL.impl("ball_query", _ball_query_cuda, "CUDA")


def _no_cpu(*args, **kwargs):
    raise RuntimeError("b2pn ops run on a B200 only: there is no CPU fallback (tensors must be CUDA tensors)")


_IMPL.impl("fps", _no_cpu, "CPU")
_IMPL.impl("ball_query", _no_cpu, "CPU")


@torch.library.register_fake("b2pn::fps")
def _fps_fake(pos, ptr, out_ptr, start, max_n, num_out, seed=0, rng_state=None, cluster=0, threads=0):
    return (pos.new_empty(num_out, dtype=torch.int64), pos.new_empty(num_out, 3),
            pos.new_empty(num_out, dtype=torch.int64))


@torch.library.register_fake("b2pn::ball_query")
def _bq_fake(src, qry, src_ptr, qry_ptr, max_src, max_qry, r, K, mode="auto"):
    return qry.new_empty(qry.size(0), K, dtype=torch.int32), qry.new_empty(qry.size(0), dtype=torch.int32)


# --------------------------------------------------------------------------------------------------
#  host-side layout of the hierarchy (cloud offsets at every level, computed without device reads)
# --------------------------------------------------------------------------------------------------
def fps_num_samples(n: int, ratio: float) -> int:
    """ceil(float32(n)*float32(ratio)) -- the sizing rule of torch_cluster.fps."""
    return int(_lib.lib().b2pn_fps_num_samples(int(n), float(ratio)))


@dataclass
class Level:
    sizes: List[int]          # points per cloud (host)
    ptr: torch.Tensor         # [B+1] int64 on the device
    total: int
    max_n: int
    sizes_f32: Optional[torch.Tensor] = None  # [B] float32 on the device (scales the random FPS start)


_LEVEL_CACHE: dict = {}
_LEVEL_CACHE_MAX = 64


def build_levels(sizes0: Sequence[int], ratios: Sequence[float], device: torch.device) -> List[Level]:
    """Level 0 = input clouds; level i+1 = after fps with ratios[i].  One pinned H2D copy in total, and none at
    all when the same batch layout was seen before (the offsets only depend on the cloud sizes): a training
    loop with fixed-size clouds builds them once, which also keeps the forward pass capturable in a CUDA graph."""
    key = (tuple(int(s) for s in sizes0), tuple(float(r) for r in ratios), str(device))
    hit = _LEVEL_CACHE.get(key)
    if hit is not None:
        return hit
    lv = _build_levels(sizes0, ratios, device)
    if len(_LEVEL_CACHE) >= _LEVEL_CACHE_MAX:
        _LEVEL_CACHE.pop(next(iter(_LEVEL_CACHE)))
    _LEVEL_CACHE[key] = lv
    return lv


def _build_levels(sizes0: Sequence[int], ratios: Sequence[float], device: torch.device) -> List[Level]:
    all_sizes = [list(int(s) for s in sizes0)]
    for r in ratios:
        all_sizes.append([fps_num_samples(n, r) for n in all_sizes[-1]])
    B = len(all_sizes[0])
    host = torch.zeros(len(all_sizes), B + 1, dtype=torch.int64)
    for i, s in enumerate(all_sizes):
        host[i, 1:] = torch.cumsum(torch.tensor(s, dtype=torch.int64), 0)
    hostf = torch.tensor(all_sizes, dtype=torch.float32).reshape(len(all_sizes), B)
    if device.type == "cuda":
        # asynchronous: torch's pinned-memory allocator keeps a staging buffer alive until the copy that reads it has
        # run, so no synchronisation is needed here.  (A host-side wait at this point made every step of a RAGGED
        # training loop -- a new layout per batch -- wait for the GPU to drain before the next launch could be queued.)
        dev = host.pin_memory().to(device, non_blocking=True)
        devf = hostf.pin_memory().to(device, non_blocking=True)
    else:
        dev, devf = host, hostf
    return [Level(s, dev[i], int(host[i, -1]), max(s) if s else 0, devf[i]) for i, s in enumerate(all_sizes)]


def fps(pos: torch.Tensor, src: Level, dst: Level, start: Optional[torch.Tensor] = None, *, seed: int = 0,
        rng_state: Optional[torch.Tensor] = None, cluster: int = 0, threads: int = 0
        ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """idx [M] int64 (global), pos[idx] [M,3], batch[idx] [M] -- pointnet2_regressor.py:13,19.
    ``start``: explicit cloud-local start indices; else, with ``rng_state`` (int64[2] on the device, zeroed once), the
    kernel draws torch_cluster's ``random_start`` itself from (seed, rng_state[0], cloud) and advances the counter; else
    every cloud starts at its point 0.  ``cluster`` / ``threads``: force a kernel variant (benchmark sweeps)."""
    return torch.ops.b2pn.fps(pos, src.ptr, dst.ptr, start, src.max_n, dst.total, int(seed), rng_state, cluster, threads)


def ball_query(src_pos: torch.Tensor, qry_pos: torch.Tensor, src: Level, qry: Level, r: float, K: int = 64,
               mode: str = "auto") -> Tuple[torch.Tensor, torch.Tensor]:
    """Fixed-width neighbour slots nbr [M,K] int32 / cnt [M] -- pointnet2_regressor.py:14-16."""
    return torch.ops.b2pn.ball_query(src_pos, qry_pos, src.ptr, qry.ptr, src.max_n, qry.max_n, float(r), K, mode)
