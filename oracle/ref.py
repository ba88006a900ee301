"""b2pn CPU oracle (Python side) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; nothing under ``dl_biomass_b200/`` does.

It restates on the CPU what /root/reference/pointnet2_regressor.py:5-58 computes through
torch_geometric / torch_cluster / torch_scatter (un-vendored, unpinned -- SURVEY.md §0.2):

* ``fps_ref`` / ``ball_query_ref``  -> C library ``oracle/b2pn_oracle.c``
* ``MLPRef``                        -> PyG ``MLP`` (A.4): Lin -> BN -> act -> dropout ... plain last
* ``point_conv_ref``                -> ``PointNetConv`` message/aggregate (A.3, A.5)
* ``SAModuleRef`` / ``GlobalSAModuleRef`` / ``NetRef`` -> the reference classes, line by line
* ``weighted_mse`` / ``make_adam``  -> /root/reference/main.py:157-169, :84

PARITY STATUS: FPS is pinned against the reference's own numpy FPS
(/root/reference/downsampling_point_clouds.py:55-92, golden file
tests/golden/fps_reference_numpy.npz).  Everything else is PARITY UNPINNED: the reference has
no tests or golden vectors and its third-party kernels cannot be installed here; the
canonicalisation rules are SURVEY.md A.7.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libb2pn_oracle.so")
_lib = None



def ptr_from_batch(batch: torch.Tensor, num_clouds=None) -> torch.Tensor:
    """``ptr`` from a sorted ``batch`` vector, as PyG's fps wrapper derives it (scatter_add of ones,
    cumsum).  The oracle's own copy: nothing under oracle/ imports the product package."""
    if num_clouds is None:
        num_clouds = int(batch.max().item()) + 1 if batch.numel() else 0
    counts = torch.bincount(batch, minlength=num_clouds)
    ptr = torch.zeros(num_clouds + 1, dtype=torch.int64)
    ptr[1:] = torch.cumsum(counts, 0)
    return ptr


def build_oracle(force: bool = False) -> str:
    src = os.path.join(_HERE, "b2pn_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH)):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build_oracle()
        lib = ctypes.CDLL(_LIB_PATH)
        i64p, f32p, i32p = (ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_float),
                            ctypes.POINTER(ctypes.c_int32))
        lib.oracle_fps_num_samples.restype = ctypes.c_int64
        lib.oracle_fps_num_samples.argtypes = [ctypes.c_int64, ctypes.c_float]
        lib.oracle_fps_f32.restype = ctypes.c_int
        lib.oracle_fps_f32.argtypes = [f32p, i64p, i64p, i64p, ctypes.c_int32, i64p, ctypes.c_int32]
        lib.oracle_fps_f64.restype = ctypes.c_int
        lib.oracle_fps_f64.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.c_int64, ctypes.c_int64,
                                       ctypes.c_int64, i64p]
        lib.oracle_ball_query_f32.restype = ctypes.c_int
        lib.oracle_ball_query_f32.argtypes = [f32p, f32p, i64p, i64p, ctypes.c_int32, ctypes.c_double,
                                              ctypes.c_int32, i32p, i32p, ctypes.c_int32]
        lib.oracle_num_threads.restype = ctypes.c_int
        _lib = lib
    return _lib


def _p(t: torch.Tensor, ctype):
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctype))


def num_threads() -> int:
    return int(_load().oracle_num_threads())


def fps_num_samples(n: int, ratio: float) -> int:
    """m = ceil(float32(n) * float32(ratio))  (SURVEY.md A.1)."""
    return int(_load().oracle_fps_num_samples(int(n), float(np.float32(ratio))))


def sample_ptr(ptr: torch.Tensor, ratio: float) -> torch.Tensor:
    sizes = (ptr[1:] - ptr[:-1]).tolist()
    out = torch.zeros(len(sizes) + 1, dtype=torch.int64)
    out[1:] = torch.cumsum(torch.tensor([fps_num_samples(n, ratio) for n in sizes], dtype=torch.int64), 0)
    return out


def fps_ref(pos: torch.Tensor, ptr: torch.Tensor, ratio: float,
            start: Optional[torch.Tensor] = None, threads: int = 0) -> torch.Tensor:
    """Global int64 sample indices, clouds concatenated, selection order (A.1)."""
    pos = pos.detach().to(torch.float32).contiguous().cpu()
    ptr = ptr.to(torch.int64).contiguous().cpu()
    out_ptr = sample_ptr(ptr, ratio)
    out = torch.empty(int(out_ptr[-1]), dtype=torch.int64)
    st = None if start is None else start.to(torch.int64).contiguous().cpu()
    rc = _load().oracle_fps_f32(_p(pos, ctypes.c_float), _p(ptr, ctypes.c_int64), _p(out_ptr, ctypes.c_int64),
                                None if st is None else _p(st, ctypes.c_int64), ptr.numel() - 1,
                                _p(out, ctypes.c_int64), threads)
    if rc != 0:
        raise RuntimeError(f"oracle_fps_f32 failed rc={rc}")
    return out


def fps_ref_f64(pos: np.ndarray, m: int, start: int = 0) -> np.ndarray:
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    out = np.empty(m, dtype=np.int64)
    rc = _load().oracle_fps_f64(pos.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), pos.shape[0], m, start,
                                out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    if rc != 0:
        raise RuntimeError(f"oracle_fps_f64 failed rc={rc}")
    return out


def ball_query_ref(src: torch.Tensor, qry: torch.Tensor, src_ptr: torch.Tensor, qry_ptr: torch.Tensor,
                   r: float, K: int = 64, threads: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Fixed-width neighbour slots ``nbr [M,K] int32`` (global source index, -1 pad), ``cnt [M]`` (A.2)."""
    src = src.detach().to(torch.float32).contiguous().cpu()
    qry = qry.detach().to(torch.float32).contiguous().cpu()
    sp = src_ptr.to(torch.int64).contiguous().cpu()
    qp = qry_ptr.to(torch.int64).contiguous().cpu()
    M = qry.size(0)
    nbr = torch.empty(M, K, dtype=torch.int32)
    cnt = torch.empty(M, dtype=torch.int32)
    rc = _load().oracle_ball_query_f32(_p(src, ctypes.c_float), _p(qry, ctypes.c_float), _p(sp, ctypes.c_int64),
                                       _p(qp, ctypes.c_int64), sp.numel() - 1, float(r), K,
                                       _p(nbr, ctypes.c_int32), _p(cnt, ctypes.c_int32), threads)
    if rc != 0:
        raise RuntimeError(f"oracle_ball_query_f32 failed rc={rc}")
    return nbr, cnt


def slots_to_edges(nbr: torch.Tensor, cnt: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(row=query, col=source) edge list in canonical order: by query, ascending source index."""
    M, K = nbr.shape
    valid = torch.arange(K)[None, :] < cnt[:, None].to(torch.int64)
    row = torch.arange(M)[:, None].expand(M, K)[valid]
    col = nbr.to(torch.int64)[valid]
    return row, col


# --------------------------------------------------------------------------------------------
#  MLP / PointNetConv / modules
# --------------------------------------------------------------------------------------------
def _resolve_act(act):
    if act is None:
        return None
    if callable(act) and not isinstance(act, str):
        return act
    name = str(act).lower()
    table = {"relu": torch.nn.ReLU, "leakyrelu": torch.nn.LeakyReLU, "leaky_relu": torch.nn.LeakyReLU,
             "elu": torch.nn.ELU}
    if name not in table:
        raise ValueError(f"unsupported activation {act!r}")
    return table[name]()


class _RoundBF16(torch.autograd.Function):
    """Round to the 16-bit operand format in forward, identity in backward (models where the B200 tensor-core mode
    rounds).  ``fmt``: class attribute; the kernels keep forward-domain operands in fp16 (torch.bfloat16 reproduces the
    round-1 behaviour for error-attribution experiments)."""

    fmt = torch.float16

    @staticmethod
    def forward(ctx, x):
        return x.to(_RoundBF16.fmt).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


class MLPRef(torch.nn.Module):
    """PyG ``MLP(channel_list, act=..., dropout=..., batch_norm=True)``.

    ``emulate_bf16 = True`` rounds weights and the post-BatchNorm values to the 16-bit operand format exactly where the
    tensor-core mode does (fp32 accumulation and statistics), so tests can separate kernel bugs from the
    arg-max / ReLU-mask flips that 16-bit rounding legitimately causes in the gradients."""

    emulate_bf16 = False
    # which of the bf16 mode's roundings the emulation applies (error attribution, tools/bf16_attribution.py):
    # W weights, X the MLP's input (a previous level's output arrives as bf16), Z the stored normalised value,
    # A the activation operand of the next GEMM
    emulate_parts = "WXZA"

    def __init__(self, channel_list: Sequence[int], act="relu", dropout: float = 0.0):
        super().__init__()
        self.channel_list = list(channel_list)
        self.act = _resolve_act(act)
        self.dropout = float(dropout)
        self.lins = torch.nn.ModuleList(
            [torch.nn.Linear(a, b) for a, b in zip(channel_list[:-1], channel_list[1:])])
        self.norms = torch.nn.ModuleList([torch.nn.BatchNorm1d(c) for c in channel_list[1:-1]])

    def forward(self, x):
        if self.emulate_bf16:
            parts = self.emulate_parts
            rnd = _RoundBF16.apply
            ident = lambda t: t  # noqa: E731
            q = rnd if "A" in parts else ident
            qz = rnd if "Z" in parts else ident
            qw = rnd if "W" in parts else ident
            if "X" in parts and x.size(1) > 16:   # wide feature maps go in as bf16; raw inputs as hi+lo pairs (~fp32)
                x = torch.cat([rnd(x[:, :-3]), x[:, -3:]], 1)
            x = F.linear(x, qw(self.lins[0].weight), self.lins[0].bias)
            for lin, norm in zip(self.lins[1:], self.norms):
                # the kernels store the NORMALISED value in bf16 and apply gamma/beta afterwards
                xhat = F.batch_norm(x, norm.running_mean, norm.running_var, None, None, self.training, norm.momentum,
                                    norm.eps)
                if self.training and norm.num_batches_tracked is not None:
                    norm.num_batches_tracked += 1
                x = q(qz(xhat) * norm.weight + norm.bias)   # second rounding: the MMA operand itself is bf16
                if self.act is not None:
                    x = self.act(x)
                x = F.linear(x, qw(lin.weight), lin.bias)
            return x
        x = self.lins[0](x)
        for lin, norm in zip(self.lins[1:], self.norms):
            x = norm(x)
            if self.act is not None:
                x = self.act(x)
            x = F.dropout(x, p=self.dropout, training=self.training)
            x = lin(x)
        return x


def segment_max_first(msg: torch.Tensor, seg: torch.Tensor, num_seg: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-(segment, channel) max with the gradient routed to the FIRST arg-max row (A.5)."""
    E, C = msg.shape
    idx = seg[:, None].expand(E, C)
    with torch.no_grad():
        mx = torch.full((num_seg, C), float("-inf"), dtype=msg.dtype).scatter_reduce(
            0, idx, msg, "amax", include_self=True)
        eid = torch.where(msg == mx[seg], torch.arange(E)[:, None].expand(E, C), torch.full((1, 1), E))
        arg = torch.full((num_seg, C), E, dtype=torch.int64).scatter_reduce(0, idx, eid, "amin", include_self=True)
        empty = arg >= E
        arg = arg.clamp(max=max(E - 1, 0))
    out = msg.gather(0, arg)
    out = torch.where(empty, torch.zeros((), dtype=msg.dtype), out)  # segments with no rows -> 0 (A.3)
    return out, arg


def point_conv_ref(nn: torch.nn.Module, x: Optional[torch.Tensor], pos_src: torch.Tensor, pos_dst: torch.Tensor,
                   row: torch.Tensor, col: torch.Tensor) -> torch.Tensor:
    """``PointNetConv(local_nn=nn, add_self_loops=False)``: max_e nn([x_j || pos_j - pos_i]) (A.3)."""
    msg = pos_src[col] - pos_dst[row]
    if x is not None:
        msg = torch.cat([x[col], msg], dim=1)
    msg = nn(msg)
    out, _ = segment_max_first(msg, row, pos_dst.size(0))
    return out


class SAModuleRef(torch.nn.Module):
    """/root/reference/pointnet2_regressor.py:5-20 with the canonical fps/radius of SURVEY.md A.7."""

    def __init__(self, ratio, r, nn, max_num_neighbors: int = 64):
        super().__init__()
        self.ratio, self.r, self.K = ratio, r, max_num_neighbors
        self.conv = torch.nn.Module()
        self.conv.local_nn = nn  # same state_dict keys as PointNetConv: conv.local_nn.*

    def forward(self, x, pos, batch, ptr=None, start=None):
        if ptr is None:
            ptr = ptr_from_batch(batch)
        idx = fps_ref(pos, ptr, self.ratio, start)                              # :13
        qptr = sample_ptr(ptr, self.ratio)
        nbr, cnt = ball_query_ref(pos, pos[idx], ptr, qptr, self.r, self.K)     # :14-15
        row, col = slots_to_edges(nbr, cnt)
        x = point_conv_ref(self.conv.local_nn, x, pos, pos[idx], row, col)      # :16-18
        return x, pos[idx], batch[idx], qptr                                    # :19-20


class GlobalSAModuleRef(torch.nn.Module):
    """/root/reference/pointnet2_regressor.py:23-33."""

    def __init__(self, nn):
        super().__init__()
        self.nn = nn

    def forward(self, x, pos, batch, num_clouds):
        x = self.nn(torch.cat([x, pos], dim=1))                                 # :29
        x, _ = segment_max_first(x, batch, num_clouds)                          # :30 global_max_pool
        pos = pos.new_zeros((x.size(0), 3))
        batch = torch.arange(x.size(0))
        return x, pos, batch


class NetRef(torch.nn.Module):
    """/root/reference/pointnet2_regressor.py:36-58."""

    def __init__(self, num_features, activation_function, neuron_multiplier, dropout_probability):
        super().__init__()
        nm = 1 if neuron_multiplier == 0 else neuron_multiplier
        self.sa1_module = SAModuleRef(0.2, 2, MLPRef([3 + num_features, 64 * nm, 64 * nm, 128 * nm],
                                                     act=activation_function))
        self.sa2_module = SAModuleRef(0.25, 8, MLPRef([128 * nm + 3, 128 * nm, 128 * nm, 256 * nm],
                                                      act=activation_function))
        self.sa3_module = GlobalSAModuleRef(MLPRef([256 * nm + 3, 256 * nm, 512 * nm, 1024 * nm],
                                                   act=activation_function))
        self.mlp = MLPRef([1024 * nm, 128 * nm, 128 * nm, 4], act=None, dropout=dropout_probability)

    def forward(self, data, start=None):
        ptr = getattr(data, "ptr", None)
        if ptr is None:
            ptr = ptr_from_batch(data.batch)
        B = ptr.numel() - 1
        x1, pos1, batch1, ptr1 = self.sa1_module(data.x, data.pos, data.batch, ptr, start)
        x2, pos2, batch2, _ = self.sa2_module(x1, pos1, batch1, ptr1, None)
        x3, _, _ = self.sa3_module(x2, pos2, batch2, B)
        return self.mlp(x3)


LOSS_WEIGHTS = (1.0 / 11.0, 1.0 / 12.0, 1.0 / 5.0, 1.0 / 72.0)


def weighted_mse(outs: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """/root/reference/main.py:157-169: sum_c w_c * mse(y[:,c], outs[:,c])."""
    y = y.reshape(outs.size(0), 4).to(outs.dtype)
    w = torch.tensor(LOSS_WEIGHTS, dtype=outs.dtype, device=outs.device)
    return (((outs - y) ** 2).mean(0) * w).sum()


def make_adam(params, lr: float = 0.00179966410046844, weight_decay: float = 8.0250963438986e-05):
    """/root/reference/main.py:38-39,84."""
    return torch.optim.Adam(params, lr=lr, weight_decay=weight_decay)


def seeded_init_(model: torch.nn.Module, seed: int = 7) -> torch.nn.Module:
    """Deterministic parameter fill (CPU generator) so tests on any box rebuild the same weights.

    Linear: U(-1/sqrt(fan_in), 1/sqrt(fan_in)); BN weight U(0.5,1.5), BN bias U(-0.2,0.2).
    Works on NetRef and on the product Net alike (same parameter names and shapes).
    """
    g = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(model.named_parameters(), key=lambda kv: kv[0]):
            if "norms." in name:
                lo, hi = (0.5, 1.5) if name.endswith("weight") else (-0.2, 0.2)
            else:
                fan_in = p.shape[1] if p.dim() == 2 else p.shape[0]
                bound = 1.0 / float(np.sqrt(fan_in)) if p.dim() == 2 else 0.1
                lo, hi = -bound, bound
            v = torch.rand(p.shape, generator=g, dtype=torch.float32) * (hi - lo) + lo
            p.copy_(v.to(p.device, p.dtype))
    return model
