"""B200-native drop-in for /root/reference/pointnet2_regressor.py.

Same classes, constructor arguments and ``forward`` contracts as the reference module
(``SAModule(ratio, r, nn)``, ``GlobalSAModule(nn)``, ``Net(num_features, activation_function,
neuron_multiplier, dropout_probability)``; ``forward(data)`` over ``data.x / data.pos / data.batch``),
and the same ``state_dict`` keys as the PyG-built reference model (``sa1_module.conv.local_nn.lins.0.weight``,
``...norms.0.running_mean`` ...), so ``main.py`` can ``from pointnet2_regressor import Net`` unchanged.
Underneath, everything PyG / torch_cluster / torch_scatter did is done by libb2pn's sm_100a kernels:

    fps            -> torch.ops.b2pn.fps          (Kernel 1, csrc/fps.cu)
    radius         -> torch.ops.b2pn.ball_query   (Kernel 2, csrc/ball_query.cu)
    PointConv(nn)  -> sa.sa_apply                 (Kernel 3, csrc/sa_*.cu: gather + concat + MLP + max)
    nn(cat) + global_max_pool -> sa.sa_apply in CLOUDS mode

There is no CPU path: tensors must live on a B200.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn.functional as F

from . import head, ops, sa

_PRECISIONS = {"fp32": sa.PREC_F32, "f32": sa.PREC_F32, "float32": sa.PREC_F32,
               "bf16": sa.PREC_BF16, "bfloat16": sa.PREC_BF16}


class MLP(torch.nn.Module):
    """Parameter container shaped like ``torch_geometric.nn.MLP(channel_list, act=..., dropout=...)``
    (SURVEY.md A.4): ``lins[i]`` Linear, ``norms[i]`` BatchNorm1d after every hidden layer, last layer plain.
    ``forward`` is the plain layer-by-layer evaluation (used for the tiny regression head); inside the
    set-abstraction modules the parameters are consumed by the fused kernels instead."""

    def __init__(self, channel_list: Sequence[int], act="relu", dropout: float = 0.0, batch_norm: bool = True,
                 bias: bool = True):
        super().__init__()
        if not batch_norm or not bias:
            raise NotImplementedError("the reference only builds MLPs with batch_norm=True, bias=True")
        self.channel_list = list(int(c) for c in channel_list)
        self.act_name = None if act is None else (act if isinstance(act, str) else type(act).__name__)
        sa.act_code(self.act_name)  # fail early on an activation the kernels do not implement
        self.dropout = float(dropout)
        self.lins = torch.nn.ModuleList(
            [torch.nn.Linear(a, b) for a, b in zip(self.channel_list[:-1], self.channel_list[1:])])
        self.norms = torch.nn.ModuleList([torch.nn.BatchNorm1d(c) for c in self.channel_list[1:-1]])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.lins[0](x)
        for lin, norm in zip(self.lins[1:], self.norms):
            x = norm(x)
            if self.act_name is not None:
                x = F.relu(x)
            x = F.dropout(x, p=self.dropout, training=self.training)
            x = lin(x)
        return x

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # PyG >= 2.1 wraps BatchNorm1d as ``norms.i.module.*``; accept both spellings
        for k in list(state_dict.keys()):
            if k.startswith(prefix + "norms.") and ".module." in k[len(prefix):]:
                state_dict[k.replace(".module.", ".", 1)] = state_dict.pop(k)
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)


class PointConv(torch.nn.Module):
    """Holds ``local_nn`` under the same attribute name as PyG's PointNetConv (state_dict parity)."""

    def __init__(self, local_nn: MLP, add_self_loops: bool = False):
        super().__init__()
        if add_self_loops:
            raise NotImplementedError("the reference uses add_self_loops=False")
        self.local_nn = local_nn


def _cloud_sizes(batch: torch.Tensor, ptr=None) -> List[int]:
    """Points per cloud as host integers.  Free when a host-side ``ptr`` is available (what
    ``Batch.from_data_list`` builds on the CPU); otherwise one small device->host read, like the
    ``int(batch.max())`` inside the reference's fps/radius wrappers."""
    if ptr is not None:
        p = ptr if not ptr.is_cuda else ptr.cpu()
        return (p[1:] - p[:-1]).tolist()
    if batch.numel() == 0:
        return []
    return torch.bincount(batch).cpu().tolist()


class SAModule(torch.nn.Module):
    """/root/reference/pointnet2_regressor.py:5-20."""

    def __init__(self, ratio, r, nn, max_num_neighbors: int = 64):
        super().__init__()
        self.ratio = ratio
        self.r = r
        self.max_num_neighbors = max_num_neighbors
        self.conv = PointConv(nn, add_self_loops=False)
        self.precision = "fp32"
        self.random_start = True  # torch_cluster.fps default
        # random_start is drawn INSIDE the sampling kernel from (seed, call counter, cloud): the counter is a device
        # scalar the kernel advances, so CUDA-graph replays draw fresh start points and no ATen kernel is involved.
        # Plain attributes, not buffers: the module's state_dict stays exactly the reference's.
        self._fps_seed = int(torch.initial_seed() & 0x7fffffffffffffff) ^ (hash((float(ratio), float(r))) & 0xffffffff)
        self._fps_rng_state = None

    def _rng_state(self, device):
        st = self._fps_rng_state
        if st is None or st.device != device:
            st = torch.zeros(2, dtype=torch.int64, device=device)
            self._fps_rng_state = st
        return st

    def _sample(self, pos, src: ops.Level, dst: ops.Level, start=None):
        """fps + pos[idx] + batch[idx] (:13, :19): depends on the positions only, never on the weights."""
        if start is None and self.random_start and src.total > 0:
            return ops.fps(pos, src, dst, None, seed=self._fps_seed, rng_state=self._rng_state(pos.device))
        return ops.fps(pos, src, dst, start)

    def _group(self, pos, pos_dst, src: ops.Level, dst: ops.Level, x=None, gather: bool = False):
        """radius (:14-16) and, for the tensor-core path, the compaction of the filled slots into rows and
        (``gather``: when ``x`` is raw input data, not an activation) the gathered message inputs of :17-18:
        no weights involved, like ``_sample``.  Returns (nbr, cnt, rowmap or None, l1op or None)."""
        nbr, cnt = ops.ball_query(pos, pos_dst, src, dst, self.r, self.max_num_neighbors)
        rowmap = l1op = None
        if _PRECISIONS[self.precision] == sa.PREC_BF16:
            rowmap = sa.pack_rows(nbr, cnt, self.max_num_neighbors)
            if gather and dst.total > 0 and (x is None or not x.requires_grad):
                l1op = sa.gather_rows(x, pos, pos_dst, self.max_num_neighbors, rowmap)
        return nbr, cnt, rowmap, l1op

    def _run(self, x, pos, src: ops.Level, dst: ops.Level, start=None, sampled=None, grouped=None, x_bf16=None):
        """Returns (out, pos[idx], batch[idx], idx, bf16 copy of out or None)."""
        idx, pos_dst, batch_dst = sampled if sampled is not None else self._sample(pos, src, dst, start)
        nbr, cnt, rowmap, l1op = grouped if grouped is not None else self._group(pos, pos_dst, src, dst)
        out, _, out16 = sa.sa_apply(self.conv.local_nn, x, pos, pos_dst, nbr, cnt, None, seg_mode=sa.SEG_SLOTS,
                                    K=self.max_num_neighbors, n_dst=dst.total, precision=_PRECISIONS[self.precision],
                                    rowmap=rowmap, l1op=l1op, x_bf16=x_bf16, want_bf16_out=True)   # :18
        return out, pos_dst, batch_dst, idx, out16

    def forward(self, x, pos, batch):
        lv = ops.build_levels(_cloud_sizes(batch), [self.ratio], pos.device)
        out, pos_dst, batch_dst, _, _ = self._run(x, pos, lv[0], lv[1])
        return out, pos_dst, batch_dst


class GlobalSAModule(torch.nn.Module):
    """/root/reference/pointnet2_regressor.py:23-33."""

    def __init__(self, nn):
        super().__init__()
        self.nn = nn
        self.precision = "fp32"

    def _run(self, x, pos, batch, num_clouds: int, x_bf16=None):
        out, _ = sa.sa_apply(self.nn, x, pos, None, None, None, batch, seg_mode=sa.SEG_CLOUDS, K=0,
                             n_dst=num_clouds, precision=_PRECISIONS[self.precision], x_bf16=x_bf16)   # :29-30
        return out

    def forward(self, x, pos, batch):
        num_clouds = int(batch.max().item()) + 1 if batch.numel() else 0
        x = self._run(x, pos, batch, num_clouds)
        pos = pos.new_zeros((x.size(0), 3))                                              # :31
        batch = torch.arange(x.size(0), device=batch.device)                             # :32
        return x, pos, batch


class Sampling:
    """Everything about one batch that depends on the point positions only (``Net.sample``): per set-abstraction
    level the farthest-point samples (idx, pos, batch) and, optionally, the neighbourhoods (nbr, cnt, compacted
    rows, gathered level-1 message inputs).  A training loop can compute them for the NEXT batch on a second stream
    while the current batch trains (``train.PipelinedTrainStep``)."""

    def __init__(self, sizes, level1, level2, group1=None, group2=None):
        self.sizes, self.level1, self.level2 = tuple(sizes), tuple(level1), tuple(level2)
        self.group1, self.group2 = group1, group2

    @staticmethod
    def _flat(group):
        if group is None:
            return []
        nbr, cnt, rowmap, l1op = group
        return ([nbr, cnt] + ([] if rowmap is None else [t for t in rowmap if isinstance(t, torch.Tensor)])
                + ([] if l1op is None else [l1op]))

    def tensors(self):
        return list(self.level1) + list(self.level2) + self._flat(self.group1) + self._flat(self.group2)

    def clone(self) -> "Sampling":
        def cg(group):
            if group is None:
                return None
            nbr, cnt, rowmap, l1op = group
            rm = None if rowmap is None else tuple(t.clone() if isinstance(t, torch.Tensor) else t for t in rowmap)
            return nbr.clone(), cnt.clone(), rm, None if l1op is None else l1op.clone()
        return Sampling(self.sizes, [t.clone() for t in self.level1], [t.clone() for t in self.level2],
                        cg(self.group1), cg(self.group2))


class _CallInBackward(torch.autograd.Function):
    """Identity whose backward first runs ``fn()``: a hook at a fixed point of the backward pass."""

    @staticmethod
    def forward(ctx, x, fn):
        ctx.fn = fn
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        ctx.fn()
        return g, None


class Net(torch.nn.Module):
    """/root/reference/pointnet2_regressor.py:36-58.

    Extra, optional and keyword-only: ``precision`` ("fp32": CUDA-core FMA, the 1e-4 parity mode; "bf16":
    tcgen05 tensor-core tiles with fp32 accumulation).  It can also be switched later with
    ``set_precision``."""

    def __init__(self, num_features, activation_function, neuron_multiplier, dropout_probability, *,
                 precision: str = "fp32"):
        super().__init__()
        if neuron_multiplier == 0:
            neuron_multiplier = 1  # :40-43
        nm = neuron_multiplier
        self.sa1_module = SAModule(0.2, 2, MLP([3 + num_features, 64 * nm, 64 * nm, 128 * nm], act=activation_function))
        self.sa2_module = SAModule(0.25, 8, MLP([128 * nm + 3, 128 * nm, 128 * nm, 256 * nm], act=activation_function))
        self.sa3_module = GlobalSAModule(MLP([256 * nm + 3, 256 * nm, 512 * nm, 1024 * nm], act=activation_function))
        self.mlp = MLP([1024 * nm, 128 * nm, 128 * nm, 4], act=None, dropout=dropout_probability)
        # dropout noise of the fused head: (seed, counter, layer, element); the counter is a device scalar bumped by the
        # forward kernel, so a CUDA-graph replay draws fresh noise every step.  (A plain attribute, not a buffer: the
        # module's buffers stay exactly the reference's.)
        self._head_rng_counter = None
        self._head_seed = int(torch.initial_seed() & 0x7fffffffffffffff)
        self.set_precision(precision)

    def __getstate__(self):
        # torch.save(model) of /root/reference/main.py:245 pickles the module: the optimiser's arena is not part of it
        st = self.__dict__.copy()
        st.pop("_b2pn_arena", None)
        return st

    def set_precision(self, precision: str) -> "Net":
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
        for m in (self.sa1_module, self.sa2_module, self.sa3_module):
            m.precision = precision
        self.precision = precision
        return self

    def set_random_start(self, flag: bool) -> "Net":
        self.sa1_module.random_start = self.sa2_module.random_start = bool(flag)
        return self

    def _levels(self, data):
        pos = data.pos
        if not pos.is_cuda:
            raise RuntimeError("dl_biomass_b200.Net runs on a B200 only: move the batch to the GPU "
                               "(there is no CPU fallback)")
        sizes = getattr(data, "cloud_sizes", None)
        if sizes is None:
            sizes = _cloud_sizes(data.batch, getattr(data, "ptr", None))
        return sizes, ops.build_levels(sizes, [self.sa1_module.ratio, self.sa2_module.ratio], pos.device)

    def _usable(self, group):
        """A precomputed neighbourhood is reused only if it carries what the current precision needs."""
        if group is None or ((group[2] is None) != (_PRECISIONS[self.precision] != sa.PREC_BF16)):
            return None
        return group

    def sample(self, data, start: Optional[torch.Tensor] = None, grouping: bool = True, aux_stream=None) -> Sampling:
        """Farthest-point sampling and (``grouping``) ball query + row compaction of both levels for ``data`` (no
        weights involved); pass the result to ``forward(data, sampling=...)``.  ``aux_stream``: a second CUDA stream
        on which the level-1 grouping runs while the level-2 sampling (one CTA per cloud) occupies the current one."""
        sizes, lv = self._levels(data)
        pos = data.pos.to(torch.float32)
        l1 = self.sa1_module._sample(pos, lv[0], lv[1], start)
        g1 = g2 = None
        if grouping and aux_stream is not None:
            cur = torch.cuda.current_stream(pos.device)
            aux_stream.wait_stream(cur)
            with torch.cuda.stream(aux_stream):
                g1 = self.sa1_module._group(pos, l1[1], lv[0], lv[1], data.x, gather=True)
                for t in Sampling._flat(g1):
                    t.record_stream(cur)
        l2 = self.sa2_module._sample(l1[1], lv[1], lv[2])
        if grouping:
            if g1 is None:
                g1 = self.sa1_module._group(pos, l1[1], lv[0], lv[1], data.x, gather=True)
            g2 = self.sa2_module._group(l1[1], l2[1], lv[1], lv[2])
            if aux_stream is not None:
                cur.wait_stream(aux_stream)
        return Sampling(sizes, l1, l2, g1, g2)

    def forward(self, data, start: Optional[torch.Tensor] = None, sampling: Optional[Sampling] = None,
                after_grouping=None, before_level1_backward=None):
        """``after_grouping``: optional callable invoked when the last kernel of the two set-abstraction levels'
        FORWARD has been enqueued and again never before their BACKWARD starts: with autograd recording it fires in
        the backward pass right before the level-2 backward (after the global level and the head went both ways),
        otherwise right after the level-2 forward.  train.PipelinedTrainStep joins its sampling stream there.
        ``before_level1_backward``: optional callable fired in the backward pass between the level-2 and the level-1
        backward (only with autograd recording)."""
        x, pos, batch = data.x, data.pos, data.batch                                     # :53
        sizes, lv = self._levels(data)
        if sampling is not None and tuple(sampling.sizes) != tuple(sizes):
            raise ValueError("sampling was computed for a batch with different cloud sizes")
        pos = pos.to(torch.float32)
        s1 = None if sampling is None else sampling.level1
        s2 = None if sampling is None else sampling.level2
        g1 = None if sampling is None else self._usable(sampling.group1)
        g2 = None if sampling is None else self._usable(sampling.group2)
        x1, pos1, _, _, x1h = self.sa1_module._run(x, pos, lv[0], lv[1], start, s1, g1)  # :54
        if before_level1_backward is not None and torch.is_grad_enabled() and x1.requires_grad:
            x1 = _CallInBackward.apply(x1, before_level1_backward)
        x2, pos2, batch2, _, x2h = self.sa2_module._run(x1, pos1, lv[1], lv[2], None, s2, g2, x_bf16=x1h)  # :55
        if after_grouping is not None:
            if torch.is_grad_enabled() and x2.requires_grad:
                x2 = _CallInBackward.apply(x2, after_grouping)
            else:
                after_grouping()
        x3 = self.sa3_module._run(x2, pos2, batch2, len(sizes), x_bf16=x2h)              # :56
        if head.supported(self.mlp, x3):                                                 # :58
            if self._head_rng_counter is None or self._head_rng_counter.device != x3.device:
                self._head_rng_counter = torch.zeros((), dtype=torch.int64, device=x3.device)
            return head.head_apply(self.mlp, x3, self._head_rng_counter, self._head_seed)
        return self.mlp(x3)   # shapes outside the fused head's range (batch > 32 clouds, wide heads): ATen on the GPU
