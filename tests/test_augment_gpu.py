"""GPU: libb2pn's on-device augmentation (csrc/augment.cu) against oracle/augment_ref.py, through the C-ABI."""
import random

import numpy as np
import pytest
import torch

from dl_biomass_b200.augment import CloudCache, MIN_POINTS
from dl_biomass_b200.data import Data, synthetic_clouds
from oracle import augment_ref as ar

pytestmark = pytest.mark.gpu


def _clouds(sizes, F, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for n in sizes:
        pos = torch.randn(n, 3, generator=g) * torch.tensor([4.0, 4.0, 8.0])
        x = torch.rand(n, F, generator=g) * 20 if F > 0 else None
        out.append(Data(x=x, pos=pos, y=torch.rand(4, generator=g)))
    return out


@pytest.mark.parametrize("F", [1, 0, 2])
def test_augmented_batch_matches_oracle(cuda_device, F):
    sizes = [100, 513, 7168, 16384, 2, 1000]
    clouds = _clouds(sizes, F, seed=F)
    cache = CloudCache(clouds, cuda_device)
    rng = random.Random(21)
    plan = cache.plan(range(len(sizes)), rng, epoch=3)
    assert [r[0] for r in plan] == [0, 1, 2, 3, 5] or len(plan) >= 4    # the 2-point cloud is dropped (< MIN_POINTS)
    assert all(r[2] + r[3] >= MIN_POINTS for r in plan)
    b = cache.batch(None, seed=77, plan=plan, return_source=True)
    assert b.cloud_sizes == [r[2] + r[3] for r in plan] and b.ptr.tolist()[-1] == b.pos.shape[0]
    pos, src, bat = b.pos.cpu().numpy(), b.source_index.cpu().numpy(), b.batch.cpu().numpy()
    xs = None if F == 0 else b.x.cpu().numpy()
    for i, (cid, n, n_keep, n_dup, sd, angle, uid) in enumerate(plan):
        lo, hi = int(b.ptr[i]), int(b.ptr[i + 1])
        want_pos, want_x, want_src = ar.augment_cloud_ref(clouds[cid].pos.numpy(), None if F == 0 else clouds[cid].x.numpy(),
                                                          n_keep, n_dup, np.float32(sd).item(), angle, 77, uid)
        assert np.array_equal(src[lo:hi], want_src)                   # selection and order: bit-exact
        assert np.all(bat[lo:hi] == i)
        # fp32 rotation / Box-Muller on the device vs float64 in the oracle: 2e-5 m on coordinates of up to ~40 m
        assert np.abs(pos[lo:hi] - want_pos).max() <= 2e-5, (i, np.abs(pos[lo:hi] - want_pos).max())
        if F:
            assert np.abs(xs[lo:hi] - want_x).max() <= 2e-5
        assert torch.equal(b.y.reshape(-1, 4)[i].cpu(), clouds[cid].y)
    assert b.cloud_ids == [r[0] for r in plan]


def test_augmentation_is_a_pure_function_of_seed_and_epoch(cuda_device):
    cache = CloudCache(synthetic_clouds(50, 6, 2048), cuda_device)
    ids = [4, 0, 5, 2]
    a = cache.batch(ids, random.Random(1), seed=9, epoch=0)
    b = cache.batch(ids, random.Random(1), seed=9, epoch=0)
    c = cache.batch(ids, random.Random(1), seed=9, epoch=1)
    assert torch.equal(a.pos, b.pos) and torch.equal(a.x, b.x) and torch.equal(a.batch, b.batch)
    assert a.pos.shape == c.pos.shape and not torch.equal(a.pos, c.pos)
    # un-augmented assembly: the same points, shuffled
    d = cache.batch(ids, augment=False, seed=9, return_source=True)
    for i, cid in enumerate(ids):
        lo, hi = int(d.ptr[i]), int(d.ptr[i + 1])
        assert hi - lo == cache.sizes[cid]
        src = d.source_index[lo:hi].long()
        assert torch.equal(torch.sort(src).values, torch.arange(hi - lo, device=src.device))
        ref = cache.pos[cache.offsets[cid]:cache.offsets[cid + 1]][src]
        assert torch.equal(d.pos[lo:hi], ref)


def test_many_clouds_and_training_on_augmented_batches(cuda_device):
    """More clouds than one launch takes (64), and the batches feed the model's ragged path."""
    from dl_biomass_b200.pointnet2_regressor import Net
    from dl_biomass_b200.train import make_optimizer, train_step
    cache = CloudCache(synthetic_clouds(70, 80, 256), cuda_device)
    big = cache.batch(list(range(80)), random.Random(2), seed=3)
    assert big.num_graphs == 80 and int(big.batch.max()) == 79
    cnt = torch.bincount(big.batch).cpu().tolist()
    assert cnt == big.cloud_sizes
    torch.manual_seed(0)
    net = Net(1, "ReLU", 0, 0.0, precision="bf16").to(cuda_device)
    net.train()
    opt = make_optimizer(net.parameters())
    rng = random.Random(5)
    losses = []
    for ep in range(3):
        bt = cache.batch([1, 7, 9, 30], rng, seed=13, epoch=ep)
        losses.append(float(train_step(net, opt, bt)))
    assert all(np.isfinite(losses))


def test_copies_of_a_cloud_in_one_epoch_are_augmented_independently(cuda_device):
    """The reference's training set holds every cloud 1 + num_augs times per epoch, each copy augmented on its own
    (/root/reference/main.py:100-112): two records of the SAME cloud in the same epoch must get different per-point
    streams (different permutations / duplicated points), not just different scalar draws."""
    clouds = _clouds([800], 1, seed=5)
    cache = CloudCache(clouds, cuda_device)
    rng = random.Random(3)
    plan = cache.plan([0, 0, 0], rng, epoch=1)
    assert len({r[6] for r in plan}) == 3                     # three distinct generator streams
    b = cache.batch(None, seed=9, plan=plan, return_source=True)
    src = b.source_index.cpu().numpy()
    off = np.cumsum([0] + b.cloud_sizes)
    firsts = [tuple(src[off[i]:off[i] + 32].tolist()) for i in range(3)]
    assert len(set(firsts)) == 3                              # different point orders, not nested prefixes of one
