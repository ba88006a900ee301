"""CPU: host-side logic around the hot path that needs no GPU -- the Sampling container handed from Net.sample to
Net.forward, the augmentation's scalar draws, argument checks of the pipelined step."""
import random

import pytest
import torch

from dl_biomass_b200.pointnet2_regressor import Sampling


def _group(n, k, with_rows):
    nbr = torch.arange(n * k, dtype=torch.int32).reshape(n, k)
    cnt = torch.full((n,), k, dtype=torch.int32)
    rowmap = None
    if with_rows:
        rowmap = (torch.zeros(16, dtype=torch.int32), torch.zeros(128, dtype=torch.int32),
                  torch.zeros(2, dtype=torch.int64), 128, torch.zeros(128, dtype=torch.bfloat16))
    return nbr, cnt, rowmap, (torch.ones(9, 128, dtype=torch.bfloat16) if with_rows else None)


def test_sampling_container_flattens_and_clones_everything():
    l1 = (torch.arange(6), torch.zeros(6, 3), torch.zeros(6, dtype=torch.int64))
    l2 = (torch.arange(2), torch.zeros(2, 3), torch.zeros(2, dtype=torch.int64))
    plain = Sampling([10, 20], l1, l2)
    assert len(plain.tensors()) == 6 and plain.group1 is None
    full = Sampling([10, 20], l1, l2, _group(6, 4, True), _group(2, 4, False))
    ts = full.tensors()
    # 6 sampling tensors + (nbr, cnt, 4 row tensors, l1op) + (nbr, cnt): the integer capacity is not a tensor
    assert len(ts) == 6 + 7 + 2 and all(isinstance(t, torch.Tensor) for t in ts)
    cl = full.clone()
    assert cl.sizes == full.sizes and cl.group1[2][3] == 128 and cl.group2[2] is None and cl.group2[3] is None
    for a, b in zip(ts, cl.tensors()):
        assert a.data_ptr() != b.data_ptr() and torch.equal(a, b)


def test_augmentation_scalar_draws_match_the_oracle_stream():
    """dl_biomass_b200.augment.draw_scalars and oracle.augment_ref.draw_scalars consume a random.Random identically
    (ranges of /root/reference/augmentation.py:55,79,94,97,115)."""
    from dl_biomass_b200 import augment
    from oracle import augment_ref as ar
    ra, rb = random.Random(123), random.Random(123)
    for n in (100, 513, 7168, 16384, 2, 1):
        for _ in range(25):
            assert augment.draw_scalars(ra, n) == ar.draw_scalars(rb, n)
    assert augment.MIN_POINTS == 100


def test_pipelined_step_rejects_bad_arguments_before_touching_the_gpu():
    from dl_biomass_b200.data import Batch, synthetic_clouds
    from dl_biomass_b200.train import PipelinedTrainStep, weighted_mse_loss
    with pytest.raises(ValueError, match="join"):
        PipelinedTrainStep(None, None, None, join="sometimes")
    b = Batch.from_data_list(synthetic_clouds(1, 2, 64))
    with pytest.raises(RuntimeError, match="B200"):
        PipelinedTrainStep(None, None, b)                    # CPU batch
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        weighted_mse_loss(torch.zeros(2, 4), torch.zeros(8))
