"""CPU: the augmentation oracle against the reference's own functions (golden fixture), the counter-based generator
of libb2pn against its numpy restatement, and host-side argument checks of b2pn_augment_batch."""
import ctypes
import os
import random

import numpy as np
import pytest

from oracle import augment_ref as ar

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "augment_reference.npz")


@pytest.fixture(scope="module")
def built_lib():
    from dl_biomass_b200 import _lib
    _lib.build()
    return _lib


def test_oracle_transformation_matches_reference_functions():
    """tests/golden/augment_reference.npz holds outputs of /root/reference/augmentation.py's point_removal,
    random_noise and rotate_points themselves (oracle/gen_golden_augment.py) together with the draws they consumed:
    given those draws the restatement must reproduce them to the last bit."""
    g = np.load(GOLD)
    assert int(g["num_cases"]) >= 5
    for i in range(int(g["num_cases"])):
        c, x = ar.apply_augmentation(g[f"c{i}_coords"], g[f"c{i}_x"], g[f"c{i}_keep"], g[f"c{i}_noise_c"],
                                     g[f"c{i}_noise_x"], bool(g[f"c{i}_add"]), g[f"c{i}_use"], float(g[f"c{i}_angle"]))
        assert c.shape == g[f"c{i}_out_coords"].shape and x.shape == g[f"c{i}_out_x"].shape
        assert np.array_equal(c, g[f"c{i}_out_coords"]) and np.array_equal(x, g[f"c{i}_out_x"])
        n, k = g[f"c{i}_coords"].shape[0], g[f"c{i}_keep"].shape[0]
        assert round(n * 0.9) <= k <= n and g[f"c{i}_use"].shape[0] <= round(k * 0.1)


def test_counter_generator_matches_library(built_lib):
    h = built_lib.lib()
    rng = random.Random(3)
    for _ in range(200):
        seed, uid = rng.getrandbits(64), rng.getrandbits(40)
        stream, ctr = rng.randrange(3), rng.getrandbits(rng.choice((8, 20, 45)))
        assert int(h.b2pn_augment_draw(seed, uid, stream, ctr)) == int(ar.draw(seed, uid, stream, ctr))


def test_counter_draws_have_the_reference_distributions():
    rng = random.Random(11)
    for n in (100, 1000, 7168):
        for _ in range(20):
            n_keep, n_dup, sd, angle = ar.draw_scalars(rng, n)
            assert round(n * 0.9) <= n_keep <= n and 0 <= n_dup <= round(n_keep * 0.1)
            assert 0.01 <= abs(sd) <= 0.025 and -180.0 <= angle <= 180.0
    keep, use, zc, zx = ar.counter_draws(7168, 6800, 500, 1, seed=5, uid=9)
    assert len(set(keep.tolist())) == 6800 and keep.min() >= 0 and keep.max() < 7168
    assert len(set(use.tolist())) == 500 and use.max() < 6800
    assert not np.array_equal(keep, np.sort(keep))                     # random ORDER, not just a random subset
    z = np.concatenate([zc.ravel(), zx.ravel()])
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1.0) < 0.02 and abs((z ** 3).mean()) < 0.05
    # every point is kept equally often: 200 samples of a 90 % subset
    cnt = np.zeros(500)
    for uid in range(200):
        k, _, _, _ = ar.counter_draws(500, 450, 0, 0, seed=1, uid=uid)
        cnt[k] += 1
    assert abs(cnt.mean() - 180.0) < 1e-9 and cnt.min() > 150 and cnt.max() <= 200
    first = np.array([ar.counter_draws(500, 450, 0, 0, seed=1, uid=u)[0][0] for u in range(400)])
    assert len(set(first.tolist())) > 250                              # the first kept point varies


def test_augment_batch_argument_checks_need_no_gpu(built_lib):
    h = built_lib.lib()
    assert h.b2pn_augment_max_points() == 16384
    one = (built_lib.AugmentCloud * 1)()
    dummy = ctypes.c_void_p(64)
    one[0].n_src, one[0].n_keep, one[0].n_dup = 1000, 1001, 0
    assert h.b2pn_augment_batch(dummy, None, 0, one, 1, 1, dummy, None, None, None, None) == -1
    one[0].n_keep, one[0].n_dup = 900, 901
    assert h.b2pn_augment_batch(dummy, None, 0, one, 1, 1, dummy, None, None, None, None) == -1
    one[0].n_src, one[0].n_keep, one[0].n_dup = 20000, 19000, 0
    assert h.b2pn_augment_batch(dummy, None, 0, one, 1, 1, dummy, None, None, None, None) == -2
    assert h.b2pn_augment_batch(None, None, 0, one, 1, 1, dummy, None, None, None, None) == -1
    assert h.b2pn_augment_batch(dummy, None, 1, one, 1, 1, dummy, None, None, None, None) == -1   # F > 0 needs x
    assert h.b2pn_augment_batch(None, None, 0, None, 0, 1, None, None, None, None, None) == 0


def test_cloud_cache_fails_loudly_without_gpu():
    import torch
    from dl_biomass_b200.augment import CloudCache
    from dl_biomass_b200.data import synthetic_clouds
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        CloudCache(synthetic_clouds(1, 2, 128), torch.device("cpu"))
