"""CPU restatement of the reference's training augmentation -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product
(dl_biomass_b200/augment.py -> libb2pn's augment.cu) never does.

Follows /root/reference/augmentation.py:
    point_removal  :73-89    random subset of round(0.9 n)..n points, in random order
    random_noise   :92-122   coords/x +- N(0, sd) with sd ~ U(0.01, 0.025), then 0..round(0.1 n') distinct jittered
                             points are appended after the (un-jittered) input
    rotate_points  :54-70    coords @ [[c,-s,0],[s,c,0],[0,0,1]], angle ~ U(-180, 180) degrees
in the order AugmentPointCloudsInFiles.__getitem__ applies them (:287-289).

PARITY: ``apply_augmentation`` -- the transformation GIVEN the random draws -- is pinned against the reference's own
three functions executed in place (tests/golden/augment_reference.npz, written by oracle/gen_golden_augment.py).
The DRAWS themselves cannot be pinned: the reference pulls them from the global ``random`` / ``numpy.random``
Mersenne twisters in call order; the device path uses the counter-based generator restated below (``draw``), which
produces the same distributions (uniform random subset in uniformly random order, iid normal deviates).
"""
from __future__ import annotations

import math

import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(z: np.ndarray) -> np.ndarray:
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z ^ (z >> np.uint64(30))
        z = z * np.uint64(0xBF58476D1CE4E5B9)
        z = z ^ (z >> np.uint64(27))
        z = z * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def draw(seed: int, uid: int, stream: int, counter) -> np.ndarray:
    """64-bit draw number ``counter`` of generator stream (seed, uid, stream): csrc/augment.cu aug_draw(aug_stream())."""
    with np.errstate(over="ignore"):
        a = _mix(np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * np.uint64(uid + 1))
        b = _mix(a + np.uint64(0xD1B54A32D192ED03) * np.uint64(stream + 1))
        return _mix(b + np.uint64(0x9E3779B97F4A7C15) * (np.asarray(counter, dtype=np.uint64) + np.uint64(1)))


def normal_from_draw(r: np.ndarray) -> np.ndarray:
    """Box-Muller on bits 63..40 (u1 in (0,1]) and 39..16 (u2 in [0,1)); float64 here, float32 on the device."""
    u1 = ((r >> np.uint64(40)).astype(np.float64) + 1.0) * 2.0 ** -24
    u2 = ((r >> np.uint64(16)) & np.uint64(0xFFFFFF)).astype(np.float64) * 2.0 ** -24
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def apply_augmentation(coords, x, keep_idx, noise_c, noise_x, add: bool, use_idx, angle_deg: float):
    """The three reference steps with every random draw passed in.

    keep_idx  [n_keep]      indices kept by point_removal, in output order              (augmentation.py:75-82)
    noise_c   [n_keep, 3]   deviates for the coordinates, noise_x [n_keep, dim] for x   (:98-111)
    add       True: deviates are added, False: subtracted                               (:97, :105)
    use_idx   [n_dup]       jittered points that get appended                           (:114-120)
    angle_deg rotation about z                                                          (:55-69)
    """
    coords = np.asarray(coords, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    c1, x1 = coords[keep_idx, :], x[keep_idx, :]                       # point_removal
    sgn = 1.0 if add else -1.0
    jc, jx = c1 + sgn * noise_c, x1 + sgn * noise_x                     # random_noise
    c2 = np.append(c1, jc[use_idx, :], axis=0)
    x2 = np.append(x1, jx[use_idx, :], axis=0)
    a = np.radians(angle_deg)                                           # rotate_points
    rot = np.array([[np.cos(a), -np.sin(a), 0.0], [np.sin(a), np.cos(a), 0.0], [0.0, 0.0, 1.0]])
    c3 = c2.copy()
    c3[:, :3] = np.matmul(c2[:, :3], rot)
    return c3, x2


def draw_scalars(rng, n: int):
    """The per-sample scalar draws with the reference's ranges (``rng``: a ``random.Random``):
    n_keep = randint(round(0.9 n), n) (:79); sd ~ U(0.01, 0.025) (:94); add iff U(0,1) >= 0.5 (:97);
    n_dup = randint(0, round(0.1 n_keep)) (:115); angle ~ U(-180, 180) (:55)."""
    n_keep = rng.randint(round(n * 0.9), n)
    sd = rng.uniform(0.01, 0.025)
    add = rng.uniform(0.0, 1.0) >= 0.5
    n_dup = rng.randint(0, round(n_keep * 0.1))
    angle = rng.uniform(-180.0, 180.0)
    return n_keep, n_dup, (sd if add else -sd), angle


def counter_draws(n: int, n_keep: int, n_dup: int, dim: int, seed: int, uid: int):
    """keep_idx, use_idx and the unit normal deviates the device generates for (seed, uid)."""
    k0 = draw(seed, uid, 0, np.arange(n)) >> np.uint64(32)
    keep_idx = np.lexsort((np.arange(n), k0))[:n_keep]                  # by key, ties by index
    k1 = draw(seed, uid, 1, np.arange(n_keep)) >> np.uint64(32)
    use_idx = np.lexsort((np.arange(n_keep), k1))[:n_dup]
    ctr = np.arange(n_keep, dtype=np.uint64)[:, None] * np.uint64(3 + dim) + np.arange(3 + dim, dtype=np.uint64)[None, :]
    z = normal_from_draw(draw(seed, uid, 2, ctr))
    return keep_idx, use_idx, z[:, :3], z[:, 3:]


def augment_cloud_ref(pos, x, n_keep: int, n_dup: int, noise_sd: float, angle_deg: float, seed: int, uid: int):
    """What b2pn_augment_batch writes for one cloud: (out_pos, out_x, out_src), float64."""
    pos = np.asarray(pos, dtype=np.float64)
    n = pos.shape[0]
    xx = np.zeros((n, 0)) if x is None else np.asarray(x, dtype=np.float64)
    keep_idx, use_idx, zc, zx = counter_draws(n, n_keep, n_dup, xx.shape[1], seed, uid)
    sd = abs(noise_sd)
    out_pos, out_x = apply_augmentation(pos, xx, keep_idx, sd * zc, sd * zx, noise_sd >= 0, use_idx, angle_deg)
    out_src = np.concatenate([keep_idx, keep_idx[use_idx]])
    return out_pos, out_x, out_src
