"""GPU parity on the inputs a lidar pipeline produces at its edges: clouds of 1..7 points (FPS keeps ceil(ratio * n) >= 1
point, a level-2 cloud of one point), points that see no neighbour but themselves, clouds made of one repeated point,
neighbourhoods saturated at K = 64 (/root/reference/pointnet2_regressor.py:14-16: ``max_num_neighbors=64`` keeps the
first 64 in index order), and an empty batch.  Grouping is compared bit for bit, the network in both precision modes
against the oracle (evaluation mode: BatchNorm on running statistics, so a batch of tiny clouds is well defined)."""
import numpy as np
import pytest
import torch

from oracle import ref
from dl_biomass_b200 import ops
from dl_biomass_b200.data import Batch, Data

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def _batch(sizes, seed, spread=6.0, F=1):
    g = torch.Generator().manual_seed(seed)
    items = []
    for n in sizes:
        pos = (torch.rand(n, 3, generator=g) - 0.5) * spread
        items.append(Data(x=torch.randn(n, F, generator=g), pos=pos, y=torch.rand(1, 4, generator=g) * 10))
    return Batch.from_data_list(items)


def _levels_vs_oracle(b, dev, ratio, r, K=64):
    sizes = (b.ptr[1:] - b.ptr[:-1]).tolist()
    lv = ops.build_levels(sizes, [ratio], dev)
    pos = b.pos.to(dev)
    idx, pos_out, batch_out = ops.fps(pos, lv[0], lv[1], None)
    nbr, cnt = ops.ball_query(pos, pos_out, lv[0], lv[1], r, K)
    torch.cuda.synchronize()
    want_idx = ref.fps_ref(b.pos, b.ptr, ratio, None)
    qptr = ref.sample_ptr(b.ptr, ratio)
    want_nbr, want_cnt = ref.ball_query_ref(b.pos, b.pos[want_idx], b.ptr, qptr, r, K)
    assert torch.equal(idx.cpu(), want_idx)
    assert torch.equal(cnt.cpu(), want_cnt)
    assert torch.equal(nbr.cpu(), want_nbr)
    return want_idx, want_cnt


def _net_pair(dev, precision):
    from dl_biomass_b200.pointnet2_regressor import Net
    netr = ref.seeded_init_(ref.NetRef(1, "ReLU", 0, 0.0), 11)
    net = Net(1, "ReLU", 0, 0.0, precision=precision).to(dev).set_random_start(False)
    ref.seeded_init_(net, 11)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():   # running statistics that are not the identity
        for (k, v), (_, vr) in zip(net.named_buffers(), netr.named_buffers()):
            if k.endswith("running_mean"):
                t = torch.randn(vr.shape, generator=g) * 0.1
            elif k.endswith("running_var"):
                t = torch.rand(vr.shape, generator=g) + 0.5
            else:
                continue
            vr.copy_(t)
            v.copy_(t.to(dev))
    return netr.eval(), net.eval()


@pytest.mark.parametrize("sizes", [[1, 2, 3, 5, 7], [1], [4, 1, 4], [6, 600, 2]])
def test_tiny_clouds_group_and_evaluate_like_the_oracle(cuda_device, sizes):
    b = _batch(sizes, 3)
    idx1, _ = _levels_vs_oracle(b, cuda_device, 0.2, 2.0)
    b1 = Batch.from_data_list([Data(pos=b.pos[idx1][s:e], x=None) for s, e in
                               zip(ref.sample_ptr(b.ptr, 0.2)[:-1].tolist(), ref.sample_ptr(b.ptr, 0.2)[1:].tolist())])
    _levels_vs_oracle(b1, cuda_device, 0.25, 8.0)
    for precision, tol in (("fp32", 1e-4), ("bf16", 2e-2)):
        netr, net = _net_pair(cuda_device, precision)
        with torch.no_grad():
            want = netr(b)
            got = net(b.to(cuda_device))
        torch.cuda.synchronize()
        assert got.shape == want.shape == (len(sizes), 4)
        assert rel_err(got, want) < tol, (precision, rel_err(got, want))


def test_isolated_points_see_only_themselves(cuda_device):
    """Points further apart than both radii: every centroid's neighbourhood is the centroid itself."""
    g = torch.Generator().manual_seed(1)
    grid = torch.stack(torch.meshgrid(torch.arange(6.), torch.arange(6.), torch.arange(4.), indexing="ij"), -1).reshape(-1, 3)
    pos = grid * 40.0 + torch.rand(grid.shape, generator=g)
    b = Batch.from_data_list([Data(x=torch.randn(pos.size(0), 1, generator=g), pos=pos, y=torch.ones(1, 4)),
                              Data(x=torch.randn(50, 1, generator=g), pos=pos[:50] + 1000.0, y=torch.ones(1, 4))])
    _, cnt = _levels_vs_oracle(b, cuda_device, 0.2, 2.0)
    assert int(cnt.max()) == 1 and int(cnt.min()) == 1
    netr, net = _net_pair(cuda_device, "fp32")
    with torch.no_grad():
        assert rel_err(net(b.to(cuda_device)), netr(b)) < 1e-4


def test_a_cloud_of_one_repeated_point(cuda_device):
    """All distances are exactly zero: FPS ties go to the lowest index, the ball query keeps the first K duplicates."""
    pos = torch.cat([torch.full((300, 3), 1.5), (torch.rand(200, 3, generator=torch.Generator().manual_seed(2)) - 0.5) * 4])
    b = Batch.from_data_list([Data(x=torch.ones(300, 1), pos=pos[:300], y=torch.ones(1, 4)),
                              Data(x=torch.ones(200, 1), pos=pos[300:], y=torch.ones(1, 4))])
    _, cnt = _levels_vs_oracle(b, cuda_device, 0.2, 2.0)
    assert int(cnt[:60].min()) == 64          # saturated: 300 candidates, 64 slots
    netr, net = _net_pair(cuda_device, "fp32")
    with torch.no_grad():
        assert rel_err(net(b.to(cuda_device)), netr(b)) < 1e-4


def test_saturated_neighbourhoods_keep_the_first_k_in_index_order(cuda_device):
    rng = np.random.default_rng(0)
    pos = torch.from_numpy(rng.normal(size=(5000, 3)).astype(np.float32) * 0.6)   # ~all points within r = 2 of each other
    b = Batch.from_data_list([Data(x=torch.ones(5000, 1), pos=pos, y=torch.ones(1, 4))])
    for K in (64, 16, 1):
        _, cnt = _levels_vs_oracle(b, cuda_device, 0.2, 2.0, K=K)
        assert int(cnt.min()) == K


def test_empty_batch(cuda_device):
    """No clouds: the operators return empty tensors (nothing to launch), they do not crash."""
    lv = ops.build_levels([], [0.2], cuda_device)
    pos = torch.empty(0, 3, device=cuda_device)
    idx, pos_out, batch_out = ops.fps(pos, lv[0], lv[1], None)
    nbr, cnt = ops.ball_query(pos, pos_out, lv[0], lv[1], 2.0, 64)
    torch.cuda.synchronize()
    assert idx.numel() == 0 and pos_out.shape == (0, 3) and batch_out.numel() == 0
    assert cnt.numel() == 0 and nbr.shape[0] == 0


def test_training_step_on_tiny_clouds_fp32(cuda_device):
    """Train-mode BatchNorm over a handful of rows (3 clouds of 3 / 40 / 9 points): outputs, loss, every gradient and the
    updated running statistics against the oracle."""
    from dl_biomass_b200.pointnet2_regressor import Net
    b = _batch([3, 40, 9], 9, spread=3.0)
    netr = ref.seeded_init_(ref.NetRef(1, "ReLU", 0, 0.0), 11).train()
    net = Net(1, "ReLU", 0, 0.0, precision="fp32").to(cuda_device).set_random_start(False)
    ref.seeded_init_(net, 11).train()
    want = netr(b)
    lw = ref.weighted_mse(want, b.y)
    lw.backward()
    out = net(b.to(cuda_device))
    loss = ref.weighted_mse(out, b.y.to(cuda_device))
    loss.backward()
    torch.cuda.synchronize()
    assert rel_err(out, want) < 1e-4 and rel_err(loss, lw) < 1e-4
    for (k, p), (_, pr) in zip(net.named_parameters(), netr.named_parameters()):
        if pr.grad.abs().max() < 1e-6 * max(1.0, float(lw.detach())):   # biases in front of a BatchNorm: 0 in theory
            continue
        assert rel_err(p.grad, pr.grad) < 2e-3, (k, rel_err(p.grad, pr.grad))
    for (k, v), (_, vr) in zip(net.named_buffers(), netr.named_buffers()):
        assert rel_err(v.float(), vr.float()) < 1e-4, k
