"""CPU, gloo, world_size 2: the data-parallel gradient reducer (dl_biomass_b200/parallel.py) and cloud sharding.
The module under the reducer is the CPU oracle network (the product network has no CPU path)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dl_biomass_b200.data import Batch, synthetic_clouds
from dl_biomass_b200.parallel import DEFAULT_BUCKETS, GradReducer, shard_clouds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_net():
    from oracle import ref
    return ref.seeded_init_(ref.NetRef(1, "ReLU", 0, 0.0), seed=7)


def _rank_batch(rank):
    return Batch.from_data_list(synthetic_clouds(500 + 10 * rank, 2, 160, 1, True))


def _local_grads(rank):
    from oracle import ref
    net = _make_net()
    b = _rank_batch(rank)
    ref.weighted_mse(net(b), b.y).backward()
    return {k: p.grad.clone() for k, p in net.named_parameters()}


def _worker(rank, world, port, out_dir):
    from oracle import ref
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)          # replicas start different on purpose ...
    net = ref.NetRef(1, "ReLU", 0, 0.0)
    if rank == 0:
        net = _make_net()
    red = GradReducer(net)                 # ... and the reducer broadcasts rank 0's parameters
    opt = ref.make_adam(net.parameters())
    b = _rank_batch(rank)
    opt.zero_grad(set_to_none=False)
    red.prepare()
    ref.weighted_mse(net(b), b.y).backward()
    red.finish()
    grads = {k: p.grad.clone() for k, p in net.named_parameters()}
    opt.step()
    params = {k: p.detach().clone() for k, p in net.named_parameters()}
    torch.save({"grads": grads, "params": params, "buckets": [f.numel() for f in red.flat],
                "wire": red.wire_bytes_per_step()}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_clouds_partitions_everything():
    for n, w in [(24, 2), (12, 8), (13, 4), (3, 8)]:
        got = [i for r in range(w) for i in shard_clouds(n, r, w)]
        assert got == list(range(n))


@pytest.mark.timeout(300)
def test_grad_reducer_gloo_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    g0, g1 = _local_grads(0), _local_grads(1)
    for k in g0:
        want = 0.5 * (g0[k] + g1[k])
        torch.testing.assert_close(r0["grads"][k], want, rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(r1["grads"][k], want, rtol=1e-5, atol=1e-7)
        assert torch.equal(r0["params"][k], r1["params"][k]), k    # replicas stay bit-identical after the step
    assert len(r0["buckets"]) == len(DEFAULT_BUCKETS) and sum(r0["buckets"]) == 953732
    assert r0["buckets"][0] == 724992 + 148740                       # head + SA3 first (backward order)
    assert r0["wire"] == 953732 * 4                                  # 2*(G-1)/G * bytes at G = 2
