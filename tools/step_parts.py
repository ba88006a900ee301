"""Where a pipelined training step spends its time: the two branches of the step graph timed ALONE (each as its own
CUDA graph, L2 flushed between replays, CUDA events on the replay stream).

  train(cap)  forward + loss + backward + Adam of a pre-sampled batch, persistent kernels capped to `cap` SMs (0 = all)
  sample      Net.sample(batch): FPS x2, ball query x2, row compaction, the gathered level-1 operand
  fwd / bwd   the training branch split at the loss

    python tools/step_parts.py [--batch 12 --points 10000]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dl_biomass_b200 import sa  # noqa: E402
from dl_biomass_b200.data import Batch, synthetic_clouds  # noqa: E402
from dl_biomass_b200.pointnet2_regressor import Net  # noqa: E402
from dl_biomass_b200.train import forward_backward, loss_and_grad, make_optimizer, train_step  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=12)
ap.add_argument("--points", type=int, default=10000)
ap.add_argument("--reps", type=int, default=30)
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(7)
net = Net(1, "ReLU", 0, 0.5, precision="bf16").to(dev)
net.train()
opt = make_optimizer(net)
b = Batch.from_data_list(synthetic_clouds(1234, a.batch, a.points, 1, False)).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


S = torch.cuda.Stream(dev)


def graph_of(fn, warm=3):
    s = S
    s.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s):
        for _ in range(warm):
            fn()
    torch.cuda.current_stream(dev).wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=S):
        fn()
    return g


def time_graph(g, reps=a.reps):
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


samp = net.sample(b, grouping=True)
sms = torch.cuda.get_device_properties(dev).multi_processor_count
res = {}
for cap in (0, sms - a.batch):
    sa.set_sm_limit(cap)
    g = graph_of(lambda: train_step(net, opt, b, None, sampling=samp))
    res[f"train(cap={cap})"] = time_graph(g)
    del g
sa.set_sm_limit(0)
g = graph_of(lambda: net.sample(b, grouping=True))
res["sample"] = time_graph(g)
del g

# forward alone / backward alone (the loss gradient is fixed, the graph of backward re-uses forward's saved tensors)
keep = {}


def fwd():
    keep["out"] = net(b, sampling=samp)
    return keep["out"]


g = graph_of(fwd)
res["forward"] = time_graph(g)
del g
with torch.cuda.stream(S):
    out = fwd()
    loss, grad = loss_and_grad(out, b.y)
torch.cuda.synchronize()


def bwd():
    opt.zero_grad()
    out.backward(grad, retain_graph=True)


g = graph_of(bwd)
res["backward"] = time_graph(g)
del g
for k, v in res.items():
    print(f"{k:>18s}  {v:7.3f} ms")
