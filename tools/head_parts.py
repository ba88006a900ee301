"""The regression head + loss alone (forward, weighted MSE, backward) as one CUDA graph: what this latency-bound tail of
the training stream costs inside the step (cold L2: flushed between replays; warm: back-to-back replays)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dl_biomass_b200 import head  # noqa: E402
from dl_biomass_b200.pointnet2_regressor import Net  # noqa: E402
from dl_biomass_b200.train import loss_and_grad, make_optimizer  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(1)
net = Net(1, "ReLU", 0, 0.5, precision="bf16").to(dev).train()
opt = make_optimizer(net)
x3 = torch.randn(12, 1024, device=dev, requires_grad=True)
y = torch.rand(12, 4, device=dev)
S = torch.cuda.Stream(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
net._head_rng_counter = torch.zeros((), dtype=torch.int64, device=dev) if net._head_rng_counter is None else net._head_rng_counter


def fn():
    opt.zero_grad()
    out = head.head_apply(net.mlp, x3, net._head_rng_counter, net._head_seed)
    loss, grad = loss_and_grad(out, y)
    out.backward(grad)


S.wait_stream(torch.cuda.current_stream(dev))
with torch.cuda.stream(S):
    for _ in range(3):
        fn()
torch.cuda.current_stream(dev).wait_stream(S)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=S):
    fn()


def timeit(cold):
    ts = []
    for _ in range(30):
        if cold:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1000)
    ts.sort()
    return ts[len(ts) // 2]


print(f"head fwd + loss + bwd as one graph: cold L2 {timeit(True):.1f} us, warm {timeit(False):.1f} us")

# ---- the three library calls separately (each as its own graph)
out_keep = {}


def f_fwd():
    out_keep["out"] = head.head_apply(net.mlp, x3, net._head_rng_counter, net._head_seed)


def graph_of(fn):
    S.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(S):
        for _ in range(2):
            fn()
    torch.cuda.current_stream(dev).wait_stream(S)
    torch.cuda.synchronize()
    gg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gg, stream=S):
        fn()
    return gg


g = graph_of(f_fwd)
print(f"  forward (lin0 + rest, 2 launches): cold {timeit(True):.1f} us, warm {timeit(False):.1f} us")
with torch.cuda.stream(S):
    f_fwd()
    out = out_keep["out"]
    loss, grad = loss_and_grad(out, y)
torch.cuda.synchronize()
g = graph_of(lambda: loss_and_grad(out, y))
print(f"  weighted MSE (1 launch): cold {timeit(True):.1f} us, warm {timeit(False):.1f} us")


def f_bwd():
    opt.zero_grad()
    out.backward(grad, retain_graph=True)


g = graph_of(f_bwd)
print(f"  backward (1 launch): cold {timeit(True):.1f} us, warm {timeit(False):.1f} us")
g = graph_of(lambda: None) if False else None
