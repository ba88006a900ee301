"""GPU debug: per-level error of the bf16 mode against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import ref
from dl_biomass_b200 import ops, sa
from dl_biomass_b200.data import Batch, synthetic_clouds
from dl_biomass_b200.pointnet2_regressor import Net

dev = torch.device("cuda:0")
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max())
for (B, n) in [(3, 768), (12, 640), (12, 2000)]:
    for train in (True, False):
        b = Batch.from_data_list(synthetic_clouds(4321, B, n, 1, True))
        netr = ref.seeded_init_(ref.NetRef(1, "ReLU", 0, 0.0), seed=7)
        netr.train(train)
        with torch.no_grad():
            x1, p1, b1, ptr1 = netr.sa1_module(b.x, b.pos, b.batch, b.ptr)
            x2, p2, b2, _ = netr.sa2_module(x1, p1, b1, ptr1)
            x3, _, _ = netr.sa3_module(x2, p2, b2, B)
            want = netr.mlp(x3)
        for prec in ("fp32", "bf16"):
            net = Net(1, "ReLU", 0, 0.0, precision=prec)
            net.load_state_dict(ref.seeded_init_(ref.NetRef(1, "ReLU", 0, 0.0), seed=7).state_dict())
            net = net.to(dev).set_random_start(False)
            net.train(train)
            bg = b.to(dev)
            with torch.no_grad():
                lv = ops.build_levels(b.cloud_sizes, [0.2, 0.25], dev)
                g1, q1, _, _ = net.sa1_module._run(bg.x, bg.pos, lv[0], lv[1])
                g2, q2, bb2, _ = net.sa2_module._run(g1, q1, lv[1], lv[2])
                g3 = net.sa3_module._run(g2, q2, bb2, B)
                out = net.mlp(g3)
            print(f"B={B} n={n} train={train} {prec}: x1 {rel(g1,x1):.2e} x2 {rel(g2,x2):.2e} x3 {rel(g3,x3):.2e} out {rel(out,want):.2e}")
