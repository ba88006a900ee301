"""Kernel 1 alone on BASELINE configs[1] shapes (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dl_biomass_b200 import ops
from dl_biomass_b200.data import Batch, synthetic_clouds
dev = torch.device("cuda:0")
b = Batch.from_data_list(synthetic_clouds(1234, 12, 10000, 1, False))
pos = b.pos.to(dev)
lv = ops.build_levels(b.cloud_sizes, [0.2], dev)
for _ in range(3):
    idx, p1, _ = ops.fps(pos, lv[0], lv[1])
nbr, cnt = ops.ball_query(pos, p1, lv[0], lv[1], 2.0, 64)
torch.cuda.synchronize()
print("ok", int(idx[1]), float(cnt.float().mean()))
