"""Run Kernel 1 (FPS) alone on the bench shape -- the target of `ncu --set full` captures.
usage: run_fps.py [cluster threads [n [ratio]]]   (cluster -1 = spatially pruned kernel, 0 = auto)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dl_biomass_b200 import ops  # noqa: E402
from dl_biomass_b200.data import Batch, synthetic_clouds  # noqa: E402

cluster = int(sys.argv[1]) if len(sys.argv) > 1 else 0
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
ratio = float(sys.argv[4]) if len(sys.argv) > 4 else 0.2
dev = torch.device("cuda:0")
b = Batch.from_data_list(synthetic_clouds(1234, 12, n, 1, False))
pos = b.pos.to(dev)
lv = ops.build_levels([n] * 12, [ratio], dev)
for _ in range(3):
    idx, _, _ = ops.fps(pos, lv[0], lv[1], cluster=cluster, threads=threads)   # the variant is a per-call option
torch.cuda.synchronize()
print("ok", idx[:4].tolist())
