"""Host-side (Python) cost of the eager training step: cProfile over 30 steps (the ragged-batch loop is CPU-bound)."""
import cProfile, os, pstats, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dl_biomass_b200.data import Batch, synthetic_clouds
from dl_biomass_b200.pointnet2_regressor import Net
from dl_biomass_b200.train import make_optimizer, train_step
dev = torch.device("cuda:0")
torch.manual_seed(7)
net = Net(1, "ReLU", 0, 0.5, precision="bf16").to(dev)
opt = make_optimizer(net)
bs = [Batch.from_data_list(synthetic_clouds(1234 + 50 * i, 12, 7168, 1, True)).to(dev) for i in range(8)]
for i in range(8):
    train_step(net, opt, bs[i])
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
for i in range(30):
    train_step(net, opt, bs[i % 8])
pr.disable()
t_launch = time.perf_counter() - t0
torch.cuda.synchronize()
print("host ms/step (launch only)", 1e3 * t_launch / 30, " wall incl. drain", 1e3 * (time.perf_counter() - t0) / 30)
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
