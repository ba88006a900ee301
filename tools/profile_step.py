"""One training step inside a cudaProfilerStart/Stop range (use with `ncu --profile-from-start off`)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dl_biomass_b200.data import Batch, synthetic_clouds
from dl_biomass_b200.pointnet2_regressor import Net
from dl_biomass_b200.train import make_optimizer, train_step

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="bf16")
ap.add_argument("--batch", type=int, default=12)
ap.add_argument("--points", type=int, default=10000)
ap.add_argument("--eval", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(7)
net = Net(1, "ReLU", 0, 0.5, precision=a.precision).to(dev)
opt = make_optimizer(net)   # FlatAdam over the parameter arena
b = Batch.from_data_list(synthetic_clouds(1234, a.batch, a.points, 1, False)).to(dev)
def step():
    if a.eval:
        net.eval()
        with torch.no_grad():
            return net(b)
    return train_step(net, opt, b)
for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
torch.cuda.cudart().cudaProfilerStart()
t0.record(); step(); t1.record()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("step ms", t0.elapsed_time(t1))
