"""on-device augmentation + batch assembly from the resident cloud cache vs the numpy oracle, and a ragged
training loop fed by it.  Writes gpurun_out/bench_augment.json."""
import json, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from dl_biomass_b200.augment import CloudCache
from dl_biomass_b200.data import synthetic_clouds
from dl_biomass_b200.pointnet2_regressor import Net
from dl_biomass_b200.train import PipelinedTrainStep, make_optimizer
from oracle import augment_ref as ar

dev = torch.device("cuda:0")
N, B, PLOTS = 7168, 12, 240
clouds = synthetic_clouds(4000, PLOTS, N)
cache = CloudCache(clouds, dev)
rng = random.Random(1)
out = {"gpu": torch.cuda.get_device_name(0), "points_per_cloud": N, "batch": B, "plots_cached": PLOTS}

# ---- augmentation + batch assembly alone
ids = [rng.sample(range(PLOTS), B) for _ in range(60)]
for i in range(5):
    cache.batch(ids[i], rng, seed=1, epoch=i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for i in range(5, 55):
    cache.batch(ids[i], rng, seed=1, epoch=i)
e1.record()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 50
out["augment_batch_ms_device"] = round(e0.elapsed_time(e1) / 50, 4)
out["augment_batch_ms_wall"] = round(wall * 1e3, 4)
out["augment_clouds_per_s"] = round(B / wall, 1)

# ---- the numpy oracle on the same work (what the reference's loader workers do per sample, minus the LAS read)
nprng = np.random.default_rng(3)
t0 = time.perf_counter()
reps = 3
for r in range(reps):
    for cid in ids[r]:
        c = clouds[cid]
        n_keep, n_dup, sd, angle = ar.draw_scalars(rng, N)
        keep = nprng.permutation(N)[:n_keep]
        use = nprng.permutation(n_keep)[:n_dup]
        ar.apply_augmentation(c.pos.numpy(), c.x.numpy(), keep, nprng.normal(0, abs(sd), (n_keep, 3)),
                              nprng.normal(0, abs(sd), (n_keep, 1)), sd >= 0, use, angle)
cpu = (time.perf_counter() - t0) / reps
out["numpy_oracle_batch_ms_1_thread"] = round(cpu * 1e3, 3)
out["numpy_oracle_clouds_per_s"] = round(B / cpu, 1)

# ---- ragged training loop: every step trains on a freshly augmented batch (eager pipelined step)
torch.manual_seed(0)
net = Net(1, "ReLU", 0, 0.5, precision="bf16").to(dev)
net.train()
opt = make_optimizer(net)   # FlatAdam over the parameter arena
nb = lambda i: cache.batch(ids[i % len(ids)], rng, seed=2, epoch=i)
with PipelinedTrainStep(net, opt, nb(0), graph=False) as stepper:
    for i in range(1, 8):
        stepper.step(nb(i))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    steps = 40
    for i in range(8, 8 + steps):
        loss = stepper.step(nb(i))
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / steps
out["ragged_train_step_ms_wall"] = round(t * 1e3, 3)
out["ragged_train_clouds_per_s"] = round(B / t, 1)
out["last_loss"] = float(loss)
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_augment.json"), "w"), indent=1)
