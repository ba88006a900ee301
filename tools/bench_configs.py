"""BASELINE.json configs[2] (eval sweep) and configs[3] (dense clouds) on one B200 -- parity-test cases with timings,
not bench.py lines.  Writes gpurun_out/bench_configs.json.
    python tools/bench_configs.py [eval] [dense]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dl_biomass_b200.data import Batch, synthetic_clouds  # noqa: E402
from dl_biomass_b200.pointnet2_regressor import Net  # noqa: E402
from dl_biomass_b200.train import make_optimizer, train_step  # noqa: E402

dev = torch.device("cuda:0")
which = set(sys.argv[1:]) or {"eval", "dense"}
out = {"gpu": torch.cuda.get_device_name(0)}


def timed(fn, warm, reps):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


torch.manual_seed(7)
if "eval" in which:
    net = Net(1, "ReLU", 0, 0.5, precision="bf16").to(dev).eval()
    base = synthetic_clouds(4321, 64, 10000, 1, False)
    rows = []
    for B in (64, 128, 256, 512):
        clouds = [base[i % 64] for i in range(B)]
        b = Batch.from_data_list(clouds).to(dev)
        with torch.no_grad():
            ms = timed(lambda: net(b), 2, 5)
        rec = {"batch": B, "points": 10000, "ms": round(ms, 3), "clouds_per_s": round(B / (ms * 1e-3), 1),
               "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}
        print("eval", rec, flush=True)
        rows.append(rec)
        del b
        torch.cuda.empty_cache()
    out["eval_sweep_bf16"] = rows

if "dense" in which:
    for prec in ("bf16",):
        net = Net(1, "ReLU", 0, 0.5, precision=prec).to(dev).train()
        net.sa1_module.r, net.sa2_module.r = 4.0, 16.0   # the reference's "..._w_doubled_radius" runs
        opt = make_optimizer(net.parameters())
        b = Batch.from_data_list(synthetic_clouds(777, 8, 100000, 1, False)).to(dev)
        ms = timed(lambda: train_step(net, opt, b), 2, 5)
        rec = {"precision": prec, "batch": 8, "points": 100000, "radii": [4.0, 16.0], "ms_per_step": round(ms, 2),
               "clouds_per_s": round(8 / (ms * 1e-3), 2), "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}
        print("dense", rec, flush=True)
        out.setdefault("dense_train", []).append(rec)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_configs.json"), "w"), indent=1)
