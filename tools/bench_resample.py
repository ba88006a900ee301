"""offline FPS resampler (float64) -- plots -> 7 168 points, B200 kernel vs the CPU oracle.
Writes gpurun_out/bench_resample.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from dl_biomass_b200 import resample
from oracle import ref

rng = np.random.default_rng(5)
out = {"gpu": torch.cuda.get_device_name(0), "k": 7168, "rows": []}
for n_pts, n_plots in ((20000, 1), (20000, 148), (60000, 148)):
    plots = [rng.normal(size=(n_pts, 3)) * np.array([5.0, 5.0, 9.0]) + np.array([431234.5, 5312345.25, 250.0]) for _ in range(n_plots)]
    resample.farthest_point_sampling_batch(plots[:1], 7168)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    idx = resample.farthest_point_sampling_batch(plots, 7168)
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    rec = {"points_per_plot": n_pts, "plots": n_plots, "seconds": round(t, 4), "plots_per_s": round(n_plots / t, 2),
           "scan_GBps": round(n_plots * 7168 * n_pts * 32 / t / 1e9, 1)}
    if n_plots == 1:
        t0 = time.perf_counter()
        want = ref.fps_ref_f64(plots[0], 7168, 0)
        rec["cpu_oracle_seconds_1_thread"] = round(time.perf_counter() - t0, 3)
        rec["equal_to_oracle"] = bool(np.array_equal(idx[0], want))
    print(rec, flush=True)
    out["rows"].append(rec)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_resample.json"), "w"), indent=1)
