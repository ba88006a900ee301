"""GPU: time every FPS variant and the ball query on BASELINE configs (writes gpurun_out/sweep_grouping.json)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dl_biomass_b200 import _lib, ops  # noqa: E402
from dl_biomass_b200.data import Batch, synthetic_clouds  # noqa: E402


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    dev = torch.device("cuda:0")
    out = {"gpu": torch.cuda.get_device_name(0), "fps": [], "bq": []}
    quick = os.environ.get("QUICK") == "1"
    cfgs = [(12, 10000, 0.2, "sa1_10k"), (12, 2000, 0.25, "sa2_10k"), (8, 100000, 0.2, "sa1_100k"),
            (64, 10000, 0.2, "sa1_10k_b64")]
    for (B, n, ratio, tag) in (cfgs[:2] if quick else cfgs):
        b = Batch.from_data_list(synthetic_clouds(1234, B, n, 1, False))
        pos = b.pos.to(dev)
        lv = ops.build_levels([n] * B, [ratio], dev)
        m = lv[1].sizes[0]
        ref_idx = None
        for cluster in ((1, -2) if quick else (1, -2, -1, 2, 4, 8, 16)):
            for threads in ((256, 512, 640, 768, 1024) if cluster == -2 else (256, 512, 1024)):
                try:
                    idx, _, _ = ops.fps(pos, lv[0], lv[1], cluster=cluster, threads=threads)
                    torch.cuda.synchronize()
                except RuntimeError as e:
                    out["fps"].append({"cfg": tag, "cluster": cluster, "threads": threads, "error": str(e)[:80]})
                    continue
                if ref_idx is None:
                    ref_idx = idx.clone()
                same = bool(torch.equal(idx, ref_idx))
                ms = timeit(lambda: ops.fps(pos, lv[0], lv[1], cluster=cluster, threads=threads), iters=3 if n > 20000 else 7)
                scan_gb = B * m * n * 16 / 1e9
                rec = {"cfg": tag, "cluster": cluster, "threads": threads, "ms": round(ms, 4),
                       "us_per_iter": round(ms * 1e3 / m, 4), "scan_GBps": round(scan_gb / (ms * 1e-3), 1),
                       "same_as_first": same}
                print(rec, flush=True)
                out["fps"].append(rec)
        ms = timeit(lambda: ops.fps(pos, lv[0], lv[1]))
        out["fps"].append({"cfg": tag, "cluster": "auto", "ms": round(ms, 4)})
        print(out["fps"][-1], flush=True)
        # ball query at this level
        idx, pos_out, _ = ops.fps(pos, lv[0], lv[1])
        for r in ((2.0, 4.0) if ratio == 0.2 else (8.0, 16.0)):
            ms = timeit(lambda: ops.ball_query(pos, pos_out, lv[0], lv[1], r, 64))
            nbr, cnt = ops.ball_query(pos, pos_out, lv[0], lv[1], r, 64)
            rec = {"cfg": tag, "r": r, "ms": round(ms, 4), "avg_deg": round(float(cnt.float().mean()), 2),
                   "scan_GBps": round(B * m * n * 12 / 1e9 / (ms * 1e-3), 1)}
            print(rec, flush=True)
            out["bq"].append(rec)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sweep_grouping.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
