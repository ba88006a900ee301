#!/usr/bin/env python
"""bench.py -- train tree-clouds/sec of the PointNet++ biomass regressor hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--precision bf16|fp32] [--impl reference]

A "step" is one pass of /root/reference/main.py:150-172 over one batch of synthetic tree clouds
(forward + weighted MSE + backward + gradient all-reduce + Adam), BASELINE.json configs[1]:
10,000-point clouds, 12 clouds per GPU.  Rank 0 prints ONE JSON line (contract in the task brief):
value = device-resident throughput, e2e = through Net.forward with host buffers (H2D + loss D2H inside
the timed region), roofline = Kernel 1 (FPS, level 1) timed alone against the measured HBM peak,
cpu_baseline = the CPU oracle on this box's host cores.  `--impl reference` times the reference's CPU
path (its third-party kernels restated by oracle/) instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "train tree-clouds/sec @10k pts"
UNIT = "tree-clouds/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=None, choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=12, help="tree clouds per GPU per step")
    ap.add_argument("--points", type=int, default=10000)
    ap.add_argument("--pool", type=int, default=4, help="distinct batches rotated through")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of one CUDA graph per step")
    ap.add_argument("--join", default="end", choices=["backward", "end"], help="where the sampling branch joins")
    ap.add_argument("--no-uncap-l1", action="store_true", help="keep the SM cap through the level-1 backward")
    ap.add_argument("--no-cap", action="store_true", help="do not cap persistent kernels while the branch runs")
    ap.add_argument("--no-pregroup", action="store_true", help="ball query / row packing inside forward")
    ap.add_argument("--aux", action="store_true", help="third stream for the level-1 grouping")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="sample (FPS) every batch inside its own step instead of one step ahead on a second stream")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
#  helpers
# ------------------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 1590.0)), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc, self.lines, self.thread = None, [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except (OSError, ValueError):
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_pool(args, rank, pin):
    """`pool` distinct batches of `batch` clouds for this rank (seeds as SURVEY.md 8(d): S + rank*1e6 + c)."""
    from dl_biomass_b200.data import Batch, synthetic_clouds
    pool = []
    for i in range(args.pool):
        clouds = synthetic_clouds(1234 + rank * 1_000_000 + i * args.batch, args.batch, args.points, 1, False)
        b = Batch.from_data_list(clouds)
        pool.append(b.pin_memory() if pin else b)
    return pool


def batch_h2d_bytes(b, with_batch_vector=True):
    ts = (b.x, b.pos, b.batch, b.y) if with_batch_vector else (b.x, b.pos, b.y)
    return sum(t.numel() * t.element_size() for t in ts if t is not None)


# ------------------------------------------------------------------------------------------------
#  CPU arm: the reference's CPU path as restated by oracle/ (torch CPU + C fps/radius)
# ------------------------------------------------------------------------------------------------
def cpu_train_steps(args, steps, warmup, budget_s=150.0):
    from oracle import ref
    torch.set_num_threads(os.cpu_count() or 1)
    from dl_biomass_b200.data import Batch, synthetic_clouds
    net = ref.seeded_init_(ref.NetRef(1, "ReLU", 0, 0.5), seed=7)
    net.train()
    opt = ref.make_adam(net.parameters())
    clouds_per_step = args.batch

    def one(i, ncl):
        b = Batch.from_data_list(synthetic_clouds(1234 + i * args.batch, ncl, args.points, 1, False))
        t0 = time.perf_counter()
        opt.zero_grad()
        loss = ref.weighted_mse(net(b), b.y)
        loss.backward()
        opt.step()
        return time.perf_counter() - t0

    t_first = one(0, clouds_per_step)  # also serves as warm-up
    if t_first * (steps + max(warmup - 1, 0)) > budget_s:  # bound the sample: fewer clouds per step
        clouds_per_step = max(1, int(clouds_per_step * budget_s / (t_first * (steps + max(warmup - 1, 0)))))
    for i in range(1, warmup):
        one(i, clouds_per_step)
    ts = [one(100 + i, clouds_per_step) for i in range(steps)]
    total = sum(ts)
    return {"value": clouds_per_step * steps / total, "ms_per_step": 1e3 * total / steps,
            "clouds_per_step": clouds_per_step, "cores": torch.get_num_threads(), "host_cores": os.cpu_count(),
            "oracle_threads": ref.num_threads()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_train_steps(args, args.steps, args.warmup)
    sample = (f"{r['clouds_per_step']} of {args.batch} clouds x {args.points} pts per step, "
              f"{args.steps} steps, full train step (fwd+loss+bwd+Adam)")
    line = {"impl": "reference", "metric": METRIC, "value": round(r["value"], 4), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(r["ms_per_step"], 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"train step, {args.points}-pt clouds, batch {args.batch} (BASELINE configs[0])",
                       "note": "reference CPU path: torch_cluster/torch_scatter/PyG are not installable here, so their "
                               "kernels are the oracle's restatement (oracle/ref.py, oracle/b2pn_oracle.c)"},
            "cpu_baseline": {"value": round(r["value"], 4), "unit": UNIT, "cores": r["cores"], "kind": "port",
                             "sample": sample},
            "e2e": {"value": round(r["value"], 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
#  B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    from dl_biomass_b200 import _lib, ops
    from dl_biomass_b200.parallel import GradReducer
    from dl_biomass_b200.pointnet2_regressor import Net
    from dl_biomass_b200.train import GraphedTrainStep, PipelinedTrainStep, make_optimizer, train_step

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200: no CUDA device (the product path has no CPU fallback)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = _lib.lib()

    precision = args.precision
    if precision is None:
        from dl_biomass_b200 import sa
        precision = "bf16" if getattr(sa, "bf16_available", lambda: False)() else "fp32"

    torch.manual_seed(7)
    net = Net(1, "ReLU", 0, 0.5, precision=precision).to(dev)
    net.train()
    # one graph per step; with world > 1 the graph holds forward + backward and the NCCL all-reduce + Adam follow it
    # eagerly (capturing the collective itself hung on the 2-GPU box in round 1)
    use_graph = not args.no_graph
    opt = make_optimizer(net.parameters(), capturable=use_graph and world == 1)
    reducer = GradReducer(net) if world > 1 else None

    pool_host = make_pool(args, rank, pin=True)
    pool_dev = [b.to(dev) for b in pool_host]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(step_fn, steps):
        """Per-step CUDA events on the current stream, L2 flushed (untimed) before every step."""
        evs = []
        barrier()
        for i in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step_fn(i)
            b.record()
            evs.append((a, b))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # The step (forward, loss, backward, all-reduce, Adam) is captured once and replayed (single GPU), and the
    # farthest-point sampling of batch i+1 runs on a second stream while batch i trains (train.PipelinedTrainStep):
    # every timed step still contains one full sampling and one full training pass, only overlapped.
    pipeline = not args.no_pipeline
    graphed = stepper = None
    if pipeline:
        stepper = PipelinedTrainStep(net, opt, pool_dev[0], reducer, graph=use_graph, join=args.join,
                                     cap=not args.no_cap, grouping=not args.no_pregroup, aux=args.aux,
                                     uncap_level1_backward=not args.no_uncap_l1)
    elif use_graph:
        graphed = GraphedTrainStep(net, opt, pool_dev[0], reducer)

    def run_step(batch):
        if stepper is not None:
            return stepper.step(batch)
        return graphed(batch) if graphed is not None else train_step(net, opt, batch, reducer)

    # ---- resident-input arm ("value") ------------------------------------------------------------
    def step_resident(i):
        run_step(pool_dev[(i + 1) % len(pool_dev)])

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = lib.b2pn_launch_count()
    total_ms = timed(step_resident, args.steps)
    launches = lib.b2pn_launch_count() - l0
    clocks = sampler.stop() if sampler else None
    if stepper is not None and use_graph:  # replayed kernels do not pass through the library's launch counter
        launches = stepper.launches_per_step * args.steps
    elif graphed is not None:
        launches = graphed.launches_per_replay * args.steps
    value = world * args.batch * args.steps / (total_ms * 1e-3)

    # ---- end-to-end arm: host buffers in, loss out ---------------------------------------------------
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def step_e2e(i):
        hb = pool_host[(i + 1) % len(pool_host)]
        if stepper is not None or graphed is not None:
            loss = run_step(hb)  # pinned host tensors are copied straight into the step's input buffers
        else:
            loss = train_step(net, opt, hb.to(dev, non_blocking=True), reducer)
        loss_host.copy_(loss, non_blocking=True)

    for i in range(3):
        step_e2e(i)
    e2e_ms = timed(step_e2e, args.steps)
    e2e_value = world * args.batch * args.steps / (e2e_ms * 1e-3)
    sm_limit_note = stepper.sm_limit if stepper is not None else 0
    if stepper is not None:
        stepper.close()

    # ---- roofline of the dominant grouping kernel: FPS level 1, timed alone -----------------------------
    hbm_peak, _, peak_kind = measured_peaks()
    b0 = pool_dev[0]
    lv = ops.build_levels(b0.cloud_sizes, [0.2], dev)
    torch.cuda.synchronize(dev)
    for _ in range(3):
        ops.fps(b0.pos, lv[0], lv[1])
    reps = 10
    evs = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.fps(b0.pos, lv[0], lv[1])
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize(dev)
    fps_ms = sum(a.elapsed_time(b) for a, b in evs) / reps
    scan_bytes = sum(16 * n * m for n, m in zip(lv[0].sizes, lv[1].sizes))  # m*n*16 per cloud, SURVEY 8(d)
    achieved = scan_bytes / 1e9 / (fps_ms * 1e-3)
    traffic = None
    tfile = os.path.join(ROOT, "profiles", "fps_dram_bytes.json")
    if os.path.exists(tfile):
        traffic = json.load(open(tfile)).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "fps_kernel (Kernel 1, SA1: 10000->2000 per cloud)",
                "achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(achieved / hbm_peak, 4),
                "traffic": traffic, "peak_kind": f"{peak_kind} (burst copy bandwidth)",
                "bytes_model": "scan-equivalent m*n*16 B per cloud (SURVEY 8(d)); register-resident, so DRAM traffic is "
                               "the compulsory n*12+m*8 B",
                "ms_per_launch": round(fps_ms, 4), "share_of_step": round(fps_ms / (total_ms / args.steps), 3),
                "note": "timed alone here; inside the step it runs on the sampling branch of the graph, concurrently with "
                        "the training kernels" if stepper is not None else "timed alone"}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu_baseline:
        r = cpu_train_steps(args, steps=1, warmup=1, budget_s=30.0)
        cpu = {"value": round(r["value"], 4), "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": f"1 warm-up + 1 timed train step of {r['clouds_per_step']} clouds x {args.points} pts "
                         f"(oracle NetRef: torch CPU + C fps/radius, {r['oracle_threads']} threads)"}

    line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if precision == "bf16" else "f32",
            "data": "synthetic",
            "config": {"workload": f"PointNet++ regressor train step, {args.points}-pt clouds, batch {args.batch}/GPU, "
                                   f"F=1, {precision} MLPs (BASELINE configs[1])",
                       "global_batch": world * args.batch, "points_per_cloud": args.points,
                       "parallelism": f"dp{world}", "l2": "256 MB buffer rewritten before every timed step",
                       "timing": "per-step CUDA events on the launching stream, summed; max over ranks",
                       "launch": "one CUDA graph replay per step" if use_graph else "eager kernel launches",
                       "pipeline": ("FPS" + (" + ball query + row compaction + level-1 gather" if stepper.grouping else "")
                                    + " of batch i+1 on a second stream during step i (every timed step contains one "
                                    f"full sampling and one full training pass); join at {stepper.join_at}; persistent "
                                    f"kernels capped at {sm_limit_note} CTAs" + (" until the level-1 backward" if stepper.uncap_l1 else ""))
                                   if stepper is not None else "none"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "ms_per_step": round(e2e_ms / args.steps, 4),
                    "h2d_bytes_per_step": batch_h2d_bytes(pool_host[0], with_batch_vector=not use_graph),
                    "d2h_bytes_per_step": 4},
            "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
