"""Run-to-run spread of the bf16 gradients (VERDICT r01 weak #5 / ADVICE): the same batch, the same weights, N backward
passes; per parameter tensor the largest |g_run - g_run0| relative to the tensor's own max and to the level's max, in the
default mode (fp32 atomics for the dW splits and the level-2 -> level-1 feature-gradient scatter) and in the
deterministic mode (fixed-order dW reduction, fixed-point scatter).  Writes gpurun_out/grad_spread.md.
    python tools/grad_spread.py [clouds points runs]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dl_biomass_b200 import sa  # noqa: E402
from dl_biomass_b200.data import Batch, synthetic_clouds  # noqa: E402
from dl_biomass_b200.pointnet2_regressor import Net  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 12
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
RUNS = int(sys.argv[3]) if len(sys.argv) > 3 else 20
dev = torch.device("cuda:0")
lines = [f"# Run-to-run spread of the bf16 gradients ({RUNS} runs, same input, same weights)", ""]
for (b_, n_) in ((3, 640), (B, N)):
    batch = Batch.from_data_list(synthetic_clouds(321, b_, n_, 1, True)).to(dev)
    torch.manual_seed(11)
    net = Net(1, "ReLU", 0, 0.0, precision="bf16").to(dev).set_random_start(False)
    net.train()
    state = {k: v.clone() for k, v in net.state_dict().items()}
    names = [n for n, _ in net.named_parameters()]

    def grads():
        net.load_state_dict(state)
        net.zero_grad(set_to_none=True)
        out = net(batch)
        out.square().sum().backward()
        torch.cuda.synchronize()
        return out.detach().clone(), [p.grad.detach().clone() for p in net.parameters()]

    for det in (False, True):
        with sa.options(deterministic=det):
            o0, g0 = grads()
            worst_own = {n: 0.0 for n in names}
            worst_lvl = {n: 0.0 for n in names}
            out_diff = 0.0
            lvl_scale = {}
            for n, g in zip(names, g0):
                k = n.split(".")[0]
                lvl_scale[k] = max(lvl_scale.get(k, 0.0), float(g.abs().max()))
            for _ in range(RUNS - 1):
                o, g = grads()
                out_diff = max(out_diff, float((o - o0).abs().max()))
                for n, a, c in zip(names, g0, g):
                    d = float((a - c).abs().max())
                    worst_own[n] = max(worst_own[n], d / max(float(a.abs().max()), 1e-30))
                    worst_lvl[n] = max(worst_lvl[n], d / max(lvl_scale[n.split(".")[0]], 1e-30))
        lines += [f"## {b_} clouds x {n_} points, {'deterministic' if det else 'default (atomic)'} mode", "",
                  f"forward outputs: max |diff| over runs = {out_diff:.3e}", "",
                  "| parameter | max diff / max|g| of the tensor | max diff / max|g| of its level |", "|---|---|---|"]
        for n in names:
            lines.append(f"| {n} | {worst_own[n]:.3e} | {worst_lvl[n]:.3e} |")
        per_level = {}
        for n in names:
            k = n.split(".")[0]
            per_level[k] = max(per_level.get(k, 0.0), worst_lvl[n])
        lines += ["", "worst per level (relative to the level's largest gradient): " +
                  ", ".join(f"{k} {v:.3e}" for k, v in per_level.items()), ""]
        print(lines[-2], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "grad_spread.md"), "w").write("\n".join(lines) + "\n")
