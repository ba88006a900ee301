"""CPU experiment: which of the bf16 mode's roundings cost how much of the 2e-2 output tolerance on the HEADLINE shape
(12 clouds x 10 000 points, train-mode BatchNorm)?  Runs the float64 oracle with the bf16 roundings emulated per level /
per rounding site and prints the relative error of the [12,4] outputs against the unrounded float64 oracle.
    python tools/bf16_attribution.py [clouds points [bf16|fp16]]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from oracle import ref  # noqa: E402
from dl_biomass_b200.data import Batch, synthetic_clouds  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 12
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
b = Batch.from_data_list(synthetic_clouds(1234, B, N, 1, False))


def run(cfg):
    net = ref.seeded_init_(ref.NetRef(1, "ReLU", 0, 0.0), seed=7).double()
    net.train()
    for name, parts in cfg.items():
        m = {"sa1": net.sa1_module.conv.local_nn, "sa2": net.sa2_module.conv.local_nn, "sa3": net.sa3_module.nn}[name]
        m.emulate_bf16, m.emulate_parts = True, parts
    d = type("D", (), {})()
    d.x, d.pos, d.batch, d.ptr = b.x.double(), b.pos.double(), b.batch, b.ptr
    with torch.no_grad():
        return net(d)


fmt = {"bf16": torch.bfloat16, "fp16": torch.float16}[sys.argv[3] if len(sys.argv) > 3 else "bf16"]
ref._RoundBF16.fmt = fmt
print("operand format:", fmt)
want = run({})
cases = {"all levels, all roundings": {"sa1": "WXZA", "sa2": "WXZA", "sa3": "WXZA"},
         "sa1 only": {"sa1": "WXZA"}, "sa2 only": {"sa2": "WXZA"}, "sa3 only": {"sa3": "WXZA"},
         "weights only": {"sa1": "W", "sa2": "W", "sa3": "W"},
         "level inputs only": {"sa1": "X", "sa2": "X", "sa3": "X"},
         "zhat only": {"sa1": "Z", "sa2": "Z", "sa3": "Z"},
         "activation operand only": {"sa1": "A", "sa2": "A", "sa3": "A"},
         "all but zhat": {"sa1": "WXA", "sa2": "WXA", "sa3": "WXA"},
         "sa1+sa2 only": {"sa1": "WXZA", "sa2": "WXZA"}}
for name, cfg in cases.items():
    got = run(cfg)
    print(f"{name:32s} out rel err {float((got - want).abs().max() / want.abs().max()):.3e}", flush=True)
