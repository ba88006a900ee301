"""GPU debug helper: run the tcgen05 self-test with structured inputs and print where it differs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_tc_gpu import _run

dev = torch.device("cuda:0")
for mode in (0, 1):
    for (m, k, rows) in [(64, 64, 128), (128, 128, 256), (256, 128, 256), (128, 192, 128)]:
        try:
            err, got, want = _run(m, k, rows, mode, dev)
        except Exception as e:  # noqa
            print("mode", mode, (m, k, rows), "EXC", e)
            continue
        bad = ((got - want).abs() > 1e-2 * want.abs().max())
        print("mode", mode, (m, k, rows), "rel err", err, "bad frac", float(bad.float().mean()),
              "nan", int(torch.isnan(got).sum()))
        if bad.any():
            rows_bad = bad.any(1).nonzero().flatten()[:8].tolist()
            cols_bad = bad.any(0).nonzero().flatten()[:16].tolist()
            print("   bad out-channels:", rows_bad, " bad rows:", cols_bad)
            print("   got ", got[:2, :6].tolist())
            print("   want", want[:2, :6].tolist())
