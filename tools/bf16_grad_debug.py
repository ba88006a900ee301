"""GPU debug: per-parameter gradient error of the bf16 set-abstraction levels against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import ref
from dl_biomass_b200 import sa
from dl_biomass_b200.data import Batch, synthetic_clouds
from dl_biomass_b200.pointnet2_regressor import MLP

dev = torch.device("cuda:0")
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
def pair(chans, seed):
    mref = ref.seeded_init_(ref.MLPRef(chans, act="ReLU"), seed)
    m = MLP(chans, act="ReLU"); m.load_state_dict(mref.state_dict())
    return mref, m.to(dev)

for prec in (sa.PREC_BF16,):
    print("==== precision", prec)
    for (c_in, chans, K) in [(1, [4, 64, 64, 128], 64), (16, [19, 32, 48, 40], 16), (128, [131, 128, 128, 256], 64)]:
        b = Batch.from_data_list(synthetic_clouds(21, 3, 600, max(c_in, 1), True))
        x = torch.randn(b.pos.size(0), c_in, generator=torch.Generator().manual_seed(1))
        idx = ref.fps_ref(b.pos, b.ptr, 0.2); qptr = ref.sample_ptr(b.ptr, 0.2)
        nbr, cnt = ref.ball_query_ref(b.pos, b.pos[idx], b.ptr, qptr, 2.5, K)
        row, col = ref.slots_to_edges(nbr, cnt)
        mref, m = pair(chans, 3)
        mref.emulate_bf16 = True
        if c_in > 16: x = x.to(torch.bfloat16).float()
        xr = x.clone().requires_grad_(True)
        want = ref.point_conv_ref(mref, xr, b.pos, b.pos[idx], row, col)
        gout = torch.randn(want.shape, generator=torch.Generator().manual_seed(2))
        want.backward(gout)
        xg = x.to(dev).requires_grad_(True)
        out, arg = sa.sa_apply(m, xg, b.pos.to(dev), b.pos[idx].to(dev), nbr.to(dev), cnt.to(dev), None,
                               seg_mode=sa.SEG_SLOTS, K=K, n_dst=idx.numel(), precision=prec)
        out.backward(gout.to(dev)); torch.cuda.synchronize()
        print(f"slots c_in={c_in} K={K}: out {rel(out,want):.2e} dx {rel(xg.grad,xr.grad):.2e} " +
              " ".join(f"{k}:{rel(p.grad,pr.grad):.1e}" for (k,p),(_,pr) in zip(m.named_parameters(), mref.named_parameters())))
    g = torch.Generator().manual_seed(5)
    sizes = [130, 257, 64]; n = sum(sizes)
    x = torch.randn(n, 32, generator=g); pos = torch.randn(n, 3, generator=g) * 3
    batch = torch.repeat_interleave(torch.arange(3), torch.tensor(sizes))
    for chans in ([35, 64, 96, 200], [35, 256, 512, 1024]):
        mref, m = pair(chans, 9)
        mref.emulate_bf16 = True
        x = x.to(torch.bfloat16).float()
        xr = x.clone().requires_grad_(True)
        want, _, _ = ref.GlobalSAModuleRef(mref)(xr, pos, batch, 3)
        gout = torch.randn(want.shape, generator=g)
        want.backward(gout)
        xg = x.to(dev).requires_grad_(True)
        out, arg = sa.sa_apply(m, xg, pos.to(dev), None, None, None, batch.to(dev), seg_mode=sa.SEG_CLOUDS, K=0, n_dst=3, precision=prec)
        out.backward(gout.to(dev)); torch.cuda.synchronize()
        d = (xg.grad.cpu() - xr.grad).abs()
        print(f"clouds {chans}: out {rel(out,want):.2e} dx {rel(xg.grad,xr.grad):.2e} (worst row {int(d.max(1).values.argmax())}) " +
              " ".join(f"{k}:{rel(p.grad,pr.grad):.1e}" for (k,p),(_,pr) in zip(m.named_parameters(), mref.named_parameters())))

    # per-channel look at the BatchNorm weight gradients (dgamma) of the last slots case
    for (k, p), (_, pr) in zip(m.named_parameters(), mref.named_parameters()):
        if k.startswith("norms") and k.endswith("weight"):
            d = (p.grad.cpu() - pr.grad).abs()
            top = d.topk(4).indices.tolist()
            print(k, "top err channels", top, "got", [round(float(p.grad[i]), 4) for i in top],
                  "want", [round(float(pr.grad[i]), 4) for i in top], "max|want|", float(pr.grad.abs().max()))
