"""GPU parity of the fused regression head (csrc/head.cu) against the plain layer-by-layer evaluation of the same
parameters (torch on the GPU, fp32): forward, every gradient, BatchNorm running statistics, eval mode, and dropout
(the keep-masks the kernel drew are replayed in torch)."""
import pytest
import torch
import torch.nn.functional as F

from dl_biomass_b200 import head
from dl_biomass_b200.pointnet2_regressor import MLP

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach(), b.detach()
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12))


def _pair(chans, p, dev, seed=3):
    torch.manual_seed(seed)
    m = MLP(chans, act=None, dropout=p).to(dev)
    for n in m.norms:  # non-trivial affine + running stats
        n.weight.data.uniform_(0.5, 1.5)
        n.bias.data.uniform_(-0.5, 0.5)
        n.running_mean.uniform_(-0.2, 0.2)
        n.running_var.uniform_(0.5, 2.0)
    import copy
    return m, copy.deepcopy(m)


def _torch_forward(m, x, masks, p):
    """MLP.forward with the dropout masks given explicitly."""
    h = m.lins[0](x)
    for i, (lin, norm) in enumerate(zip(m.lins[1:], m.norms)):
        h = norm(h)
        if masks is not None:
            h = h * masks[i].float() / (1.0 - p)
        h = lin(h)
    return h


@pytest.mark.parametrize("B,chans,train", [(12, [1024, 128, 128, 4], True), (12, [1024, 128, 128, 4], False),
                                           (5, [300, 64, 96, 3], True), (32, [2048, 256, 256, 4], True),
                                           (1, [64, 32, 32, 4], False)])
def test_head_matches_torch(cuda_device, B, chans, train):
    m, mr = _pair(chans, 0.0, cuda_device)
    m.train(train)
    mr.train(train)
    g = torch.Generator().manual_seed(B)
    x = torch.randn(B, chans[0], generator=g).to(cuda_device).requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    counter = torch.zeros((), dtype=torch.int64, device=cuda_device)
    assert head.supported(m, x)
    out = head.head_apply(m, x, counter, 11)
    want = mr(xr)
    gout = torch.randn(B, chans[3], generator=g).to(cuda_device)
    out.backward(gout)
    want.backward(gout)
    torch.cuda.synchronize()
    assert rel(out, want) < 2e-5
    assert rel(x.grad, xr.grad) < 2e-4
    for (k, p_), (_, pr) in zip(m.named_parameters(), mr.named_parameters()):
        if train and k in ("lins.0.bias", "lins.1.bias", "norms.0.bias"):
            # additive constants in front of a train-mode BatchNorm (norms.0.bias reaches it through the linear,
            # activation-free lins.1): their gradient is zero up to rounding
            assert float(p_.grad.abs().max()) < 1e-3 * float(gout.abs().sum())
            continue
        assert rel(p_.grad, pr.grad) < 5e-4, k
    for (k, v), (_, vr) in zip(m.named_buffers(), mr.named_buffers()):
        assert rel(v.float(), vr.float()) < 1e-5, k


def test_head_dropout_masks_replayed(cuda_device):
    p = 0.5
    chans = [1024, 128, 128, 4]
    m, mr = _pair(chans, p, cuda_device)
    m.train()
    mr.train()
    x = torch.randn(12, 1024, generator=torch.Generator().manual_seed(1)).to(cuda_device).requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    counter = torch.zeros((), dtype=torch.int64, device=cuda_device)
    out = head.head_apply(m, x, counter, 5)
    saved = out.grad_fn.saved_tensors
    masks = (saved[3], saved[4])
    keep = float(torch.cat([mk.float().flatten() for mk in masks]).mean())
    assert 0.42 < keep < 0.58
    want = _torch_forward(mr, xr, masks, p)
    gout = torch.randn(12, 4, generator=torch.Generator().manual_seed(2)).to(cuda_device)
    out.backward(gout)
    want.backward(gout)
    assert rel(out, want) < 2e-5 and rel(x.grad, xr.grad) < 2e-4
    for (k, p_), (_, pr) in zip(m.named_parameters(), mr.named_parameters()):
        if k in ("lins.0.bias", "lins.1.bias"):
            continue
        assert rel(p_.grad, pr.grad) < 5e-4, k
    # (with dropout between them norms.0.bias is no longer a constant shift, so it is compared above)
    # the counter advanced: the next call draws different noise; resetting it reproduces the first call
    assert int(counter.item()) == 1
    out2 = head.head_apply(m, x.detach(), counter, 5)
    assert rel(out2, out.detach()) > 1e-3
    counter.zero_()
    m2, _ = _pair(chans, p, cuda_device)
    m2.train()
    out3 = head.head_apply(m2, x.detach(), counter, 5)
    assert torch.equal(out3, out.detach())


def test_head_falls_back_when_unsupported(cuda_device):
    m = MLP([64, 32, 32, 4], act=None).to(cuda_device)
    assert not head.supported(m, torch.zeros(33, 64, device=cuda_device))          # more than 32 clouds
    assert not head.supported(MLP([64, 32, 32, 4], act="relu").to(cuda_device), torch.zeros(4, 64, device=cuda_device))


def test_weighted_mse_loss_matches_reference_formula(cuda_device):
    """train.weighted_mse_loss == the four F.mse_loss terms of /root/reference/main.py:157-169, value and gradient."""
    import torch.nn.functional as F
    from dl_biomass_b200.train import LOSS_WEIGHTS, weighted_mse_loss
    g = torch.Generator().manual_seed(4)
    for B in (1, 12, 37):
        outs = (torch.randn(B, 4, generator=g) * 3).to(cuda_device).requires_grad_(True)
        y = (torch.rand(4 * B, generator=g) * 40).to(cuda_device)          # flat, as PyG collates it
        loss = weighted_mse_loss(outs, y)
        (loss * 1.5).backward()
        o = outs.detach().cpu().double().requires_grad_(True)
        yy = y.cpu().double().reshape(B, 4)
        want = sum(F.mse_loss(yy[:, c], o[:, c]) * LOSS_WEIGHTS[c] for c in range(4))
        (want * 1.5).backward()
        assert abs(float(loss.detach()) - float(want.detach())) <= 1e-5 * abs(float(want.detach()))
        assert torch.allclose(outs.grad.cpu().double(), o.grad, rtol=1e-5, atol=1e-7)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        weighted_mse_loss(torch.zeros(2, 4), torch.zeros(8))
