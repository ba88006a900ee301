"""GPU: the tcgen05/TMEM GEMM pipeline of the 16-bit tensor-core mode against a plain torch matmul on the same rounded
operands (K-major gather tiles, MN-major feature-major tiles built by the SIMT loaders (mode 1), the same tiles fetched by
TMA tensor-map copies (mode 2): fp16 weights x fp16 operand, the forward configuration.  (bf16 x fp16 in ONE kind::f16
instruction is an illegal instruction on sm_100a, measured in round 2; the dW GEMMs convert their fp16 side instead.)"""
import pytest
import torch

from dl_biomass_b200 import _lib

pytestmark = pytest.mark.gpu


def _run(m_out, k, rows, mode, dev):
    g = torch.Generator().manual_seed(m_out * 7 + k * 3 + rows + mode)
    w = torch.randn(m_out, k, generator=g)
    b = torch.randn(rows, k, generator=g)
    w_bf = w.to(torch.float16).float()
    b_bf = b.to(torch.float16)
    want = (w_bf.double() @ b_bf.double().t()).float()            # [m_out, rows]
    tiles = (rows + 127) // 128
    ld = tiles * 128
    wd = w.to(dev)
    if mode == 0:
        bd = b_bf.to(dev).contiguous()
        ldb = 0
    else:
        bd = torch.zeros(k, ld, dtype=torch.float16, device=dev)
        bd[:, :rows] = b_bf.t().to(dev)
        ldb = ld
    zeros3 = torch.zeros(rows, 3, device=dev)
    out = torch.full((m_out, ld), float("nan"), device=dev)
    ws = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
    rc = _lib.lib().b2pn_tc_gemm_selftest(wd.data_ptr(), m_out, k, bd.data_ptr(), mode, rows, ldb, zeros3.data_ptr(),
                                          out.data_ptr(), ld, ws.data_ptr(), ws.numel(),
                                          torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "b2pn_tc_gemm_selftest")
    torch.cuda.synchronize()
    got = out[:, :rows].cpu()
    err = float((got - want).abs().max() / want.abs().max())
    return err, got, want


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("m_out,k,rows", [(64, 64, 128), (128, 128, 1024), (128, 131, 1000), (256, 128, 5000),
                                          (200, 259, 777), (1024, 512, 6000), (64, 8, 300)])
def test_tc_gemm_selftest(cuda_device, m_out, k, rows, mode):
    err, got, want = _run(m_out, k, rows, mode, cuda_device)
    assert err < 2e-3, (err, got[:2, :4], want[:2, :4])
