"""GPU parity of the WHOLE network (forward + backward of /root/reference/pointnet2_regressor.py:52-58 under the loss of
main.py:157-169) on the HEADLINE shape of BASELINE.json configs[1]: 12 clouds x 10 000 points, fixed and ragged.

Truth is the CPU oracle evaluated in float64.  Tolerances (north_star): fp32 mode 1e-4 relative on outputs and gradients,
bf16 mode 2e-2 on the regression outputs.  A gradient tensor is held to 1e-4 unless the oracle's OWN float32 evaluation
(the precision the reference runs in) is further than that from the float64 truth: such a tensor is ill-conditioned at
float32 in ANY implementation, and it is then held to twice the float32 oracle's error instead (printed)."""
import pytest
import torch

from oracle import ref
from dl_biomass_b200.data import Batch, synthetic_clouds
from dl_biomass_b200.pointnet2_regressor import Net

pytestmark = pytest.mark.gpu


def _record(msg):
    """Measured margins go to stdout and, as evidence for profiles/, to gpurun_out/r02_headline_parity.log."""
    import os
    print(msg)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
        with open(os.path.join(root, "gpurun_out", "r02_headline_parity.log"), "a") as f:
            f.write(msg + "\n")
    except OSError:
        pass


def rel_err(got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return float((got - want).abs().max() / want.abs().max().clamp_min(1e-300))


def _oracle(b, dtype):
    net = ref.seeded_init_(ref.NetRef(1, "ReLU", 0, 0.0), seed=7).to(dtype)
    net.train()
    data = type("D", (), {})()
    data.x, data.pos, data.batch, data.ptr = b.x.to(dtype), b.pos.to(dtype), b.batch, b.ptr
    out = net(data)
    loss = ref.weighted_mse(out, b.y.to(dtype))
    loss.backward()
    return net, out.detach(), {k: p.grad.detach() for k, p in net.named_parameters()}


@pytest.fixture(scope="module", params=[False, True], ids=["fixed", "ragged"])
def headline(request):
    b = Batch.from_data_list(synthetic_clouds(1234, 12, 10000, 1, request.param))
    n64, out64, g64 = _oracle(b, torch.float64)
    n32, out32, g32 = _oracle(b, torch.float32)
    return b, n32, out64, g64, out32, g32


def _gpu(b, net32, precision, dev):
    net = Net(1, "ReLU", 0, 0.0, precision=precision)
    # the same initial state the oracle started from (net32 has already taken its training-mode step: its BatchNorm
    # running statistics are what the buffers of `net` must equal AFTER this forward pass)
    net.load_state_dict(ref.seeded_init_(ref.NetRef(1, "ReLU", 0, 0.0), seed=7).state_dict())
    net = net.to(dev).set_random_start(False)
    net.train()
    out = net(b.to(dev))
    loss = ref.weighted_mse(out, b.y.to(dev))
    loss.backward()
    torch.cuda.synchronize()
    return net, out.detach(), {k: p.grad.detach() for k, p in net.named_parameters()}


@pytest.mark.timeout(900)
def test_net_fp32_headline_shape(cuda_device, headline):
    b, n32, out64, g64, out32, g32 = headline
    net, out, g = _gpu(b, n32, "fp32", cuda_device)
    e_out = rel_err(out, out64)
    _record(f"fp32 12x{b.cloud_sizes[0]}.. out rel err vs f64 oracle: {e_out:.3e} (f32 oracle: {rel_err(out32, out64):.3e})")
    assert e_out < 1e-4
    scale = max(float(v.abs().max()) for v in g64.values())
    worst, relaxed = 0.0, []
    for k, want in g64.items():
        if float(want.abs().max()) < 1e-7 * scale:   # biases in front of a BatchNorm: exactly 0 in theory
            continue
        e, e_ref = rel_err(g[k], want), rel_err(g32[k], want)
        bound = 1e-4 if e_ref <= 1e-4 else 2.0 * e_ref
        if e_ref > 1e-4:
            relaxed.append((k, e, e_ref))
        worst = max(worst, e)
        assert e <= bound, (k, e, e_ref)
    _record(f"fp32 12x{b.cloud_sizes[0]}.. worst grad rel err vs f64 oracle: {worst:.3e}; tensors ill-conditioned at f32 "
            f"(f32 oracle itself > 1e-4; (name, this repo, f32 oracle)): {[(k, f'{e:.2e}', f'{r:.2e}') for k, e, r in relaxed]}")
    for (k, v), (_, vr) in zip(net.named_buffers(), n32.named_buffers()):
        assert rel_err(v.float(), vr.float()) < 1e-4, k


@pytest.mark.timeout(900)
def test_net_bf16_headline_shape(cuda_device, headline):
    b, n32, out64, g64, out32, g32 = headline
    net, out, g = _gpu(b, n32, "bf16", cuda_device)
    e_out = rel_err(out, out64)
    _record(f"16-bit mode 12x{b.cloud_sizes[0]}.. out rel err vs f64 oracle: {e_out:.3e}  (bound 2e-2, margin {2e-2 / max(e_out, 1e-30):.2f}x)")
    assert e_out < 2e-2
    cos, rels = [], {}
    for k, want in g64.items():
        a, r = g[k].double().cpu().flatten(), want.flatten()
        assert torch.isfinite(a).all(), k
        if float(r.abs().max()) < 1e-7 * max(float(v.abs().max()) for v in g64.values()):
            continue
        cos.append(float((a @ r) / (a.norm() * r.norm()).clamp_min(1e-300)))
        rels[k] = float((a - r).norm() / r.norm())
    worst = max(rels, key=rels.get)
    _record(f"16-bit mode 12x{b.cloud_sizes[0]}.. grads vs f64 oracle: cosine min {min(cos):.4f} mean {sum(cos) / len(cos):.4f}; "
            f"worst relative L2 error {rels[worst]:.3e} ({worst})")
    # bf16 activations flip ReLU masks / arg-max winners of near-ties, so gradients agree in direction, not to 2e-2
    assert min(cos) > 0.9 and sum(cos) / len(cos) > 0.98, (min(cos), rels)


def test_reference_spelled_checkpoint_and_evaluation_on_the_device(cuda_device):
    """SURVEY 8(f) rows f3 / f4 on the GPU: a state_dict in the reference's spelling (`module.` prefix of its DataParallel
    wrapper, PyG >= 2.1 `norms.i.module.*`; /root/reference/main.py:140,245, testing_model.py:30-37) loads into the B200
    Net on the device and reproduces the oracle's evaluation-mode outputs; `metrics.evaluate` over device batches gives
    the oracle's metric table, whether the test set goes through as one batch or in chunks."""
    from dl_biomass_b200 import checkpoint, metrics
    netr = ref.seeded_init_(ref.NetRef(1, "ReLU", 0, 0.5), seed=3)
    netr.train()
    with torch.no_grad():   # non-trivial running statistics
        for i in range(2):
            netr(Batch.from_data_list(synthetic_clouds(300 + i, 5, 700, 1, True)))
    netr.eval()
    sd = {}
    for k, v in netr.state_dict().items():
        k = checkpoint._NORM_PLAIN.sub(r"\1.module.\2", "." + k)[1:]
        sd["module." + k] = v.clone()
    assert any(".norms.0.module.running_mean" in k for k in sd) and all(k.startswith("module.") for k in sd)
    clouds = synthetic_clouds(777, 9, 900, 1, True)
    whole = Batch.from_data_list(clouds)
    with torch.no_grad():
        want = netr(whole)
    want_table = metrics.regression_metrics(whole.y.reshape(-1, 4), want)
    for precision, tol in (("fp32", 1e-4), ("bf16", 2e-2)):
        net = Net(1, "ReLU", 0, 0.5, precision=precision).to(cuda_device).set_random_start(False)
        res = checkpoint.load_reference_state_dict(net, sd)
        assert not res.missing_keys and not res.unexpected_keys
        chunks = [Batch.from_data_list(clouds[:4]).to(cuda_device), Batch.from_data_list(clouds[4:]).to(cuda_device)]
        table, (obs, pred) = metrics.evaluate(net, chunks, return_predictions=True)
        assert pred.is_cuda and net.training          # evaluate() ran in eval mode and restored the training flag
        e = rel_err(pred, want)
        _record(f"{precision} eval through a reference-spelled checkpoint, 9 clouds in 2 chunks: out rel err {e:.3e}")
        assert e < tol
        one, _ = metrics.evaluate(net, [whole.to(cuda_device)], return_predictions=True)
        for name in want_table:
            for m in ("r2", "rmse", "mape"):
                assert abs(table[name][m] - want_table[name][m]) <= 5e-2 * max(1.0, abs(want_table[name][m])), (name, m)
                assert abs(one[name][m] - table[name][m]) <= 1e-3 * max(1.0, abs(table[name][m])), (name, m)
        # and back out in the reference's spelling
        back = checkpoint.reference_state_dict(net, pyg_norm_wrapper=True, data_parallel=True)
        assert set(back) == set(sd) and all(torch.equal(back[k].cpu(), sd[k]) for k in sd)
