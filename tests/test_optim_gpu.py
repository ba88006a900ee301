"""GPU: optim.FlatAdam (one b2pn_adam_step launch over the parameter arena) against torch.optim.Adam, the optimiser of
/root/reference/main.py:84 (lr 1.8e-3, weight_decay 8e-5, L2-in-gradient form), and the thread re-entrancy of the
set-abstraction entry points."""
import threading

import pytest
import torch

from dl_biomass_b200.data import Batch, synthetic_clouds

pytestmark = pytest.mark.gpu


def test_flat_adam_matches_torch_adam(cuda_device):
    from dl_biomass_b200.optim import FlatAdam, ParamArena
    from dl_biomass_b200.train import ADAM_LR, ADAM_WEIGHT_DECAY
    torch.manual_seed(0)

    def make():
        torch.manual_seed(1)
        # (no BatchNorm here: a bias in front of one has a zero gradient in theory and rounding noise in practice, which
        #  Adam's normalisation turns into O(lr) parameter differences between ANY two implementations)
        m = torch.nn.Sequential(torch.nn.Linear(37, 53), torch.nn.Tanh(), torch.nn.Linear(53, 4)).to(cuda_device)
        return m

    a, b = make(), make()
    ref_opt = torch.optim.Adam(a.parameters(), lr=ADAM_LR, weight_decay=ADAM_WEIGHT_DECAY)
    arena = ParamArena(b, buckets=(("2.",), ("0.", "1.")))
    opt = FlatAdam(arena, lr=ADAM_LR, weight_decay=ADAM_WEIGHT_DECAY)
    assert arena.intact() and arena.numel % 4 == 0
    g = torch.Generator(device="cpu").manual_seed(3)
    for step in range(12):
        x = torch.randn(16, 37, generator=g).to(cuda_device)
        for m, o in ((a, ref_opt), (b, opt)):
            o.zero_grad(set_to_none=True)
            m(x).square().mean().backward()
            o.step()
        for (k, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
            err = float((p.detach() - q.detach()).abs().max() / p.detach().abs().max())
            assert err < 2e-6 * (step + 1), (step, k, err)
    assert opt.steps_taken == 12
    sd = opt.state_dict()
    assert sd["step"] == 12 and set(sd["state"]) == {n for n, _ in b.named_parameters()}
    st = ref_opt.state[list(a.parameters())[0]]
    torch.testing.assert_close(sd["state"]["0.weight"]["exp_avg_sq"], st["exp_avg_sq"], rtol=1e-4, atol=1e-12)
    torch.testing.assert_close(sd["state"]["0.weight"]["exp_avg"], st["exp_avg"], rtol=1e-4, atol=1e-9)
    # a CUDA-graph replay is a real optimiser step (the step counter lives on the device)
    opt2 = FlatAdam(ParamArena(make()), lr=ADAM_LR)
    opt2.arena.flat_grads.fill_(0.5)
    s = torch.cuda.Stream(cuda_device)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        opt2.step(collected=True)
    torch.cuda.current_stream().wait_stream(s)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        opt2.step(collected=True)
    gr.replay()
    gr.replay()
    assert opt2.steps_taken == 3


def test_sa_calls_are_reentrant_across_threads(cuda_device):
    """Thread-per-GPU callers (/root/reference/main.py:140): two host threads run set-abstraction forwards on their own
    streams with DIFFERENT per-call SM caps at the same time; each gets the result of its serial run.  (There is no
    process-wide setting left to trample on: the cap is a field of b2pn_sa_args, thread-local on the Python side.)"""
    from dl_biomass_b200 import sa
    from dl_biomass_b200.pointnet2_regressor import Net
    torch.manual_seed(4)
    net = Net(1, "ReLU", 0, 0.0, precision="bf16").to(cuda_device).set_random_start(False).eval()
    batches = [Batch.from_data_list(synthetic_clouds(60 + i, 3, 700, 1, True)).to(cuda_device) for i in range(2)]
    with torch.no_grad():
        want = [net(b).clone() for b in batches]
    torch.cuda.synchronize()
    got, caps, errs = [None, None], [None, None], []

    def work(i, cap):
        try:
            torch.cuda.set_device(cuda_device)
            st = torch.cuda.Stream(cuda_device)
            with torch.cuda.stream(st), torch.no_grad():
                sa.set_sm_limit(cap)
                outs = [net(batches[i]) for _ in range(5)]
                caps[i] = sa.launch_options()[0]
            st.synchronize()
            got[i] = outs[-1]
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=work, args=(0, 40)), threading.Thread(target=work, args=(1, 100))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    assert caps == [40, 100] and sa.launch_options()[0] == 0      # per-thread options, the main thread's untouched
    for g, w in zip(got, want):
        assert torch.equal(g, w)


def test_backward_kernels_write_straight_into_the_parameter_arena(cuda_device):
    """With a ParamArena (make_optimizer(model)) every parameter gradient of the network lands in the flat gradient buffer
    directly: after backward ``p.grad`` IS a view of the arena (autograd took the returned alias over without a copy), so
    the optimiser step and the data-parallel all-reduce need no flattening copies."""
    from dl_biomass_b200.optim import ParamArena
    from dl_biomass_b200.pointnet2_regressor import Net
    from dl_biomass_b200.train import forward_backward, make_optimizer
    torch.manual_seed(3)
    net = Net(1, "ReLU", 0, 0.5, precision="bf16").to(cuda_device).set_random_start(False)
    net.train()
    opt = make_optimizer(net)
    arena = ParamArena.of(net)
    b = Batch.from_data_list(synthetic_clouds(5, 4, 600, 1, True)).to(cuda_device)
    forward_backward(net, opt, b)
    torch.cuda.synchronize()
    for n, p in net.named_parameters():
        assert p.grad is not None, n
        assert p.grad.data_ptr() == arena.grad_views[p].data_ptr(), f"{n}: gradient was copied instead of written in place"
    assert float(arena.flat_grads.abs().max()) > 0
