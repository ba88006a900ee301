"""CPU: checkpoint interchange with the reference's spellings and the evaluation metrics against
scikit-learn, the library the reference calls."""
import math

import numpy as np
import pytest
import torch

from dl_biomass_b200 import checkpoint, metrics
from dl_biomass_b200.pointnet2_regressor import Net
from oracle import ref


def test_reference_state_dict_round_trip():
    torch.manual_seed(0)
    src = ref.seeded_init_(ref.NetRef(1, "ReLU", 0, 0.5), seed=3)   # PyG-shaped keys (oracle/ref.py)
    for wrapper in (False, True):
        for dp in (False, True):
            sd = {}
            for k, v in src.state_dict().items():
                if wrapper:
                    k = checkpoint._NORM_PLAIN.sub(r"\1.module.\2", "." + k)[1:]
                sd[("module." + k) if dp else k] = v.clone()
            if wrapper:
                assert any(".norms.0.module.running_mean" in k for k in sd)
            net = Net(1, "ReLU", 0, 0.5)
            res = checkpoint.load_reference_state_dict(net, sd)
            assert not res.missing_keys and not res.unexpected_keys
            for k, v in src.state_dict().items():
                assert torch.equal(net.state_dict()[k], v), k
            back = checkpoint.reference_state_dict(net, pyg_norm_wrapper=wrapper, data_parallel=dp)
            assert list(back.keys()) == list(sd.keys())
            assert all(torch.equal(back[k], sd[k]) for k in sd)
    # idempotent
    once = checkpoint.normalise_state_dict(sd)
    assert list(checkpoint.normalise_state_dict(once).keys()) == list(once.keys())


def test_metrics_match_scikit_learn():
    sk = pytest.importorskip("sklearn.metrics")
    rng = np.random.default_rng(2)
    obs = rng.uniform(0.5, 40.0, size=(57, 4))
    pred = obs * rng.normal(1.0, 0.2, size=obs.shape) + rng.normal(0.0, 0.5, size=obs.shape)
    got = metrics.regression_metrics(torch.from_numpy(obs).float(), torch.from_numpy(pred).float())
    o32, p32 = obs.astype(np.float32).astype(np.float64), pred.astype(np.float32).astype(np.float64)
    cols = {name: (o32[:, i], p32[:, i]) for i, name in enumerate(metrics.COMPONENTS)}
    cols["tree_btphr"] = (o32.sum(1), p32.sum(1))
    for name, (o, p) in cols.items():
        assert math.isclose(got[name]["r2"], sk.r2_score(o, p), rel_tol=1e-9, abs_tol=1e-12)
        assert math.isclose(got[name]["rmse"], math.sqrt(sk.mean_squared_error(o, p)), rel_tol=1e-9)
        assert math.isclose(got[name]["mape"], sk.mean_absolute_percentage_error(o, p), rel_tol=1e-9)


def test_evaluate_loop_with_a_stub_model():
    """metrics.evaluate: eval mode inside, training flag restored, chunked batches == one batch, table == scikit-learn."""
    import numpy as np
    import torch
    from sklearn import metrics as skm
    from dl_biomass_b200.data import Batch, synthetic_clouds
    from dl_biomass_b200.metrics import evaluate

    class Stub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.seen_training = []

        def forward(self, batch):
            self.seen_training.append(self.training)
            y = batch.y.reshape(-1, 4)
            return y * torch.tensor([1.1, 0.9, 1.0, 1.05]) + 0.01

    clouds = synthetic_clouds(900, 10, 64)
    chunks = [Batch.from_data_list(clouds[:4]), Batch.from_data_list(clouds[4:7]), Batch.from_data_list(clouds[7:])]
    m = Stub().train()
    table, (obs, pred) = evaluate(m, chunks, return_predictions=True)
    assert m.training and m.seen_training == [False, False, False]
    assert table == evaluate(m, [Batch.from_data_list(clouds)])
    o, p = obs.numpy().astype(np.float64), pred.numpy().astype(np.float64)
    assert abs(table["wood_btphr"]["r2"] - skm.r2_score(o[:, 3], p[:, 3])) < 1e-9
    assert abs(table["tree_btphr"]["rmse"] - np.sqrt(skm.mean_squared_error(o.sum(1), p.sum(1)))) < 1e-9
    assert abs(table["bark_btphr"]["mape"] - skm.mean_absolute_percentage_error(o[:, 0], p[:, 0])) < 1e-9
    import pytest
    with pytest.raises(ValueError):
        evaluate(m, [])
