"""Device versions of the pieces around the training kernel (``-m gpu``): sort-based ROC-AUC (against the numpy
oracle, which tests/test_oracle_golden.py pins to sklearn, including the heavy-tie fixture), the loader's epoch shuffle
and fused id gather, the serving leftovers (exclusion of seen items, hit rate on device, predict with more factors than
the tensor-core kernel takes) and the ``use_amp`` flag."""
import numpy as np
import pandas as pd
import pytest
import torch

from oracle import cf_oracle as O
from tests import _golden as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


# ---- sorted ROC-AUC -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_pos,n_neg,ties", [(1, 1, False), (7, 5, False), (1000, 3000, False), (4096, 4096, True),
                                              (100_000, 50_000, True), (1_000_000, 1_000_000, False)])
def test_sorted_auc_equals_the_oracle_exactly(dev, n_pos, n_neg, ties):
    from torchrecsys_b200 import _lib
    rng = np.random.default_rng(n_pos + n_neg)
    pos = rng.normal(0.3, 1.0, n_pos).astype(np.float32)
    neg = rng.normal(0.0, 1.0, n_neg).astype(np.float32)
    if ties:  # a few distinct values only: long runs of equal scores across both classes, signed zeros included
        pos, neg = np.round(pos * 2) / 2, np.round(neg * 2) / 2
        pos[::7], neg[::5] = -0.0, 0.0
    got = float(_lib.sorted_auc(torch.from_numpy(pos).to(dev), torch.from_numpy(neg).to(dev)).item())
    if n_pos + n_neg <= 200_000:
        want = O.roc_auc(pos, neg)
    else:  # the oracle's python loop is slow there: the same statistic through numpy's searchsorted
        s = np.sort(np.concatenate([pos, neg]).astype(np.float64))
        lo, hi = np.searchsorted(s, pos.astype(np.float64), "left"), np.searchsorted(s, pos.astype(np.float64), "right")
        want = float(((lo + hi + 1) / 2.0).sum() - n_pos * (n_pos + 1) / 2.0) / (n_pos * n_neg)
    assert got == pytest.approx(want, rel=0, abs=1e-12)   # integer rank sum: exact up to the final double division


def test_sorted_auc_on_the_golden_tie_fixture_and_edge_cases(dev):
    from torchrecsys_b200 import _lib
    from torchrecsys_b200.evaluate.metrics import Metrics
    g = G.load("eval_pairwise")
    pos, neg = g["pos"].astype(np.float32).reshape(-1), g["neg"].astype(np.float32).reshape(-1)
    got = Metrics().roc_auc(torch.from_numpy(pos).to(dev), torch.from_numpy(neg).to(dev))
    assert got == pytest.approx(O.roc_auc(pos, neg), abs=1e-12)
    one = torch.ones(5, device=dev)
    assert float(_lib.sorted_auc(one, one).item()) == 0.5                       # all tied
    assert float(_lib.sorted_auc(one + 1, one).item()) == 1.0
    assert float(_lib.sorted_auc(one - 1, one).item()) == 0.0
    assert np.isnan(float(_lib.sorted_auc(one[:0], one).item()))                 # an empty class


# ---- loader: shuffle + gather ---------------------------------------------------------------------------------
def test_epoch_shuffle_is_a_seeded_permutation_and_gather_matches_indexing(dev):
    from torchrecsys_b200 import _lib
    for n in (0, 1, 2, 1000, 2049, 300_001):
        p = _lib.epoch_shuffle(7, n, dev)
        assert p.dtype == torch.int64 and torch.equal(torch.sort(p)[0], torch.arange(n, device=dev))
        assert torch.equal(p, _lib.epoch_shuffle(7, n, dev))                      # replayable
        if n > 1000:
            assert not torch.equal(p, _lib.epoch_shuffle(8, n, dev))
            assert float((p == torch.arange(n, device=dev)).float().mean()) < 0.01  # not the identity
    n = 5000
    cols = [torch.randint(0, 1 << 40, (n,), device=dev), torch.randint(0, 99, (n, 3), device=dev),
            torch.randint(0, 5, (n, 1), device=dev)]
    perm = _lib.epoch_shuffle(3, n, dev)
    for src, got in zip(cols, _lib.gather_rows(cols, perm)):
        assert torch.equal(got, src[perm])


# ---- model-level surface ----------------------------------------------------------------------------------------
def _frame(n_u=120, n_i=60, n_int=3000, seed=0):
    rng = np.random.default_rng(seed)
    return pd.DataFrame({"user": np.r_[np.arange(n_u), rng.integers(0, n_u, n_int - n_u)],
                         "item": np.r_[np.arange(n_i), rng.integers(0, n_i, n_int - n_i)]})


def _quiet(fn, *a, **k):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def test_evaluate_reports_the_sorted_roc_auc(dev):
    from torchrecsys.model import TorchRecSys
    torch.manual_seed(0)
    np.random.seed(0)
    model = _quiet(TorchRecSys, _frame(), "user", "item", n_factors=16, net_type="fm", use_cuda=True)
    _quiet(model.fit, torch.optim.SparseAdam(list(model.parameters()), lr=0.05), epochs=2, batch_size=256)
    _quiet(model.evaluate, batch_size=128, eval_metrics=["loss", "auc", "roc_auc"])
    assert set(model.last_eval) == {"loss", "auc", "roc_auc"}
    test = model._device_split("test_data")
    with torch.no_grad():
        pos = model.net.forward({"u": test["user"], "i": test["pos"]}, "u", "i").cpu().numpy()
        neg = model.net.forward({"u": test["user"], "i": test["neg"]}, "u", "i").cpu().numpy()
    assert model.last_eval["roc_auc"] == pytest.approx(O.roc_auc(pos, neg), abs=1e-9)


def test_predict_batch_can_leave_out_seen_items_and_hit_rate_runs_on_device(dev):
    from torchrecsys.evaluate.metrics import Metrics
    from torchrecsys.model import TorchRecSys
    torch.manual_seed(1)
    np.random.seed(1)
    model = _quiet(TorchRecSys, _frame(seed=1), "user", "item", n_factors=16, net_type="linear", use_cuda=True)
    _quiet(model.fit, torch.optim.Adagrad(model.parameters(), lr=0.05), epochs=1, batch_size=256)
    users = torch.arange(0, 120, 7)
    k = 10
    plain = model.predict_batch(users, 60)                      # the full ranking
    unseen = model.predict_batch(users, k, exclude_seen=True)
    tr = model.data_processor.train_data
    for row, u in enumerate(users.tolist()):
        seen = set(tr["pos_item_id"][tr["user_id"] == u].tolist())
        want = [i for i in plain[row].tolist() if i not in seen][:k]
        assert unseen[row].tolist()[:len(want)] == want and not (set(unseen[row].tolist()) & seen)
    # hit rate with everything on the device: held-out items of the test split vs the recommendations
    te = model.data_processor.test_data
    truth = torch.full((len(users), 8), -1, dtype=torch.int64)
    for row, u in enumerate(users.tolist()):
        items = te["pos_item_id"][te["user_id"] == u][:8]
        truth[row, :len(items)] = items
    m = Metrics()
    got = m.hit_rate(truth.to(dev), unseen.to(dev))
    want = m.hit_rate(truth, unseen)
    assert got == want and 0.0 <= got <= 1.0
    with pytest.raises(IndexError):
        model.predict(10 ** 6)
    with pytest.raises(IndexError):
        model.predict_batch(torch.tensor([0, -1]))


def test_predict_works_beyond_the_tensor_core_kernels_factor_limit(dev):
    """n_factors = 256 trains and evaluates; predict must rank too (exact fp32 path), as the reference does."""
    from torchrecsys.model import TorchRecSys
    torch.manual_seed(2)
    np.random.seed(2)
    model = _quiet(TorchRecSys, _frame(seed=2), "user", "item", n_factors=256, net_type="linear", use_cuda=True)
    top = model.predict(3, top_k=5)
    with torch.no_grad():
        items = torch.arange(model.n_items, device=dev)
        s = model.net.forward({"u": torch.full_like(items, 3), "i": items}, "u", "i").view(-1)
    assert top.tolist() == torch.sort(s, descending=True, stable=True)[1][:5].cpu().tolist()
    assert model.predict_batch(torch.tensor([3, 4]), 5).shape == (2, 5)


@pytest.mark.parametrize("net_type", ["linear", "fm"])
def test_use_amp_does_not_change_the_fp32_scorers(dev, net_type):
    """use_amp only concerns the MLP tower's GEMMs (always bf16 on the tensor cores here; fp16 autocast in the
    reference).  Linear / FM compute in fp32 with the flag on or off: bit-identical training."""
    from torchrecsys.model import TorchRecSys
    out = []
    for amp in (False, True):
        torch.manual_seed(3)
        np.random.seed(3)
        model = _quiet(TorchRecSys, _frame(seed=3), "user", "item", n_factors=16, net_type=net_type, use_cuda=True,
                       use_amp=amp)
        _quiet(model.fit, torch.optim.SparseAdam(list(model.parameters()), lr=0.01), epochs=2, batch_size=128)
        out.append({k: v.clone() for k, v in model.net.state_dict().items()})
    for k in out[0]:
        assert torch.equal(out[0][k], out[1][k]), k
