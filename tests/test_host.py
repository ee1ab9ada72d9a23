"""Host-side logic that needs no GPU: data preparation, the loader, module structure, optimizer
binding arithmetic.  Mirrors the structural assertions of the reference's own tests
(tests/test_model_and_features.py:52-131,145-185; tests/test_metrics.py:6-25)."""
import math

import numpy as np
import pandas as pd
import pytest
import torch

from torchrecsys.collaborative.mlp import MLP
from torchrecsys.dataset.dataset import FastDataLoader, ProcessData
from torchrecsys.evaluate.metrics import Metrics
from torchrecsys.model import TorchRecSys
from torchrecsys_b200 import _lib, engine

N_USERS, N_ITEMS, N_INT, N_CAT = 100, 50, 1000, 5


def frame(meta=None, seed=0):
    rng = np.random.default_rng(seed)
    df = pd.DataFrame({"user_id": np.r_[np.arange(N_USERS), rng.integers(0, N_USERS, N_INT - N_USERS)],
                       "item_id": np.r_[np.arange(N_ITEMS), rng.integers(0, N_ITEMS, N_INT - N_ITEMS)]})
    cat = rng.integers(0, N_CAT, N_ITEMS)
    if meta == "int":
        df["category_ids"] = cat[df.item_id]
    elif meta == "list":
        df["category_ids"] = [[int(cat[i]), int((cat[i] + 1) % N_CAT)] for i in df.item_id]
    elif meta == "str":
        df["category_ids"] = [str([int(cat[i])]) for i in df.item_id]
    return df, cat


@pytest.mark.parametrize("dyn", [False, True])
def test_process_data_keys(dyn):
    df, _ = frame()
    p = ProcessData(df, "user_id", "item_id", dynamic_neg_sampling=dyn)
    p.prepare_data()
    assert ("neg_item_id" in p.train_data) == (not dyn)
    assert p.train_data["user_id"].dtype == torch.int64
    assert p.config == {"num_users": N_USERS, "num_items": N_ITEMS, "num_metadata": {}}
    n_tr, n_te = p.train_data["user_id"].numel(), p.test_data["user_id"].numel()
    assert n_tr + n_te == N_INT and n_te == math.ceil(N_INT * (1 - 0.9) - 1e-9)


@pytest.mark.parametrize("form", ["int", "list", "str"])
@pytest.mark.parametrize("dyn", [False, True])
def test_process_data_metadata_forms(form, dyn):
    df, cat = frame(form)
    p = ProcessData(df, "user_id", "item_id", metadata_id_col=["category_ids"], dynamic_neg_sampling=dyn)
    p.prepare_data()
    tr = p.train_data
    assert tr["pos_metadata_id"].shape == (tr["user_id"].numel(), 1)
    assert np.array_equal(tr["pos_metadata_id"][:, 0].numpy(), cat[tr["pos_item_id"].numpy()])
    assert ("neg_metadata_id" in tr) == (not dyn)
    if not dyn:
        assert tr["neg_metadata_id"].numel() > 0
        assert np.array_equal(tr["neg_metadata_id"][:, 0].numpy(), cat[tr["neg_item_id"].numpy()])
    assert np.array_equal(p.item_meta[:, 0], cat)
    assert p.item_to_metadata_map[3]["category_ids"][0] == cat[3]
    assert p.config["num_metadata"]["category_ids"] >= N_CAT


def test_split_and_static_negatives_match_reference_rng_contract():
    """Same numpy draw and same sklearn split call as the reference (dataset.py:58-60, 240)."""
    from sklearn.model_selection import train_test_split
    df, _ = frame()
    np.random.seed(7)
    p = ProcessData(df, "user_id", "item_id", split_ratio=0.8)
    p.prepare_data()
    np.random.seed(7)
    neg = np.random.randint(low=0, high=N_ITEMS, size=N_INT)
    ref = df.assign(neg_item=neg)
    tr, te = train_test_split(ref, test_size=1 - 0.8, random_state=42)
    assert np.array_equal(p.train_data["user_id"].numpy(), tr.user_id.to_numpy())
    assert np.array_equal(p.train_data["neg_item_id"].numpy(), tr.neg_item.to_numpy())
    assert np.array_equal(p.test_data["pos_item_id"].numpy(), te.item_id.to_numpy())


def test_loader_dynamic_negatives_and_shapes():
    df, _ = frame("list")
    p = ProcessData(df, "user_id", "item_id", metadata_id_col=["category_ids"], dynamic_neg_sampling=True)
    p.prepare_data()
    loader = FastDataLoader(p.train_data, batch_size=32, shuffle=False, dynamic_neg_sampling=True,
                            n_items=N_ITEMS, item_to_metadata_map=p.item_to_metadata_map,
                            metadata_id_cols=["category_ids"])
    seen = 0
    for batch in loader:
        assert batch["pos_item_id"].shape == batch["neg_item_id"].shape
        assert (batch["pos_item_id"] != batch["neg_item_id"]).all()
        assert batch["pos_metadata_id"].shape[0] == batch["neg_metadata_id"].shape[0]
        assert batch["neg_metadata_id"].dim() in (2, 3)
        assert np.array_equal(batch["neg_metadata_id"][:, 0].numpy(),
                              p.item_meta[batch["neg_item_id"].numpy(), 0])
        seen += batch["user_id"].numel()
    assert seen == p.train_data["user_id"].numel() and len(loader) == math.ceil(seen / 32)
    with pytest.raises(ValueError):
        FastDataLoader(p.train_data, dynamic_neg_sampling=True)


def test_loader_shuffle_is_a_permutation_and_empty_is_fine():
    df, _ = frame()
    p = ProcessData(df, "user_id", "item_id", split_ratio=1.0)
    p.prepare_data()
    assert p.test_data["user_id"].numel() == 0
    assert list(FastDataLoader(p.test_data, batch_size=8)) == []
    got = torch.cat([b["user_id"] for b in FastDataLoader(p.train_data, batch_size=64, shuffle=True)])
    assert torch.equal(got.sort()[0], p.train_data["user_id"].sort()[0])


def test_mlp_structure():
    m = MLP(n_users=10, n_items=7, n_metadata={}, n_factors=16, use_metadata=False, hidden_layers=[64, 32])
    assert [fc.out_features for fc in m.fcs] == [64, 32] and m.fcs[0].in_features == 32
    assert m.output_layer.in_features == 32 and len(m.bns) == 2
    names = [n for n, _ in m.named_parameters()]
    assert names[:2] == ["user.weight", "item.weight"] and names[-1] == "output_layer.bias"
    m2 = MLP(n_users=10, n_items=7, n_metadata={"c": 3}, n_factors=16, use_batch_norm=False)
    assert not hasattr(m2, "bns") and m2.hidden_layers == [1024, 128] and m2.input_shape == 48


def test_parameter_names_match_reference_state_dict():
    df, _ = frame("int")
    lin = TorchRecSys(df, "user_id", "item_id", n_factors=8, net_type="linear", metadata_id_col=["category_ids"])
    assert [n for n, _ in lin.named_parameters()] == [
        "net.metadata.0.weight", "net.user.weight", "net.item.weight", "net.user_bias.weight", "net.item_bias.weight"]
    fm = TorchRecSys(df, "user_id", "item_id", n_factors=8, net_type="fm", metadata_id_col=["category_ids"])
    assert [n for n, _ in fm.named_parameters()] == [
        "net.user.weight", "net.item.weight", "net.linear_user.weight", "net.linear_item.weight",
        "net.metadata.0.weight", "net.linear_metadata.0.weight"]
    assert lin.net.user.weight.std().item() == pytest.approx(1 / 8, rel=0.2)  # N(0, 1/D)
    assert float(lin.net.item_bias.weight.abs().sum()) == 0.0
    with pytest.raises(AssertionError):
        TorchRecSys(df, "user_id", "item_id", net_type="ease")


def test_no_cpu_fallback():
    df, _ = frame()
    model = TorchRecSys(df, "user_id", "item_id", n_factors=8, use_cuda=False)
    opt = torch.optim.SparseAdam(list(model.parameters()))
    for call in (lambda: model.fit(opt, epochs=1), lambda: model.evaluate(), lambda: model.predict(0)):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()
    batch = {"user_id": torch.zeros(4, dtype=torch.long), "pos_item_id": torch.zeros(4, dtype=torch.long)}
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.net.forward(batch, "user_id", "pos_item_id")


def test_hit_rate_known_answer():
    y_hat = torch.tensor([[0, 1, -1], [1, 2, -1], [1, 2, 3], [0, 1, -1]])
    y_pred = torch.tensor([[2, 3], [0, 2], [1, 2], [0, 3]])
    assert Metrics().hit_rate(y_hat=y_hat, y_pred=y_pred) == 3 / 4


def test_step_scales_follow_torch_formulas():
    b = engine.OptBinding(_lib.OPT_SPARSE_ADAM, ("exp_avg", "exp_avg_sq"), 1e-3, 0.9, 0.999, 1e-8, step0=4)
    s = engine.step_scales(b, 2)
    assert s[0] == 1e-3 * math.sqrt(1 - 0.999 ** 5) / (1 - 0.9 ** 5)
    b = engine.OptBinding(_lib.OPT_ADAGRAD, ("sum", None), 0.5, lr_decay=0.1, step0=0)
    assert engine.step_scales(b, 3) == [0.5, 0.5 / 1.1, 0.5 / 1.2]


def test_canonical_meta_accepts_reference_layouts():
    from torchrecsys_b200.collaborative._base import canonical_meta
    bag = torch.tensor([[3, 9, 0], [4, 0, 0]])
    assert canonical_meta(bag, 1).tolist() == [[3], [4]]                      # (B, L), one feature
    assert canonical_meta(torch.stack([bag, bag + 1], 1), 2).tolist() == [[3, 4], [4, 5]]  # (B, F, L)
    assert canonical_meta(torch.tensor([[1, 2], [3, 4]]), 2).tolist() == [[1, 2], [3, 4]]   # canonical
    assert canonical_meta(None, 0) is None


def test_precision_recall_at_k_matches_the_reference_loop():
    """helper/evaluate.py:53-76 restated as the per-user set loop it is, against the tensor version."""
    import numpy as np
    import torch
    from torchrecsys_b200.evaluate.metrics import Metrics
    rng = np.random.default_rng(3)
    n_users, n_items, k = 40, 60, 7
    tu, ti = rng.integers(0, n_users - 5, 300), rng.integers(0, n_items, 300)  # the last 5 users have no truth
    scored = rng.permutation(n_users)[:25]
    rec = np.stack([rng.permutation(n_items)[:k] for _ in scored])
    rec[3, -2:] = -1  # padded row
    want_p, want_r = [], []
    for row, u in enumerate(scored):
        truth = set(ti[tu == u].tolist())
        if truth:
            n = len(truth & set(rec[row].tolist()))
            want_p.append(n / k)
            want_r.append(n / len(truth))
    p, r = Metrics().precision_recall_at_k(torch.from_numpy(scored), torch.from_numpy(rec), tu, ti)
    assert abs(p - np.mean(want_p)) < 1e-12 and abs(r - np.mean(want_r)) < 1e-12


def test_long_epochs_are_cut_into_runs_of_whole_steps(monkeypatch):
    """engine._Chunked.run: consecutive slices of MAX_STEPS_PER_CALL * batch samples, losses concatenated."""
    import torch
    from torchrecsys_b200 import engine

    class Fake(engine._Chunked):
        def __init__(self):
            self.calls = []

        def _run(self, samples, batch_size):
            n = samples["user"].shape[0]
            self.calls.append((int(samples["user"][0]), n, tuple(samples["meta"].shape)))
            return torch.full((-(-n // batch_size),), float(len(self.calls)))

    monkeypatch.setattr(engine, "MAX_STEPS_PER_CALL", 4)
    n, B = 4 * 3 * 2 + 5, 3  # two full runs of 4 steps and a rest of 5 samples (2 steps)
    smp = {"user": torch.arange(n), "meta": torch.arange(2 * n).view(n, 2)}
    f = Fake()
    loss = f.run(smp, B)
    assert f.calls == [(0, 12, (12, 2)), (12, 12, (12, 2)), (24, 5, (5, 2))]
    assert loss.tolist() == [1.0] * 4 + [2.0] * 4 + [3.0] * 2
    g = Fake()
    g.run({k: v[:12] for k, v in smp.items()}, B)
    assert len(g.calls) == 1


def test_plan_launch_count_follows_the_batch_size():
    from types import SimpleNamespace as NS
    from torchrecsys_b200 import engine, _lib
    m = NS(n_meta=1, net=_lib.NET_FM, user=NS(n_rows=1_000_000), item=NS(n_rows=200_000), meta=[NS(n_rows=100)])
    assert engine.plan_launches(m, 8192, 200) == 1            # one CTA per (step, id space)
    assert engine.plan_launches(m, 16384, 200) == 2 * (3 + 3 + 1) + 2 + 1 + 2  # tiled sort + items + flags


def test_batchnorm_refuses_a_training_batch_of_one_row_like_torch():
    """torch.nn.BatchNorm1d raises on a [1, C] batch in train mode (what the reference's MLP would hit on a short last
    batch); the fused MLP runner refuses the epoch with torch's message instead of normalising one row by itself."""
    from torchrecsys_b200.engine import MlpEpochRunner
    bn = torch.nn.BatchNorm1d(8).train()
    with pytest.raises(ValueError) as torch_err:
        bn(torch.zeros(1, 8))
    with pytest.raises(ValueError) as ours:
        MlpEpochRunner.check_batchnorm_rows(1025, 512, 8)
    assert str(ours.value) == str(torch_err.value)
    with pytest.raises(ValueError):
        MlpEpochRunner.check_batchnorm_rows(7, 1, 8)
    MlpEpochRunner.check_batchnorm_rows(1024, 512, 8)      # full batches
    MlpEpochRunner.check_batchnorm_rows(1026, 512, 8)      # a last batch of two rows is fine
    MlpEpochRunner.check_batchnorm_rows(0, 512, 8)
