"""CPU check of tests/_mlp_emul.py (the bf16-rounding restatement the GPU tests compare the kernels with):
it must track the fp32 oracle at the bf16 error level, and with the roundings disabled it IS the oracle."""
import numpy as np
import pytest
import torch

from oracle import cf_oracle as O
from tests import _golden as G
from tests import _mlp_emul as E


@pytest.mark.parametrize("name", G.names("train_mlp_"))
def test_emulation_without_rounding_is_the_fp32_oracle(name, monkeypatch):
    g = G.load(name)
    monkeypatch.setattr(E, "bf", lambda x: x)
    p = {k: torch.from_numpy(np.asarray(v)).clone() for k, v in G.section(g, "init").items()}
    batch = {k: torch.from_numpy(v) for k, v in G.batch_at(g, 0).items()}
    loss, dense, sparse = E.mlp_grads_bf16(p, batch)
    want_loss, want_sparse, want_dense = O.mlp_grads(G.section(g, "init"), G.batch_at(g, 0))
    assert abs(loss - float(want_loss)) < 1e-5
    for k, w in want_dense.items():
        scale = np.abs(w).max() + 1e-12
        if k.startswith("fcs.") and k.endswith(".bias"):  # sum(dZ) == 0 up to rounding noise under batch norm
            continue
        np.testing.assert_allclose(dense[k].numpy(), w, rtol=0, atol=2e-4 * scale + 1e-7, err_msg=k)
    for k, (idx, vals) in want_sparse.items():
        rows, gsum = O.coalesce(idx, vals)
        acc = torch.zeros_like(p[k])
        acc.index_add_(0, sparse[k][0], sparse[k][1])
        np.testing.assert_allclose(acc.numpy()[rows], gsum, rtol=0, atol=2e-4 * np.abs(vals).max() + 1e-8, err_msg=k)


@pytest.mark.parametrize("name", G.names("train_mlp_"))
def test_emulation_tracks_fp32_oracle_at_bf16_level(name):
    """The size of the bf16 effect on this loss: per-pass gradients cancel (hinge: d s_neg - d s_pos), so a
    2^-9 rounding shows up as up to ~35 % of the largest entry of the difference.  The GPU tests use the same bound."""
    g = G.load(name)
    p = {k: torch.from_numpy(np.asarray(v)).clone() for k, v in G.section(g, "init").items()}
    batch = {k: torch.from_numpy(v) for k, v in G.batch_at(g, 0).items()}
    loss, dense, _ = E.mlp_grads_bf16(p, batch)
    want_loss, _, want_dense = O.mlp_grads(G.section(g, "init"), G.batch_at(g, 0))
    assert abs(loss - float(want_loss)) < 2e-2
    for k, w in want_dense.items():
        if k.startswith("fcs.") and k.endswith(".bias") or not np.abs(w).max():
            continue
        got = dense[k].numpy().ravel()
        cos = float(got @ w.ravel() / (np.linalg.norm(got) * np.linalg.norm(w.ravel()) + 1e-30))
        err = np.abs(got - w.ravel()).max() / np.abs(w).max()
        assert cos > 0.97 and err < 0.5, f"{k}: cosine {cos:.4f}, max error {err:.3f} of scale"
