"""Batched predict / top-k (trs_predict_topk: tcgen05 score GEMM with the top-k fused into the epilogue +
exact fp32 re-scoring) on the B200 (``-m gpu``).

Bar: the returned item ids and scores are BIT-EXACT those of the fp32 path -- ``trs_scores`` over all items
followed by ``torch.sort(stable=True, descending=True)`` (ties -> lower item id), the stated equivalent of
the reference's ``predict`` (model.py:341-452, whose own sort is unstable) -- and equal the
numpy oracle / the golden vectors recorded from the live reference on exact-arithmetic fixtures."""
import numpy as np
import pytest
import torch

from oracle import cf_oracle as O
from tests import _golden as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _make(net_type, U, I, D, C, dev, scale=None, seed=0, integer=False):
    from torchrecsys_b200.collaborative.fm import FM
    from torchrecsys_b200.collaborative.linear import Linear
    torch.manual_seed(seed)
    cls = Linear if net_type == "linear" else FM
    net = cls(U, I, {"c": C} if C else {}, D, use_metadata=bool(C), use_cuda=True)
    with torch.no_grad():
        for p in net.parameters():
            if integer:
                p.copy_(torch.randint(-2, 3, p.shape).float() / 4)
            elif scale:
                p.copy_(torch.randn_like(p) * scale)
    return net.to(dev).eval()


def _exact(net, users, I, item_meta, k, dev):
    """The fp32 path: every item scored by trs_scores, stable descending sort."""
    from torchrecsys_b200 import _lib
    items = torch.arange(I, device=dev)
    idx, sc = [], []
    for u in users.tolist():
        s = _lib.scores(net.abi_model(), torch.full_like(items, u), items, item_meta)
        v, o = torch.sort(s, descending=True, stable=True)
        idx.append(o[:k])
        sc.append(v[:k])
    return torch.stack(idx), torch.stack(sc)


@pytest.mark.parametrize("net_type,D,C", [("linear", 128, 0), ("linear", 64, 0), ("fm", 64, 0), ("fm", 64, 100),
                                          ("linear", 80, 7), ("linear", 16, 0), ("fm", 200, 0)])
def test_topk_is_bit_exact_with_the_fp32_path(dev, net_type, D, C):
    from torchrecsys_b200 import _lib
    U, I, k = 1000, 30011, 100
    net = _make(net_type, U, I, D, C, dev, scale=0.3 if net_type == "linear" else 0.08, seed=D + C)
    rng = np.random.default_rng(1)
    users = torch.from_numpy(rng.choice(U, 171, replace=False)).to(dev)
    item_meta = (torch.arange(I, device=dev) % C).view(-1, 1).contiguous() if C else None
    idx, score, over = _lib.predict_topk(net.abi_model(), users, k, item_meta)
    torch.cuda.synchronize()
    assert int(over.sum()) == 0
    want_idx, want_score = _exact(net, users, I, item_meta, k, dev)
    assert torch.equal(idx, want_idx)
    assert torch.equal(score, want_score)


def test_topk_tie_break_is_lower_item_id_first(dev):
    """Quarter-integer weights: scores are exact in any summation order and full of ties."""
    from torchrecsys_b200 import _lib
    U, I, D, k = 64, 5000, 16, 50
    net = _make("linear", U, I, D, 0, dev, integer=True, seed=5)
    users = torch.arange(U, device=dev)
    idx, score, over = _lib.predict_topk(net.abi_model(), users, k)
    assert int(over.sum()) == 0
    want_idx, want_score = _exact(net, users, I, None, k, dev)
    assert torch.equal(idx, want_idx) and torch.equal(score, want_score)
    # and against the numpy oracle's stable top-k
    params = {n: p.detach().cpu().numpy() for n, p in net.state_dict().items()}
    for u in (0, 17, 63):
        np.testing.assert_array_equal(idx[u].cpu().numpy(), O.predict_topk("linear", params, u, k))


@pytest.mark.parametrize("net_type", ["linear", "fm"])
def test_topk_matches_golden_reference_predict(dev, net_type):
    from torchrecsys_b200 import _lib
    from torchrecsys_b200.collaborative.fm import FM
    from torchrecsys_b200.collaborative.linear import Linear
    g = G.load(f"predict_{net_type}")
    nu, ni, D, k = (int(x) for x in g["meta"])
    net = (Linear if net_type == "linear" else FM)(nu, ni, {}, D, use_metadata=False, use_cuda=True)
    net.load_state_dict({n: torch.from_numpy(v) for n, v in G.section(g, "init").items()})
    net = net.to(dev).eval()
    idx, score, over = _lib.predict_topk(net.abi_model(), torch.arange(nu, device=dev), k)
    assert int(over.sum()) == 0
    np.testing.assert_array_equal(idx.cpu().numpy(), g["stable_topk"])
    want = np.take_along_axis(g["scores"], g["stable_topk"], axis=1)
    if net_type == "linear":
        np.testing.assert_array_equal(score.cpu().numpy(), want)
    else:
        np.testing.assert_allclose(score.cpu().numpy(), want, rtol=1e-6, atol=1e-7)


def test_topk_small_catalogue_and_offsets(dev):
    """Fewer items than k (tail filled with -1), one user, item_offset for a sharded catalogue."""
    from torchrecsys_b200 import _lib
    net = _make("linear", 10, 37, 32, 0, dev, scale=0.5, seed=9)
    users = torch.tensor([3], device=dev)
    idx, score, over = _lib.predict_topk(net.abi_model(), users, 50, item_offset=1000)
    want_idx, want_score = _exact(net, users, 37, None, 37, dev)
    assert torch.equal(idx[:, :37], want_idx + 1000) and torch.equal(score[:, :37], want_score)
    assert bool((idx[:, 37:] == -1).all())


def test_saturated_fm_scores_overflow_and_fall_back_to_the_exact_path(dev):
    """FM logits far beyond the sigmoid's fp32 saturation: thousands of items tie at exactly 1.0, the
    candidate superset cannot hold them, the kernel says so, and predict() falls back to the fp32 path."""
    import pandas as pd
    from torchrecsys_b200 import _lib
    from torchrecsys_b200.model import TorchRecSys
    U, I, D = 40, 6000, 16
    rng = np.random.default_rng(0)
    df = pd.DataFrame({"user": np.concatenate([np.arange(U), rng.integers(0, U, 20000)]),
                       "item": np.concatenate([np.arange(U) % I, rng.integers(0, I, 20000)])})
    df = pd.concat([df, pd.DataFrame({"user": 0, "item": np.arange(I)})], ignore_index=True)
    model = TorchRecSys(df, "user", "item", n_factors=D, net_type="fm", use_cuda=True)
    with torch.no_grad():
        model.net.user.weight.fill_(1.0)
        model.net.item.weight.fill_(2.0)
    idx, score, over = _lib.predict_topk(model.net.abi_model(), torch.arange(4, device=dev), 10)
    assert int(over.sum()) == 4
    top = model.predict(2, top_k=10)
    assert top.tolist() == list(range(10))  # every score is 1.0: lowest ids first


def test_topk_long_item_stream_many_user_tiles(dev):
    """A catalogue long enough (> 768 tiles of 192 items) for the producers' lock-step window to engage, more user
    tiles than one wave of CTAs shares an item range with, a ragged last tile; checked for a sample of users."""
    from torchrecsys_b200 import _lib
    U, I, D, k = 40000, 400_003, 32, 20
    net = _make("linear", U, I, D, 0, dev, scale=0.3, seed=11)
    users = torch.arange(0, U, 2, device=dev)  # 20000 users = 157 user tiles: more CTAs than SMs
    idx, score, over = _lib.predict_topk(net.abi_model(), users, k)
    torch.cuda.synchronize()
    assert int(over.sum()) == 0
    pick = torch.tensor([0, 1, 127, 128, 9999, 19871, 19999], device=dev)
    want_idx, want_score = _exact(net, users[pick], I, None, k, dev)
    assert torch.equal(idx[pick], want_idx)
    assert torch.equal(score[pick], want_score)
    # every row is a descending list of distinct valid items
    assert bool((score[:, :-1] >= score[:, 1:]).all())
    assert int(idx.min()) >= 0 and int(idx.max()) < I
    assert bool((torch.sort(idx, dim=1).values[:, 1:] != torch.sort(idx, dim=1).values[:, :-1]).all())


def test_item_operand_reuse_gives_identical_results_and_is_dropped_when_the_model_changes(dev):
    """trs_predict_topk_reuse: a second call on the same workspace skips the re-cast of the item tables; a changed key
    (the model was trained / loaded) rebuilds it."""
    from torchrecsys_b200 import _lib
    from torchrecsys_b200.collaborative.linear import Linear
    torch.manual_seed(5)
    net = Linear(400, 3000, {}, 64, use_metadata=False, use_cuda=True).to(dev).eval()
    with torch.no_grad():
        net.item_bias.weight.normal_(0, 0.05)
    users = torch.arange(300, device=dev)
    cache = _lib.TopkCache()
    a = _lib.predict_topk(net.abi_model(), users, 20, cache=cache, cache_key=1)
    ws = cache.ws
    b = _lib.predict_topk(net.abi_model(), users, 20, cache=cache, cache_key=1)      # reuses the operand
    assert cache.ws is ws and torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    with torch.no_grad():
        net.item.weight.mul_(-1.0)                                                   # the model changes ...
    c = _lib.predict_topk(net.abi_model(), users, 20, cache=cache, cache_key=2)      # ... and so does the key
    fresh = _lib.predict_topk(net.abi_model(), users, 20)
    assert cache.ws is not ws and torch.equal(c[0], fresh[0]) and not torch.equal(c[0], a[0])
