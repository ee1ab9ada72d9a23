"""The two plan builders -- one CTA per (step, id space) in shared memory (batches up to 8192) and the tiled
multi-launch radix sort (any batch) -- must agree: identical sorted (row, lookup) pairs, identical singleton
flags and work-item sets, and bit-identical parameters after training on either plan (``-m gpu``).  The plan is
integer work: it is also compared, entry by entry, with a numpy restatement (stable argsort per step and id space --
the order coalesce() sums duplicates in, torch: optim/_functional.py:44 --, run lengths, flag bytes)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _plan_arrays(_lib, model, epoch, dev, tiled):
    os.environ["TRS_PLAN_TILED"] = "1" if tiled else "0"
    try:
        plan = _lib.plan_build(model, epoch, dev)
        torch.cuda.synchronize()
    finally:
        os.environ.pop("TRS_PLAN_TILED", None)
    return plan.cpu().numpy().copy()


@pytest.mark.parametrize("B,n,F,U,I", [(1024, 3000, 1, 5000, 300), (8192, 20000, 2, 1 << 20, 70000), (100, 250, 0, 40, 7)])
def test_fused_plan_equals_tiled_plan(B, n, F, U, I):
    from torchrecsys_b200 import _lib
    from torchrecsys_b200.collaborative.fm import FM
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(B + F)
    C = 11
    net = FM(U, I, {f"m{f}": C for f in range(F)}, 8, use_metadata=F > 0, use_cuda=True).to(dev)
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    ids = {"user": rng.integers(0, U, n), "pos": rng.zipf(1.3, n) % I, "neg": rng.integers(0, I, n)}
    if F:
        meta = rng.integers(0, C, (I, F))
        ids["pos_meta"], ids["neg_meta"] = meta[ids["pos"]], meta[ids["neg"]]
    t = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.int64)).to(dev) for k, v in ids.items()}
    model = net.abi_model(opt.state, (None, None))
    epoch = _lib.make_epoch(t["user"], t["pos"], t["neg"], t.get("pos_meta"), t.get("neg_meta"), B)
    a = _plan_arrays(_lib, model, epoch, dev, tiled=False)
    b = _plan_arrays(_lib, model, epoch, dev, tiled=True)
    steps = -(-n // B)
    lookups = B * (3 + 2 * F)
    # layout of the plan buffer (plan.cuh): 256-byte aligned arrays in this order
    off = 0

    def take(nbytes):
        nonlocal off
        o = off
        off += (nbytes + 255) // 256 * 256
        return o

    spans = {"user_key": take(4 * n), "user_perm": take(4 * n), "item_key": take(8 * n), "item_perm": take(8 * n)}
    for f in range(F):
        spans[f"meta_key{f}"], spans[f"meta_perm{f}"] = take(8 * n), take(8 * n)
    item_cnt, long_cnt = take(4 * steps), take(4 * steps)
    single_user, single_item = take((n + 3) // 4 * 4), take((2 * n + 3) // 4 * 4)
    item_cap, long_cap = lookups, lookups // 9 + 1
    items, long_segs = take(16 * steps * item_cap), take(16 * steps * long_cap)
    assert off == a.size == b.size
    for name, o in spans.items():
        nb = 4 * n * (1 if name.startswith("user") else 2)
        assert np.array_equal(a[o:o + nb], b[o:o + nb]), name
    # ... and against a numpy restatement: stable argsort per step and id space, segment lengths, flags
    lookups_of = {"user": lambda lo, hi: ids["user"][lo:hi],
                  "item": lambda lo, hi: np.concatenate([ids["pos"][lo:hi], ids["neg"][lo:hi]])}
    for f in range(F):
        lookups_of[f"meta{f}"] = (lambda lo, hi, f=f: np.concatenate([ids["pos_meta"][lo:hi, f], ids["neg_meta"][lo:hi, f]]))
    want_single = {"user": np.zeros(n, np.uint8), "item": np.zeros(2 * n, np.uint8)}
    want_items, want_long = [set() for _ in range(steps)], [set() for _ in range(steps)]
    for sp_i, (name, mult) in enumerate([("user", 1), ("item", 2)] + [(f"meta{f}", 2) for f in range(F)]):
        key_name = {"user": "user_key", "item": "item_key"}.get(name, name.replace("meta", "meta_key"))
        perm_name = key_name.replace("key", "perm")
        got_k = a[spans[key_name]:spans[key_name] + 4 * mult * n].view(np.uint32)
        got_p = a[spans[perm_name]:spans[perm_name] + 4 * mult * n].view(np.uint32)
        prev_rows = None
        for st in range(steps):
            lo, hi = st * B, min((st + 1) * B, n)
            look = lookups_of[name](lo, hi).astype(np.int64)
            order = np.argsort(look, kind="stable")
            base = mult * st * B
            assert np.array_equal(got_k[base:base + len(look)], look[order].astype(np.uint32)), (name, st)
            assert np.array_equal(got_p[base:base + len(look)], order.astype(np.uint32)), (name, st)
            rows, first, cnt = np.unique(look[order], return_index=True, return_counts=True)
            for r, k0, c in zip(rows.tolist(), first.tolist(), cnt.tolist()):
                if c > 8:
                    want_long[st].add((sp_i, k0, c, r))
                elif c > 1 or name not in want_single:
                    want_items[st].add((sp_i | (c << 8), k0, r, int(order[k0])))
            if name in want_single:
                flags = np.zeros(len(look), np.uint8)
                flags[order[first[cnt == 1]]] |= 1
                if prev_rows is not None:
                    flags[np.isin(look, prev_rows)] |= 2
                want_single[name][base:base + len(look)] = flags
            prev_rows = rows
    assert np.array_equal(a[single_user:single_user + n], want_single["user"])
    assert np.array_equal(a[single_item:single_item + 2 * n], want_single["item"])
    assert np.array_equal(a[item_cnt:single_user], b[item_cnt:single_user]), "segment counts"
    assert np.array_equal(a[single_user:single_user + n], b[single_user:single_user + n])
    assert np.array_equal(a[single_item:single_item + 2 * n], b[single_item:single_item + 2 * n])
    ca, la = a[item_cnt:item_cnt + 4 * steps].view(np.uint32), a[long_cnt:long_cnt + 4 * steps].view(np.uint32)
    assert la.max() > 0 or F == 0
    for s in range(steps):  # the lists are sets: slots are handed out by atomics
        for o, cap, cnt in ((items, item_cap, ca[s]), (long_segs, long_cap, la[s])):
            xa = a[o + 16 * s * cap:o + 16 * (s * cap + cnt)].view(np.uint32).reshape(-1, 4)
            xb = b[o + 16 * s * cap:o + 16 * (s * cap + cnt)].view(np.uint32).reshape(-1, 4)
            ka = np.lexsort(xa.T[::-1])
            kb = np.lexsort(xb.T[::-1])
            assert np.array_equal(xa[ka], xb[kb]), f"step {s}"
            want = want_items[s] if o == items else want_long[s]
            assert set(map(tuple, xa.tolist())) == want, f"step {s}: work items differ from the numpy restatement"


@pytest.mark.parametrize("opt_name", ["adagrad", "sparse_adam"])
def test_training_is_bit_identical_on_either_plan(opt_name):
    from torchrecsys_b200.collaborative.fm import FM
    from torchrecsys_b200.engine import EpochRunner
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(5)
    U, I, C, B, n, D = 4000, 900, 13, 2048, 2048 * 4 - 100, 32
    meta = rng.integers(0, C, (I, 1))
    ids = {"user": rng.integers(0, U, n), "pos": rng.zipf(1.2, n) % I, "neg": rng.integers(0, I, n)}
    ids["pos_meta"], ids["neg_meta"] = meta[ids["pos"]], meta[ids["neg"]]
    t = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.int64)).to(dev) for k, v in ids.items()}
    out = []
    for tiled in ("0", "1"):
        torch.manual_seed(3)
        net = FM(U, I, {"m0": C}, D, use_metadata=True, use_cuda=True).to(dev)
        opt = (torch.optim.Adagrad(net.parameters(), lr=0.05) if opt_name == "adagrad"
               else torch.optim.SparseAdam(list(net.parameters()), lr=0.01))
        os.environ["TRS_PLAN_TILED"] = tiled
        try:
            loss = EpochRunner(net, opt).run(t, B)
            torch.cuda.synchronize()
        finally:
            os.environ.pop("TRS_PLAN_TILED", None)
        out.append((loss.cpu().numpy(), {k: v.detach().cpu().numpy() for k, v in net.state_dict().items()}))
    assert np.array_equal(out[0][0], out[1][0])
    for k in out[0][1]:
        assert np.array_equal(out[0][1][k], out[1][1][k]), k


def test_long_epoch_is_cut_into_calls_of_whole_steps(monkeypatch):
    """More steps than one plan launch indexes (gridDim.y): EpochRunner.run runs them as consecutive calls;
    the result is bit-identical to a single call."""
    from torchrecsys_b200 import engine
    from torchrecsys_b200.collaborative.linear import Linear
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(9)
    U, I, B, D = 500, 300, 2, 8
    n = B * 1000 + 1
    ids = {"user": rng.integers(0, U, n), "pos": rng.integers(0, I, n), "neg": rng.integers(0, I, n)}
    t = {k: torch.from_numpy(v).to(dev) for k, v in ids.items()}
    out = []
    for cap in (1 << 20, 64):  # one call / 16 calls of 64 steps
        monkeypatch.setattr(engine, "MAX_STEPS_PER_CALL", cap)
        torch.manual_seed(4)
        net = Linear(U, I, {}, D, use_metadata=False, use_cuda=True).to(dev)
        opt = torch.optim.Adagrad(net.parameters(), lr=0.05)
        loss = engine.EpochRunner(net, opt).run(t, B)
        torch.cuda.synchronize()
        assert loss.shape[0] == 1001
        out.append((loss.cpu().numpy(), {k: v.detach().cpu().numpy() for k, v in net.state_dict().items()},
                    int(opt.state[next(iter(net.parameters()))]["step"])))
    assert out[0][2] == out[1][2] == 1001
    assert np.array_equal(out[0][0], out[1][0])
    for k in out[0][1]:
        assert np.array_equal(out[0][1][k], out[1][1][k]), k
