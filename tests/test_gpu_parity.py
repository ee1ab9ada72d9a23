"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the golden vectors
recorded from the live reference.  Everything here needs a B200 (``-m gpu``).

Tolerances: gathers, Philox negatives, sort plans and top-k indices are bit-exact;
fp32 scores / losses rtol 1e-5 atol 1e-6; rows after one update atol 2e-6; after 20 updates the
bound loosens for Adam, whose m/(sqrt(v)+eps) amplifies rounding where the gradient is ~0."""
import numpy as np
import pytest
import torch

from oracle import cf_oracle as O
from tests import _golden as G

pytestmark = pytest.mark.gpu

TRAIN = [n for n in G.names("train_") if "_mlp_" not in n]


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _lib():
    from torchrecsys_b200 import _lib as L
    return L


def _net_from_golden(g, net_type, F, dev):
    from torchrecsys_b200.collaborative.fm import FM
    from torchrecsys_b200.collaborative.linear import Linear
    U, I, C, D = (int(x) for x in g["meta"][:4])
    cls = Linear if net_type == "linear" else FM
    net = cls(U, I, {f"m{f}": C for f in range(F)}, D, use_metadata=F > 0, use_cuda=True)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in G.section(g, "init").items()})
    return net.to(dev)


def _torch_opt(kind, net, lr):
    if kind == "sparse_adam":
        return torch.optim.SparseAdam(list(net.parameters()), lr=lr)
    if kind == "adagrad":
        return torch.optim.Adagrad(net.parameters(), lr=lr)
    return torch.optim.SGD(net.parameters(), lr=lr)


def _samples(g, steps, dev):
    out = {}
    for k in ("user", "pos", "neg", "pos_meta", "neg_meta"):
        if f"batch/{k}" in g:
            v = g[f"batch/{k}"][:steps]
            out[k] = torch.from_numpy(v.reshape((-1,) + v.shape[2:])).to(dev).contiguous()
    return out


# ------------------------------------------------------------------------------------------
def test_device_info(dev):
    sm, grid, block = _lib().device_info()
    assert sm >= 100 and grid % sm == 0 and block % 32 == 0


@pytest.mark.parametrize("dim", [4, 16, 64, 80, 128, 256, 6, 33])
@pytest.mark.parametrize("n_meta", [0, 2])
def test_gather_sum_bit_exact(dev, dim, n_meta):
    rng = np.random.default_rng(dim * 10 + n_meta)
    table = rng.standard_normal((1000, dim)).astype(np.float32)
    metas = [rng.standard_normal((13, dim)).astype(np.float32) for _ in range(n_meta)]
    for n in (0, 1, 7, 1025):
        idx = rng.integers(0, 1000, n)
        midx = rng.integers(0, 13, (n, n_meta)) if n_meta else None
        got = _lib().embed_gather_sum(torch.from_numpy(table).to(dev), torch.from_numpy(idx).to(dev),
                                      [torch.from_numpy(m).to(dev) for m in metas],
                                      None if midx is None else torch.from_numpy(midx).to(dev))
        want = O.gather_sum(table, idx, metas, midx)
        assert np.array_equal(got.cpu().numpy(), want.reshape(n, dim))


@pytest.mark.parametrize("name", TRAIN)
def test_scores_match_reference(dev, name):
    g = G.load(name)
    net_type, F, _ = G.parse_train_name(name)
    net = _net_from_golden(g, net_type, F, dev)
    b = {k: torch.from_numpy(v).to(dev) for k, v in G.batch_at(g, 0).items()}
    batch = {"user_id": b["user"], "pos_item_id": b["pos"], "neg_item_id": b["neg"]}
    if F:
        batch["pos_metadata_id"], batch["neg_metadata_id"] = b["pos_meta"], b["neg_meta"]
    pos = net.forward(batch, "user_id", "pos_item_id", "pos_metadata_id")
    neg = net.forward(batch, "user_id", "neg_item_id", "neg_metadata_id")
    assert pos.shape == g["pos0"].shape  # (B,1) linear, (B,) fm
    assert pos.requires_grad  # forward is autograd-visible, as in the reference (tests/test_gpu_autograd.py)
    np.testing.assert_allclose(pos.detach().cpu().numpy(), g["pos0"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(neg.detach().cpu().numpy(), g["neg0"], rtol=1e-5, atol=1e-6)


def test_scores_reference_batch_layouts(dev):
    """(B, L) bags for one feature and (B, F, L) for several give the same scores as [B, F]."""
    g = G.load("train_fm_F2_sparse_adam")
    net = _net_from_golden(g, "fm", 2, dev)
    b = {k: torch.from_numpy(v).to(dev) for k, v in G.batch_at(g, 0).items()}
    base = {"user_id": b["user"], "pos_item_id": b["pos"]}
    a = net.forward(dict(base, pos_metadata_id=b["pos_meta"]), "user_id", "pos_item_id", "pos_metadata_id")
    padded = torch.stack([b["pos_meta"], torch.zeros_like(b["pos_meta"])], dim=2)  # (B, F, L=2)
    c = net.forward(dict(base, pos_metadata_id=padded), "user_id", "pos_item_id", "pos_metadata_id")
    assert torch.equal(a, c)
    np.testing.assert_allclose(a.detach().cpu().numpy(), g["pos0"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("n_items,n,first", [(3, 5000, 0), (200_000, 100_000, 10_000_000_000), (2, 999, 7)])
def test_philox_negatives_bit_exact(dev, n_items, n, first):
    rng = np.random.default_rng(n)
    pos = rng.integers(0, n_items, n)
    item_meta = rng.integers(0, 100, (n_items, 2))
    neg, neg_meta = _lib().philox_negatives(1234, first, torch.from_numpy(pos).to(dev), n_items,
                                            torch.from_numpy(item_meta).to(dev))
    want = O.philox_negatives(1234, first, pos, n_items)
    assert np.array_equal(neg.cpu().numpy(), want)
    assert (want != pos).all()
    assert np.array_equal(neg_meta.cpu().numpy(), item_meta[want])


def _plan_arrays(plan, n, F):
    """Decode the plan buffer with the layout of plan.cuh (uint32 arrays, 256-byte aligned)."""
    raw = plan.cpu().numpy()
    off = 0
    out = {}

    def take(name, count):
        nonlocal off
        out[name] = raw[off:off + 4 * count].view(np.uint32)
        off += (4 * count + 255) // 256 * 256

    take("user_key", n), take("user_perm", n), take("item_key", 2 * n), take("item_perm", 2 * n)
    for f in range(F):
        take(f"meta_key{f}", 2 * n), take(f"meta_perm{f}", 2 * n)
    return out


@pytest.mark.parametrize("n,batch,n_users,n_items,F", [
    (1000, 64, 50, 30, 1),          # many duplicates, ragged last step (1000 = 15*64 + 40)
    (20000, 8192, 1_000_000, 200_000, 1),   # C2 shape: 3 radix passes, multi-tile segments
    (5000, 4096, 70_000, 300, 2),   # 17-bit users, 9-bit items
    (300, 512, 5, 2, 0),            # batch larger than the data, tiny id spaces
])
def test_plan_is_stable_sort_per_step(dev, n, batch, n_users, n_items, F):
    L = _lib()
    rng = np.random.default_rng(n)
    user, pos, neg = rng.integers(0, n_users, n), rng.integers(0, n_items, n), rng.integers(0, n_items, n)
    pm = rng.integers(0, 100, (n, F)) if F else None
    nm = rng.integers(0, 100, (n, F)) if F else None
    t = lambda a: None if a is None else torch.from_numpy(a).to(dev)
    dummy = torch.zeros(4, dtype=torch.float32, device=dev)

    def table(rows):
        tb = L.make_table(dummy)
        tb.n_rows = rows
        return tb

    model = L.make_model(L.NET_FM, 4, table(n_users), table(n_items), [table(100)] * F)
    tu, tp, tn, tpm, tnm = t(user), t(pos), t(neg), t(pm), t(nm)
    epoch = L.make_epoch(tu, tp, tn, tpm, tnm, batch)
    arrays = _plan_arrays(L.plan_build(model, epoch, dev), n, F)
    for s in range(-(-n // batch)):
        lo, hi = s * batch, min((s + 1) * batch, n)
        spaces = {"user": user[lo:hi], "item": np.r_[pos[lo:hi], neg[lo:hi]]}
        for f in range(F):
            spaces[f"meta{f}"] = np.r_[pm[lo:hi, f], nm[lo:hi, f]]
        for name, ids in spaces.items():
            mult = 1 if name == "user" else 2
            key_name = name + "_key" if not name.startswith("meta") else f"meta_key{name[4:]}"
            perm_name = key_name.replace("key", "perm")
            seg = slice(mult * lo, mult * lo + len(ids))
            order = np.argsort(ids, kind="stable")
            assert np.array_equal(arrays[perm_name][seg], order.astype(np.uint32)), (name, s)
            assert np.array_equal(arrays[key_name][seg], ids[order].astype(np.uint32)), (name, s)


@pytest.mark.parametrize("name", TRAIN)
def test_train_steps_match_reference(dev, name):
    """Fused fwd+bwd+update against the reference's forward/backward/optimizer.step(), driven with a
    real torch optimizer bound to the module's parameters (state lands in optimizer.state)."""
    from torchrecsys_b200.engine import EpochRunner
    g = G.load(name)
    net_type, F, opt = G.parse_train_name(name)
    B, steps = int(g["meta"][4]), int(g["meta"][5])
    for n_steps, section, tol in ((1, "after1", dict(rtol=1e-5, atol=2e-6)),
                                  (steps, "final", dict(rtol=2e-3, atol=2e-4) if opt == "sparse_adam"
                                   else dict(rtol=1e-4, atol=1e-5))):
        net = _net_from_golden(g, net_type, F, dev)
        optim = _torch_opt(opt, net, float(g["lr"]))
        runner = EpochRunner(net, optim)
        loss = runner.run(_samples(g, n_steps, dev), B)
        torch.cuda.synchronize()
        got_loss = loss.cpu().numpy()
        np.testing.assert_allclose(got_loss[0], g["loss"][0], rtol=1e-5, atol=1e-6)
        # later losses inherit the ill-conditioned rows explained below (a saturated sigmoid under
        # Adagrad/Adam): 2 of 800 user rows can take a different first step
        np.testing.assert_allclose(got_loss, g["loss"][:n_steps], rtol=5e-4, atol=1e-5)
        sd = {k: v.cpu().numpy() for k, v in net.state_dict().items()}
        init = G.section(g, "init")
        for k, v in G.section(g, section).items():
            ok = np.ones(v.shape[0], dtype=bool)
            if n_steps == 1:
                ok = G.well_conditioned_rows(net_type, opt, init, G.batch_at(g, 0), k, v.shape[0])
            np.testing.assert_allclose(sd[k][ok], v[ok], err_msg=f"{k} after {n_steps} steps", **tol)
        if n_steps == steps:
            named = dict(net.named_parameters())
            for k, v in G.section(g, "state").items():
                pname, sk = k.rsplit("/", 1)
                got = optim.state[named[pname]][sk]
                if sk == "step":
                    assert float(got) == steps
                elif not (net_type == "linear" and pname == "user_bias.weight"):
                    np.testing.assert_allclose(got.cpu().numpy(), v, rtol=1e-3, atol=1e-6, err_msg=k)
            # the optimizer object stays usable: state_dict round-trips
            optim.load_state_dict(optim.state_dict())


@pytest.mark.parametrize("net_type", ["linear", "fm"])
@pytest.mark.parametrize("opt", ["adagrad", "sparse_adam"])
def test_pos_equal_neg_gives_exactly_zero_update(dev, net_type, opt):
    """A sample whose negative equals its positive has an exactly-zero gradient in the reference
    (its two lookups cancel bit for bit), so Adagrad/Adam must not move anything -- not even by the
    lr-sized step that g/(|g|+eps) makes out of a 1e-10 rounding residue.  (Rows shared by several
    samples are left out: there the sum order decides, in the reference too.)"""
    from torchrecsys_b200.engine import EpochRunner
    g = G.load(f"train_{net_type}_F0_{opt}")
    U, I = int(g["meta"][0]), int(g["meta"][1])
    net = _net_from_golden(g, net_type, 0, dev)
    before = {k: v.detach().clone() for k, v in net.state_dict().items()}
    B = 25
    rng = np.random.default_rng(5)
    user = torch.from_numpy(rng.permutation(U)[:B]).to(dev)
    pos = torch.from_numpy(rng.permutation(I)[:B]).to(dev)
    optim = _torch_opt(opt, net, 0.05)
    loss = EpochRunner(net, optim).run({"user": user, "pos": pos, "neg": pos.clone()}, B)
    assert float(loss[0]) == 1.0  # hinge of (s - s + 1)
    after = net.state_dict()
    for k in before:
        assert torch.equal(after[k], before[k]), k


def test_train_split_launches_equal_one_launch(dev):
    """Steps [0,7) + [7,20) in two launches == [0,20) in one, bit for bit (deterministic reduce)."""
    L = _lib()
    from torchrecsys_b200 import engine
    g = G.load("train_fm_F1_sparse_adam")
    B, steps = int(g["meta"][4]), int(g["meta"][5])
    results = []
    for cuts in ([0, steps], [0, 7, steps], [0, steps]):
        net = _net_from_golden(g, "fm", 1, dev)
        optim = _torch_opt("sparse_adam", net, float(g["lr"]))
        b = engine.bind_optimizer(optim, list(net.parameters()))
        model = net.abi_model(optim.state, b.keys)
        smp = _samples(g, steps, dev)
        epoch = L.make_epoch(smp["user"], smp["pos"], smp["neg"], smp["pos_meta"], smp["neg_meta"], B)
        scales = torch.tensor(engine.step_scales(b, steps), dtype=torch.float64).float().to(dev)
        optim_c = L.Optim(b.kind, 0, b.beta1, b.beta2, b.eps, scales.data_ptr())
        plan, ws = L.plan_build(model, epoch, dev), L.train_workspace(model, epoch, dev)
        loss = torch.zeros(steps, device=dev)
        for a, z in zip(cuts[:-1], cuts[1:]):
            L.train_steps(model, epoch, optim_c, plan, ws, a, z - a, loss[a:z])
        torch.cuda.synchronize()
        results.append([loss.clone()] + [p.detach().clone() for p in net.parameters()])
    for other in results[1:]:
        for x, y in zip(results[0], other):
            assert torch.equal(x, y)


def test_eval_pairwise_matches_reference(dev):
    L = _lib()
    ge = G.load("eval_pairwise")
    # scores equal to the fixture: dim-4 linear model with zero embeddings, scores carried by the item bias
    vals = np.unique(np.r_[ge["pos"], ge["neg"]])
    n = len(ge["pos"])
    lut = {v: i for i, v in enumerate(vals)}
    zeros = torch.zeros((len(vals), 4), device=dev)
    bias = torch.from_numpy(vals.astype(np.float32)).to(dev).view(-1, 1).contiguous()
    model = L.make_model(L.NET_LINEAR, 4, L.make_table(zeros), L.make_table(zeros, lin=bias), [])
    user = torch.zeros(n, dtype=torch.int64, device=dev)
    pos = torch.tensor([lut[v] for v in ge["pos"]], device=dev)
    neg = torch.tensor([lut[v] for v in ge["neg"]], device=dev)
    loss, auc, sp, sn = L.eval_pairwise(model, L.make_epoch(user, pos, neg, batch=n), want_scores=True)
    assert np.array_equal(sp.cpu().numpy(), ge["pos"]) and np.array_equal(sn.cpu().numpy(), ge["neg"])
    assert float(auc[0]) == float(ge["auc"])
    np.testing.assert_allclose(float(loss[0]), float(ge["hinge"]), rtol=1e-6)
    # ragged batches: unweighted per-batch values as evaluate() averages them (model.py:329-333)
    loss, auc, _, _ = L.eval_pairwise(model, L.make_epoch(user, pos, neg, batch=100))
    for bt in range(3):
        sl = slice(bt * 100, min((bt + 1) * 100, n))
        assert float(auc[bt]) == float(O.pairwise_auc(ge["pos"][sl], ge["neg"][sl]))
        np.testing.assert_allclose(float(loss[bt]), O.hinge_loss(ge["pos"][sl], ge["neg"][sl]), rtol=1e-6)


@pytest.mark.parametrize("net_type", ["linear", "fm"])
def test_predict_topk_matches_reference(dev, net_type):
    import pandas as pd
    from torchrecsys.model import TorchRecSys
    g = G.load(f"predict_{net_type}")
    nu, ni, D, k = (int(x) for x in g["meta"])
    df = pd.DataFrame({"user_id": np.resize(np.arange(nu), 100), "item_id": np.resize(np.arange(ni), 100)})
    model = TorchRecSys(df, "user_id", "item_id", n_factors=D, net_type=net_type, use_cuda=True)
    model.net.load_state_dict({k_: torch.from_numpy(v) for k_, v in G.section(g, "init").items()})
    for u in range(nu):
        top = model.predict(u, top_k=k, prediction_batch_size=7)
        assert top.dtype == torch.int64 and not top.is_cuda and top.shape == (k,)
        if net_type == "linear":  # exact-arithmetic fixture: bit-exact ranking, ties -> lower id
            assert np.array_equal(top.numpy(), g["stable_topk"][u])
        sref = g["scores"][u]
        np.testing.assert_allclose(np.sort(sref[top.numpy()])[::-1], np.sort(sref)[::-1][:k], rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------
# full-size property checks (BASELINE configs): no oracle run at this size, invariants instead
# ------------------------------------------------------------------------------------------
def _c2_model(dev, n_users=1_000_000, n_items=200_000, D=64, C=100, opt="sparse_adam", net="fm"):
    from torchrecsys_b200.collaborative.fm import FM
    from torchrecsys_b200.collaborative.linear import Linear
    torch.manual_seed(1234)
    cls = FM if net == "fm" else Linear
    model = cls(n_users, n_items, {"product_category": C}, D, use_metadata=True, use_cuda=True).to(dev)
    optim = _torch_opt(opt, model, {"sparse_adam": 1e-3, "adagrad": 1e-2, "sgd": 1e-2}[opt])
    return model, optim


def test_c2_full_size_invariants(dev):
    """FM + metadata, 1M x 200k, D=64, B=8192, Philox negatives: (i) untouched rows are bit-identical,
    (ii) touched rows = exactly the rows the batch names, (iii) two identical runs agree bitwise,
    (iv) a step on a batch equals the oracle's closed form on the touched rows."""
    L = _lib()
    from torchrecsys_b200.engine import EpochRunner
    B, steps = 8192, 3
    n = B * steps
    rng = np.random.default_rng(1234)
    user = torch.from_numpy(rng.integers(0, 1_000_000, n)).to(dev)
    pos = torch.from_numpy(rng.integers(0, 200_000, n)).to(dev)
    item_meta = (torch.arange(200_000, device=dev) % 100).view(-1, 1).contiguous()
    neg, neg_meta = L.philox_negatives(1234, 0, pos, 200_000, item_meta)
    smp = {"user": user, "pos": pos, "neg": neg, "pos_meta": item_meta[pos].contiguous(), "neg_meta": neg_meta}
    runs = []
    for _ in range(2):
        net, optim = _c2_model(dev)
        before = {k: v.detach().clone() for k, v in net.state_dict().items()}
        loss = EpochRunner(net, optim).run(smp, B)
        torch.cuda.synchronize()
        runs.append((loss.clone(), {k: v.detach().clone() for k, v in net.state_dict().items()}, before))
    (l0, a0, before), (l1, a1, _) = runs
    assert torch.equal(l0, l1) and all(torch.equal(a0[k], a1[k]) for k in a0)
    assert torch.isfinite(l0).all() and (l0 > 0).all()
    touched_u = torch.zeros(1_000_000, dtype=torch.bool, device=dev).index_fill_(0, user, True)
    touched_i = torch.zeros(200_000, dtype=torch.bool, device=dev).index_fill_(0, torch.cat([pos, neg]), True)
    for key, mask in (("user.weight", touched_u), ("item.weight", touched_i),
                      ("linear_user.weight", touched_u), ("linear_item.weight", touched_i)):
        changed = (a0[key] != before[key]).any(dim=1)
        assert not (changed & ~mask).any(), f"{key}: an untouched row changed"
        assert (changed | ~mask).float().mean() > 0.999, f"{key}: touched rows did not move"

    # (iv) one step against the numpy oracle restricted to the touched rows
    net, optim = _c2_model(dev, opt="adagrad")
    p0 = {k: v.detach().cpu().numpy().copy() for k, v in net.state_dict().items()}
    one = {k: v[:B].contiguous() for k, v in smp.items()}
    loss = EpochRunner(net, optim).run(one, B)
    batch = {k: v.cpu().numpy() for k, v in one.items()}
    spec = O.OptSpec("adagrad", lr=1e-2)
    state = O.init_opt_state(p0, spec)
    want_loss = O.train_step("fm", p0, state, batch, spec, 1)
    np.testing.assert_allclose(float(loss[0]), want_loss, rtol=1e-5)
    got = {k: v.detach().cpu().numpy() for k, v in net.state_dict().items()}
    for k in p0:  # Adagrad's first step is lr*g/(|g|+eps) with lr = 1e-2: a few ulp of g -> ~3e-6
        np.testing.assert_allclose(got[k], p0[k], rtol=1e-5, atol=1e-5, err_msg=k)


def test_fit_evaluate_predict_end_to_end_against_cpu_port(dev):
    """README quickstart shape (C1) end to end: same split and negatives as the CPU port of the reference, the port
    fed the device loader's own (replayable) shuffles -> same per-epoch losses and the same top-k."""
    import pandas as pd
    from oracle import torch_port as TP
    from torchrecsys.model import TorchRecSys
    rng = np.random.default_rng(1234)
    n_u, n_i, n_int, D, B = 300, 100, 6000, 16, 256
    df = pd.DataFrame({"user": np.r_[np.arange(n_u), rng.integers(0, n_u, n_int - n_u)],
                       "item": np.r_[np.arange(n_i), rng.integers(0, n_i, n_int - n_i)]})
    np.random.seed(1234)
    torch.manual_seed(1234)
    model = TorchRecSys(df, "user", "item", n_factors=D, net_type="linear", use_cuda=True)
    init = {k: v.detach().cpu().clone() for k, v in model.net.state_dict().items()}
    train = model.data_processor.train_data
    opt = torch.optim.SparseAdam(list(model.parameters()), lr=1e-2)

    port = TP.make_net("linear", n_u, n_i, [], D)
    port.load_state_dict(init)
    popt = TP.make_optimizer("sparse_adam", port, lr=1e-2)

    torch.manual_seed(99)
    import io, contextlib
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        model.fit(opt, epochs=3, batch_size=B)
        model.evaluate(batch_size=B)
    lines = [l for l in buf.getvalue().splitlines() if "Training Loss" in l]
    got_losses = [float(l.rsplit(":", 1)[1]) for l in lines]

    n = train["user_id"].numel()
    want_losses = []
    for e in range(3):
        perm = model._epoch_permutation(n, dev, epoch_index=e).cpu()   # the device loader's shuffle of epoch e
        assert torch.equal(torch.sort(perm)[0], torch.arange(n))
        tot, nb = 0.0, 0
        for lo in range(0, n, B):
            sel = perm[lo:lo + B]
            tot += TP.train_step(port, popt, {"user": train["user_id"][sel], "pos": train["pos_item_id"][sel],
                                              "neg": train["neg_item_id"][sel]})
            nb += 1
        want_losses.append(tot / nb)
    np.testing.assert_allclose(got_losses, want_losses, atol=2e-4)
    sd = model.net.state_dict()
    for k, v in port.state_dict().items():
        np.testing.assert_allclose(sd[k].cpu().numpy(), v.numpy(), rtol=5e-3, atol=5e-4, err_msg=k)
    assert set(model.last_eval) == {"loss", "auc"} and 0.0 <= model.last_eval["auc"] <= 1.0
    top = model.predict(0, top_k=5)
    assert top.shape == (5,) and int(top.max()) < n_i
