"""Host logic of the peer-mapped row-sharded trainer (torchrecsys_b200/sharded.py) without a GPU: the arena layout,
shard ownership (row r -> rank r % G, local row r // G), checkpoint scatter / gather in the reference's
single-process layout -- in-process (``emulate_world``) and over a world-size-2 gloo group -- and that training
without CUDA raises instead of falling back."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from torchrecsys_b200.sharded import ShardedLinearTrainer, shard_rows


def _full(rng, U, I, D):
    return {"user.weight": torch.from_numpy(rng.normal(0, 1, (U, D)).astype(np.float32)),
            "item.weight": torch.from_numpy(rng.normal(0, 1, (I, D)).astype(np.float32)),
            "user_bias.weight": torch.from_numpy(rng.normal(0, 1, (U, 1)).astype(np.float32)),
            "item_bias.weight": torch.from_numpy(rng.normal(0, 1, (I, 1)).astype(np.float32))}


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_rows_partition_every_table(world):
    for n in (0, 1, 7, 8, 9, 1000, 50_000_000):
        assert sum(shard_rows(n, r, world) for r in range(world)) == n
        for r in range(world):
            assert shard_rows(n, r, world) == len(range(r, n, world))


@pytest.mark.parametrize("world,opt", [(1, "sgd"), (2, "adagrad"), (3, "sparse_adam"), (8, "sparse_adam")])
def test_checkpoint_round_trip_on_cpu_arenas(world, opt):
    U, I, D = 103, 41, 8
    rng = np.random.default_rng(world)
    full = _full(rng, U, I, D)
    tr = ShardedLinearTrainer(U, I, D, global_batch=64, optimizer=opt, device="cpu", emulate_world=world)
    # arenas: every piece 256-byte aligned, nothing overlaps, staging and barrier words inside
    L = tr.layout
    offs = sorted([L.emb["user"], L.emb["item"], L.lin["user"], L.lin["item"], L.stage, L.sync]
                  + L.emb_state["user"] + L.emb_state["item"] + L.lin_state["user"] + L.lin_state["item"])
    assert all(o % 256 == 0 for o in offs) and len(set(offs)) == len(offs) and L.total >= L.sync + 4096
    for r in range(world):
        assert not tr.arenas[r][L.sync:L.sync + 4096].any()
    state = {k: {name: torch.from_numpy(rng.random(tuple(v.shape)).astype(np.float32))
                 for name in {"sgd": (), "adagrad": ("sum",), "sparse_adam": ("exp_avg", "exp_avg_sq")}[opt]}
             for k, v in full.items()}
    for k in state:
        state[k]["step"] = 17
    tr.load_state_dict(full, state)
    assert tr.binding.step0 == 17
    for r in range(world):  # ownership: local row l of rank r is global row l * G + r
        assert torch.equal(tr.tables[r]["item"][0], full["item.weight"][r::world])
        assert torch.equal(tr.tables[r]["user"][1], full["user_bias.weight"][r::world])
    back, sback = tr.state_dict(), tr.optimizer_state_dict()
    for k in full:
        assert torch.equal(back[k], full[k]), k
        for name, v in state[k].items():
            if name != "step":
                assert torch.equal(sback[k][name], v), (k, name)
    # a different group size reads the same checkpoint
    other = ShardedLinearTrainer(U, I, D, global_batch=64, optimizer=opt, device="cpu", emulate_world=world % 3 + 1)
    other.load_state_dict(back, sback)
    assert torch.equal(other.state_dict()["user.weight"], full["user.weight"])


def test_training_without_cuda_raises():
    tr = ShardedLinearTrainer(10, 10, 8, global_batch=4, device="cpu", emulate_world=2)
    ids = torch.zeros(8, dtype=torch.int64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tr.train_epoch(ids, ids, ids, 4)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=2)
    try:
        torch.set_num_threads(1)
        U, I, D = 57, 23, 4
        full = _full(np.random.default_rng(1), U, I, D)
        tr = ShardedLinearTrainer(U, I, D, global_batch=16, optimizer="adagrad", device="cpu")
        assert (tr.rank, tr.world, tr.local_ranks) == (rank, 2, [rank])
        tr.load_state_dict(full)
        assert torch.equal(tr.tables[rank]["user"][0], full["user.weight"][rank::2])
        tr.state[rank]["item"][0][0].fill_(float(rank + 1))        # Adagrad sums differ per rank
        sd, osd = tr.state_dict(), tr.optimizer_state_dict()       # all_gather over gloo
        ok = all(torch.equal(sd[k], full[k]) for k in full)
        want = torch.tensor([1.0, 2.0]).repeat(I)[:I, None].expand(I, D)
        ok = ok and torch.equal(osd["item.weight"]["sum"], want)
        # the loaders' id exchange: [steps, columns, B] per rank -> global columns, step-major, rank-major in a step
        steps, B = 3, 5
        g = torch.Generator().manual_seed(7)
        parts = [torch.randint(0, U, (steps, 2, B), generator=g) for _ in range(2)]
        cols = tr.gather_epoch(parts[rank])
        both = torch.stack(parts)                                  # [world, steps, 2, B]
        for j in range(2):
            ok = ok and cols[j].dtype == torch.int64 and torch.equal(cols[j], both[:, :, j].permute(1, 0, 2).reshape(-1))
        # a fitted model moves onto the two ranks and back (to_model gathers over the group: both ranks get everything)
        from torchrecsys_b200.collaborative.linear import Linear
        torch.manual_seed(11)
        net = Linear(U, I, {}, D, use_metadata=False, use_cuda=False)
        opt = torch.optim.Adagrad(net.parameters(), lr=0.1)
        for p in net.parameters():
            opt.state[p]["sum"].uniform_(0.0, 1.0)
            opt.state[p]["step"] += 5
        tr2 = ShardedLinearTrainer.from_model(net, opt, global_batch=16, device="cpu")
        ok = ok and tr2.binding.step0 == 5 and torch.equal(tr2.tables[rank]["user"][0], net.user.weight.detach()[rank::2])
        ok = ok and torch.equal(tr2.state[rank]["item"][0][0], opt.state[net.item.weight]["sum"][rank::2])
        net2 = Linear(U, I, {}, D, use_metadata=False, use_cuda=False)
        opt2 = torch.optim.Adagrad(net2.parameters(), lr=0.1)
        tr2.to_model(net2, opt2)
        for (k, a), (_, b) in zip(net.named_parameters(), net2.named_parameters()):
            ok = ok and torch.equal(a, b) and torch.equal(opt.state[a]["sum"], opt2.state[b]["sum"])
            ok = ok and float(opt2.state[b]["step"]) == 5.0
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_checkpoint_gather_over_a_gloo_group_of_two():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = dict(q.get(timeout=5) for _ in range(2))
    assert got == {0: True, 1: True}


@pytest.mark.parametrize("opt_name", ["SparseAdam", "Adagrad", "SGD"])
def test_bridge_from_a_linear_model_and_its_torch_optimizer_and_back(opt_name):
    """from_model shards TorchRecSys(net_type='linear').net together with the state of the torch optimizer bound to it
    (hyper-parameters, moments, step count); to_model hands everything back in torch's own conventions, so that a torch
    step taken afterwards equals the step the original pair would have taken."""
    from torchrecsys_b200.collaborative.linear import Linear
    U, I, D = 61, 29, 8
    torch.manual_seed(3)

    def make():
        net = Linear(U, I, {}, D, use_metadata=False, use_cuda=False)
        params = list(net.parameters())
        opt = {"SparseAdam": lambda: torch.optim.SparseAdam(params, lr=0.05, betas=(0.8, 0.95), eps=1e-7),
               "Adagrad": lambda: torch.optim.Adagrad(params, lr=0.05, lr_decay=0.01, eps=1e-9),
               "SGD": lambda: torch.optim.SGD(params, lr=0.05)}[opt_name]()
        return net, opt

    def torch_step(net, opt, seed):
        g = torch.Generator().manual_seed(seed)
        opt.zero_grad()
        for p in net.parameters():                       # a sparse gradient on a few rows, as nn.Embedding(sparse=True) gives
            rows = torch.randint(0, p.shape[0], (7,), generator=g)
            p.grad = torch.sparse_coo_tensor(rows[None], torch.randn(7, p.shape[1], generator=g), p.shape).coalesce()
        opt.step()

    net, opt = make()
    for s in range(3):
        torch_step(net, opt, s)
    tr = ShardedLinearTrainer.from_model(net, opt, global_batch=32, device="cpu", emulate_world=3)
    assert tr.kind == {"SparseAdam": "sparse_adam", "Adagrad": "adagrad", "SGD": "sgd"}[opt_name]
    b = tr.binding
    assert b.lr == 0.05 and b.step0 == (0 if opt_name == "SGD" else 3)
    if opt_name == "SparseAdam":
        assert (b.beta1, b.beta2, b.eps) == (0.8, 0.95, 1e-7)
    if opt_name == "Adagrad":
        assert (b.lr_decay, b.eps) == (0.01, 1e-9)
    for r in range(3):
        assert torch.equal(tr.tables[r]["item"][0], net.item.weight.detach()[r::3])
    net2, opt2 = make()
    tr.to_model(net2, opt2)
    for (k, p), (_, q) in zip(net.named_parameters(), net2.named_parameters()):
        assert torch.equal(p, q), k
        for name, v in opt.state[p].items():
            w = opt2.state[q][name]
            assert type(v) is type(w) and (torch.equal(v, w) if torch.is_tensor(v) else v == w), (k, name)
    torch_step(net, opt, 99)
    torch_step(net2, opt2, 99)
    for (k, p), (_, q) in zip(net.named_parameters(), net2.named_parameters()):
        assert torch.equal(p, q), k
    # not shardable: metadata tables, optimizers without a row-wise update
    with pytest.raises(ValueError):
        ShardedLinearTrainer.from_model(Linear(U, I, {"c": 5}, D, use_metadata=True), opt, 32, device="cpu", emulate_world=2)
    with pytest.raises(NotImplementedError):
        ShardedLinearTrainer.from_model(net, torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9), 32, device="cpu",
                                        emulate_world=2)


@pytest.mark.parametrize("n", [1, 7, 8, 9, 64, 150, 500, 1420, 5000, 100_000])
def test_host_fed_epochs_are_cut_into_a_few_growing_chunks(n):
    cap, b = ShardedLinearTrainer.host_chunks(n)
    sizes = np.diff(b)
    assert b[0] == 0 and b[-1] == n and (sizes > 0).all() and sizes.max() <= cap <= 2048
    assert sizes[0] <= max(8, -(-n // 16))                      # the exposed first copy is short
    assert all(sizes[i + 1] <= 4 * sizes[i] for i in range(len(sizes) - 1))   # each copy hides behind the chunk before
    assert len(sizes) <= 3 + n // 2048 + 1                      # and the epoch is not shredded
    cap, b = ShardedLinearTrainer.host_chunks(150, 64)
    assert (cap, b) == (64, [0, 10, 50, 114, 150])


def test_host_fed_epoch_control_flow_with_stub_streams(monkeypatch):
    """The chunk loop of train_epoch_host with the device pieces stubbed out (streams / events as no-ops, the training
    call recorded): every step is handed over exactly once and in order, negatives are keyed by the global sample index
    of the chunk's first sample, each chunk gets ITS slice of the optimizer scalars, losses land in the host buffer."""
    import contextlib

    class _Stub:
        def wait_stream(self, s): pass
        def wait_event(self, e): pass
        def record(self, s=None): pass

    monkeypatch.setattr(torch.cuda, "Stream", lambda *a, **k: _Stub())
    monkeypatch.setattr(torch.cuda, "Event", lambda *a, **k: _Stub())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _Stub())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    B, steps = 4, 150
    tr = ShardedLinearTrainer(50, 20, 8, global_batch=B, device="cpu", emulate_world=1)
    tr.bases = {0: 0}                                           # pretend the arena is mapped
    calls = []

    def fake_train_epoch(user, pos, neg, batch, check=True, timing=False, _scales=None):
        n = user.shape[0] // batch
        calls.append((user.clone(), pos.clone(), neg.clone(), _scales.clone(), check))
        tr.binding.step0 += n
        return user.view(n, batch).float().mean(1)

    tr.train_epoch = fake_train_epoch
    tr._step_scales = lambda n: torch.arange(n, dtype=torch.float32)
    ids = torch.arange(steps * 2 * B).view(steps, 2, B)
    loss_host = torch.zeros(steps)
    out = tr.train_epoch_host(ids, lambda pos, first: pos + first, loss_host, chunk_steps=64)
    _, bounds = ShardedLinearTrainer.host_chunks(steps, 64)
    assert len(calls) == len(bounds) - 1 and not any(c[4] for c in calls)
    assert torch.equal(torch.cat([c[0] for c in calls]), ids[:, 0].reshape(-1))
    assert torch.equal(torch.cat([c[1] for c in calls]), ids[:, 1].reshape(-1))
    assert torch.equal(torch.cat([c[3] for c in calls]), torch.arange(steps, dtype=torch.float32))
    for c, lo in zip(calls, bounds):
        assert torch.equal(c[2], c[1] + lo * B)                 # first global sample of the chunk
    want = ids[:, 0].float().mean(1)
    assert torch.equal(out, want) and torch.equal(loss_host, want) and tr.binding.step0 == steps
