"""Record golden vectors from the LIVE, UNMODIFIED reference at /root/reference.

Run in the build container only (the reference tree does not travel to the GPU box):

    python oracle/make_golden.py            # writes tests/golden/*.npz

The reference does not import as shipped (collaborative/mlp.py:16 uses ``List`` without
importing it); ``builtins.List`` is injected before the import, nothing
under /root/reference is touched.  Training fixtures drive the reference's own
``TorchRecSys.forward`` / ``TorchRecSys.backward`` methods (model.py:171-200) and its scorer
modules on explicit batches, bypassing only the data loader's RNG.
"""
import builtins
import os
import sys
import typing

import numpy as np

builtins.List = typing.List
REF = os.environ.get("TRS_REF", "/root/reference")
sys.path.insert(0, REF)

import pandas as pd  # noqa: E402
import torch  # noqa: E402
from torchrecsys.collaborative.fm import FM  # noqa: E402
from torchrecsys.collaborative.linear import Linear  # noqa: E402
from torchrecsys.collaborative.mlp import MLP  # noqa: E402
from torchrecsys.evaluate.metrics import Metrics  # noqa: E402
from torchrecsys.helper.loss import hinge_loss  # noqa: E402
from torchrecsys.model import TorchRecSys  # noqa: E402

assert os.path.realpath(sys.modules["torchrecsys"].__path__[0]).startswith(os.path.realpath(REF))

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
U, I, C, D, B, STEPS = 50, 30, 7, 16, 64, 20
LR = {"sgd": 0.05, "adagrad": 0.05, "sparse_adam": 0.01}


def bare_model():
    """A TorchRecSys shell (no dataset) so its own forward/backward methods can be driven."""
    m = TorchRecSys.__new__(TorchRecSys)
    torch.nn.Module.__init__(m)
    m.use_cuda, m.use_amp, m.grad_scaler = False, False, None
    return m


def make_opt(kind, net):
    if kind == "sparse_adam":
        return torch.optim.SparseAdam(list(net.parameters()), lr=LR[kind])
    if kind == "adagrad":
        return torch.optim.Adagrad(net.parameters(), lr=LR[kind])
    return torch.optim.SGD(net.parameters(), lr=LR[kind])


def make_batches(rng, steps, F):
    b = {
        "user": rng.integers(0, U, (steps, B)),
        "pos": rng.integers(0, I, (steps, B)),
        "neg": rng.integers(0, I, (steps, B)),
    }
    if F:
        # item -> category map, so a negative's metadata is a function of the item
        cat = rng.integers(0, C, (F, I))
        b["pos_meta"] = np.stack([cat[f][b["pos"]] for f in range(F)], axis=-1)
        b["neg_meta"] = np.stack([cat[f][b["neg"]] for f in range(F)], axis=-1)
    return {k: v.astype(np.int64) for k, v in b.items()}


def ref_batch(b, s, F):
    """Batch dict in the layout the reference loader produces (dataset.py:281-287):
    F=1 -> (B, L=1); F>=2 -> (B, F, L=1)."""
    d = {"user_id": torch.from_numpy(b["user"][s]),
         "pos_item_id": torch.from_numpy(b["pos"][s]),
         "neg_item_id": torch.from_numpy(b["neg"][s])}
    if F == 1:
        d["pos_metadata_id"] = torch.from_numpy(b["pos_meta"][s])
        d["neg_metadata_id"] = torch.from_numpy(b["neg_meta"][s])
    elif F > 1:
        d["pos_metadata_id"] = torch.from_numpy(b["pos_meta"][s])[:, :, None]
        d["neg_metadata_id"] = torch.from_numpy(b["neg_meta"][s])[:, :, None]
    return d


def snapshot(net, prefix, out):
    for k, v in net.state_dict().items():
        out[f"{prefix}/{k}"] = v.detach().numpy().copy()


def train_fixture(net_type, F, opt_kind, init_scale, seed, hidden=None):
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    n_meta = {f"m{f}": C for f in range(F)}
    if net_type == "linear":
        net = Linear(U, I, n_meta, D, use_metadata=F > 0)
    elif net_type == "fm":
        net = FM(U, I, n_meta, D, use_metadata=F > 0)
    else:
        net = MLP(U, I, n_meta, D, use_metadata=F > 0, use_batch_norm=True, hidden_layers=hidden)
    with torch.no_grad():
        for name, p in net.named_parameters():
            if init_scale is not None and ("fcs" not in name and "bns" not in name and "output" not in name):
                p.copy_(torch.randn_like(p) * init_scale)
    net.train()
    model = bare_model()
    opt = make_opt(opt_kind, net)
    steps = STEPS if net_type != "mlp" else 5
    b = make_batches(rng, steps, F)
    out = {f"batch/{k}": v for k, v in b.items()}
    out["meta"] = np.array([U, I, C, D, B, steps, F], dtype=np.int64)
    out["lr"] = np.float64(LR[opt_kind])
    if hidden:
        out["hidden"] = np.array(hidden, dtype=np.int64)
    snapshot(net, "init", out)
    losses = []
    for s in range(steps):
        batch = ref_batch(b, s, F)
        pos, neg = model.forward(net, batch)
        if s == 0:
            out["pos0"], out["neg0"] = pos.detach().numpy().copy(), neg.detach().numpy().copy()
        losses.append(model.backward(hinge_loss(pos, neg), opt))
        if s == 0:
            snapshot(net, "after1", out)
    snapshot(net, "final", out)
    out["loss"] = np.array(losses, dtype=np.float32)
    for name, p in net.named_parameters():
        for sk, sv in opt.state[p].items():
            if torch.is_tensor(sv) and sv.dim() > 0:
                out[f"state/{name}/{sk}"] = sv.detach().numpy().copy()
            else:
                out[f"state/{name}/{sk}"] = np.float64(float(sv))
    return out


def predict_fixture(net_type, seed):
    """Exact-arithmetic weights (multiples of 1/8, small range) so every product and partial
    sum of the Linear scorer is exact in fp32 in any summation order -> bit-exact scores and
    top-k.  Drives the reference's real ``predict`` (model.py:341-452)."""
    rng = np.random.default_rng(seed)
    nu, ni = 10, 25
    df = pd.DataFrame({"user_id": np.resize(np.arange(nu), 100), "item_id": np.resize(np.arange(ni), 100)})
    np.random.seed(seed)
    torch.manual_seed(seed)
    model = TorchRecSys(df, "user_id", "item_id", n_factors=D, net_type=net_type, use_cuda=False)
    assert model.n_users == nu and model.n_items == ni
    with torch.no_grad():
        for name, p in model.net.named_parameters():
            if net_type != "mlp" or name.startswith(("user", "item")):
                p.copy_(torch.from_numpy(rng.integers(-4, 5, tuple(p.shape)).astype(np.float32) / 8))
    if net_type == "mlp":  # make eval-mode BN non-trivial
        for bn in model.net.bns:
            bn.running_mean.copy_(torch.randn_like(bn.running_mean) * 0.1)
            bn.running_var.copy_(torch.rand_like(bn.running_var) + 0.5)
    out = {}
    snapshot(model.net, "init", out)
    model.net.eval()
    k = 5
    ref_top, stable_top, scores = [], [], []
    for u in range(nu):
        ref_top.append(model.predict(u, top_k=k, prediction_batch_size=7).numpy())
        batch = {"user_id": torch.full((ni,), u, dtype=torch.long), "pos_item_id": torch.arange(ni)}
        with torch.no_grad():
            s = model.net.forward(batch, "user_id", "pos_item_id", None).squeeze().float()
        scores.append(s.numpy().copy())
        stable_top.append(torch.sort(s, descending=True, stable=True)[1][:k].numpy())
    out["scores"] = np.stack(scores)
    out["ref_topk"] = np.stack(ref_top)
    out["stable_topk"] = np.stack(stable_top)
    out["meta"] = np.array([nu, ni, D, k], dtype=np.int64)
    return out


def eval_fixture(seed):
    rng = np.random.default_rng(seed)
    pos = (rng.integers(-8, 9, 257) / 4).astype(np.float32)
    neg = (rng.integers(-8, 9, 257) / 4).astype(np.float32)  # many exact ties
    tp, tn = torch.from_numpy(pos).view(-1, 1), torch.from_numpy(neg).view(-1, 1)
    return {"pos": pos, "neg": neg,
            "auc": np.float32(Metrics().auc_score(tp, tn).item()),
            "hinge": np.float32(hinge_loss(tp, tn).item())}


def main():
    os.makedirs(OUT, exist_ok=True)
    jobs = []
    for net_type in ("linear", "fm"):
        for F in (0, 1):
            for opt in ("sgd", "adagrad", "sparse_adam"):
                jobs.append((f"train_{net_type}_F{F}_{opt}", (net_type, F, opt, 0.5 if net_type == "linear" else 0.15)))
    jobs.append(("train_fm_F2_sparse_adam", ("fm", 2, "sparse_adam", 0.15)))
    jobs.append(("train_linear_F0_sparse_adam_refinit", ("linear", 0, "sparse_adam", None)))
    # FM scale 0.15: with 0.3 the D=16 pairwise term saturates the sigmoid for some samples, their
    # gradient drops to ~1e-10 and Adagrad/Adam's g/(|g|+eps) turns rounding noise into +-lr steps --
    # a fixture no two implementations (nor the reference at two thread counts) agree on.
    for i, (name, (nt, F, opt, sc)) in enumerate(jobs):
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **train_fixture(nt, F, opt, sc, 100 + i))
        print("wrote", name)
    for i, (F, opt) in enumerate([(0, "adagrad"), (0, "sgd"), (1, "adagrad")]):
        name = f"train_mlp_F{F}_{opt}"
        np.savez_compressed(os.path.join(OUT, name + ".npz"),
                            **train_fixture("mlp", F, opt, 0.5, 200 + i, hidden=[32, 16]))
        print("wrote", name)
    for i, nt in enumerate(("linear", "fm", "mlp")):
        np.savez_compressed(os.path.join(OUT, f"predict_{nt}.npz"), **predict_fixture(nt, 300 + i))
        print("wrote predict", nt)
    np.savez_compressed(os.path.join(OUT, "eval_pairwise.npz"), **eval_fixture(400))
    print("torch", torch.__version__, "numpy", np.__version__)


if __name__ == "__main__":
    main()
