"""The reference's OTHER way through a training step (``-m gpu``): ``TorchRecSys.forward -> hinge_loss ->
TorchRecSys.backward(loss, optimizer)`` (model.py:171-200), i.e. scorer kernels under torch autograd and a torch
optimizer stepping on sparse gradients -- against the golden vectors recorded from the live reference, against the
fused path, and for optimizers the fused path does not know (SGD with momentum: engine.AutogradEpochRunner)."""
import numpy as np
import pytest
import torch

from oracle import cf_oracle as O
from tests import _golden as G
from tests.test_gpu_parity import TRAIN, _net_from_golden, _samples, _torch_opt

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _batch_dict(g, s, dev):
    b = {k: torch.from_numpy(v).to(dev) for k, v in G.batch_at(g, s).items()}
    out = {"user_id": b["user"], "pos_item_id": b["pos"], "neg_item_id": b["neg"]}
    if "pos_meta" in b:
        out["pos_metadata_id"], out["neg_metadata_id"] = b["pos_meta"], b["neg_meta"]
    return out


class _Owner:
    """TorchRecSys.forward / backward without the data pipeline (they only need use_cuda)."""
    use_cuda = True
    from torchrecsys_b200.model import TorchRecSys as _T
    forward, backward = _T.forward, _T.backward


@pytest.mark.parametrize("name", TRAIN)
def test_backward_gives_the_references_sparse_gradients(dev, name):
    from torchrecsys_b200.helper.loss import hinge_loss
    g = G.load(name)
    net_type, F, _ = G.parse_train_name(name)
    net = _net_from_golden(g, net_type, F, dev)
    pos, neg = _Owner().forward(net, _batch_dict(g, 0, dev))
    loss = hinge_loss(pos, neg)
    assert loss.requires_grad and abs(float(loss) - float(g["loss"][0])) < 1e-5
    loss.backward()
    _, want = (O.linear_grads if net_type == "linear" else O.fm_grads)(G.section(g, "init"), G.batch_at(g, 0))
    named = dict(net.named_parameters())
    for key, (idx, vals) in want.items():
        gr = named[key].grad
        assert gr is not None and gr.is_sparse, key   # as nn.Embedding(sparse=True) would hand to the optimizer
        rows, gsum = O.coalesce(idx, vals)
        c = gr.coalesce()
        np.testing.assert_array_equal(c.indices()[0].cpu().numpy(), rows)
        np.testing.assert_allclose(c.values().cpu().numpy(), gsum.reshape(c.values().shape), rtol=1e-4, atol=1e-7,
                                   err_msg=key)


@pytest.mark.parametrize("name", TRAIN)
def test_forward_backward_step_equals_fit_equals_golden(dev, name):
    """One step three ways: the user-facing forward/backward with a REAL torch optimizer, the fused kernel, the
    reference's recorded result."""
    from torchrecsys_b200.engine import EpochRunner
    from torchrecsys_b200.helper.loss import hinge_loss
    g = G.load(name)
    net_type, F, opt = G.parse_train_name(name)
    B = int(g["meta"][4])
    net_a = _net_from_golden(g, net_type, F, dev)
    optim_a = _torch_opt(opt, net_a, float(g["lr"]))
    owner = _Owner()
    pos, neg = owner.forward(net_a, _batch_dict(g, 0, dev))
    loss_a = owner.backward(hinge_loss(pos, neg), optim_a)     # zero_grad -> backward -> torch's own step -> item()
    net_b = _net_from_golden(g, net_type, F, dev)
    optim_b = _torch_opt(opt, net_b, float(g["lr"]))
    loss_b = EpochRunner(net_b, optim_b).run(_samples(g, 1, dev), B)
    assert abs(loss_a - float(g["loss"][0])) < 1e-5 and abs(float(loss_b[0]) - float(g["loss"][0])) < 1e-5
    init = G.section(g, "init")
    sd_a = {k: v.detach().cpu().numpy() for k, v in net_a.state_dict().items()}
    sd_b = {k: v.detach().cpu().numpy() for k, v in net_b.state_dict().items()}
    for k, v in G.section(g, "after1").items():
        ok = G.well_conditioned_rows(net_type, opt, init, G.batch_at(g, 0), k, v.shape[0])
        np.testing.assert_allclose(sd_a[k][ok], v[ok], rtol=1e-5, atol=2e-6, err_msg=f"autograd path {k}")
        np.testing.assert_allclose(sd_b[k][ok], v[ok], rtol=1e-5, atol=2e-6, err_msg=f"fused path {k}")


def test_unknown_optimizers_fall_back_to_the_autograd_loop(dev):
    """SGD with momentum is not row-sparse (a dense momentum buffer moves every row): engine.make_runner warns and runs
    the reference's loop body under autograd; the result equals torch's own optimizer on the numpy oracle's gradients."""
    from torchrecsys_b200.engine import AutogradEpochRunner, make_runner
    g = G.load("train_linear_F1_sgd")
    B, steps = int(g["meta"][4]), 3
    net = _net_from_golden(g, "linear", 1, dev)
    optim = torch.optim.SGD(net.parameters(), lr=0.05, momentum=0.9)
    with pytest.warns(UserWarning, match="falling back"):
        runner = make_runner(net, optim, False)
    assert isinstance(runner, AutogradEpochRunner)
    loss = runner.run(_samples(g, steps, dev), B).cpu().numpy()
    params = G.section(g, "init")
    buf = {k: np.zeros_like(v) for k, v in params.items()}
    want = []
    for s in range(steps):
        l, grads = O.linear_grads(params, G.batch_at(g, s))
        want.append(l)
        for k, (idx, vals) in grads.items():
            dense = np.zeros_like(params[k])
            np.add.at(dense, idx, vals.reshape(len(idx), -1))
            buf[k] = 0.9 * buf[k] + dense if s else dense
            params[k] -= 0.05 * buf[k]
    np.testing.assert_allclose(loss, np.array(want), rtol=1e-5, atol=1e-6)
    for k, v in net.state_dict().items():
        np.testing.assert_allclose(v.cpu().numpy(), params[k], rtol=1e-4, atol=1e-6, err_msg=k)


def test_mlp_forward_is_autograd_visible_and_matches_the_kernel_path(dev):
    from torchrecsys_b200.collaborative.mlp import MLP
    torch.manual_seed(0)
    net = MLP(300, 200, {}, 32, use_metadata=False, hidden_layers=[64, 32], use_cuda=True).to(dev).train()
    u = torch.randint(0, 300, (256,), device=dev)
    i = torch.randint(0, 200, (256,), device=dev)
    batch = {"user_id": u, "pos_item_id": i}
    out = net.forward(batch, "user_id", "pos_item_id")
    assert out.requires_grad and out.shape == (256, 1)
    out.sum().backward()
    assert net.user.weight.grad.is_sparse and net.fcs[0].weight.grad is not None
    with torch.no_grad():
        ker = net.forward(batch, "user_id", "pos_item_id")        # tcgen05 path, bf16 operands
    np.testing.assert_allclose(ker.cpu().numpy(), out.detach().cpu().numpy(), rtol=2e-2,
                               atol=2e-2 * float(out.abs().max()))


def test_out_of_range_ids_raise_index_error_like_aten_embedding(dev):
    g = G.load("train_fm_F1_sparse_adam")
    net = _net_from_golden(g, "fm", 1, dev)
    b = _batch_dict(g, 0, dev)
    bad = dict(b, user_id=b["user_id"].clone())
    bad["user_id"][3] = 10 ** 6
    with pytest.raises(IndexError):
        net.forward(bad, "user_id", "pos_item_id", "pos_metadata_id")
    bad = dict(b, pos_item_id=b["pos_item_id"].clone())
    bad["pos_item_id"][0] = -1
    with pytest.raises(IndexError):
        net.forward(bad, "user_id", "pos_item_id", "pos_metadata_id")
