"""MLP tower (tcgen05 GEMMs + BatchNorm/ReLU kernels + dense and row-wise updates) on the B200 (``-m gpu``).

Two yardsticks:

* ``tests/_mlp_emul.py`` -- the reference step restated in fp32 torch with a bf16 rounding exactly where the
  kernels store bf16.  The CUDA path may differ from it only by fp32 summation order (and the 1-ulp bf16
  differences that causes downstream): gradient tensors must agree to cosine > 0.9995, entry-wise to 1e-2 of
  the tensor's gradient scale on the small golden problems.  This checks the kernels' LOGIC.
* the golden vectors recorded from the live fp32 reference and the fp32 torch port (oracle/).  Here the
  difference is bf16 arithmetic itself: scores rtol 2e-2.  Gradients of this loss are
  differences of a positive and a negative pass of nearly equal size (hinge: g*(d s_neg - d s_pos), and at
  initialisation every hinge is active), so one bf16 rounding (2^-9) of the per-pass quantities shows up as
  10-35 % of the much smaller difference; the bound is cosine > 0.97 and 0.5 of the gradient scale, and
  tests/test_mlp_emul.py shows the same level for the emulation on the CPU.  This checks the TOLERANCE claim.

BatchNorm makes the bias in front of it gradient-free up to rounding noise (sum(dZ) == 0 analytically) and
the forward result independent of it, so ``fcs.*.bias`` is left out when batch norm is on."""
import numpy as np
import pytest
import torch

from oracle import cf_oracle as O
from tests import _golden as G
from tests import _mlp_emul as E

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _net_from_golden(g, dev, F):
    from torchrecsys_b200.collaborative.mlp import MLP
    U, I, C, D = (int(x) for x in g["meta"][:4])
    net = MLP(U, I, {f"m{f}": C for f in range(F)}, D, use_metadata=F > 0, use_batch_norm=True,
              hidden_layers=[int(h) for h in g["hidden"]], use_cuda=True)
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in G.section(g, "init").items()})
    return net.to(dev)


def _batch(g, s, dev, F):
    b = {k: torch.from_numpy(v).to(dev) for k, v in G.batch_at(g, s).items()}
    d = {"user_id": b["user"], "pos_item_id": b["pos"], "neg_item_id": b["neg"]}
    if F:
        d["pos_metadata_id"], d["neg_metadata_id"] = b["pos_meta"], b["neg_meta"]
    return d


def _skip(k, use_bn=True):
    return "num_batches" in k or (use_bn and k.startswith("fcs.") and k.endswith(".bias"))


def _compare_deltas(got, want, init, keys, tol, cos_min, label, scale_of=None):
    """max |got_delta - want_delta| <= tol * scale and cosine >= cos_min, per tensor; returns problems."""
    problems, report = [], []
    for k in keys:
        wd = (want[k] - init[k]).float().flatten()
        gd = (got[k] - init[k]).float().flatten()
        scale = float(wd.abs().max())
        if scale_of and k in scale_of:
            scale = max(scale, scale_of[k])
        err = float((gd - wd).abs().max())
        cos = float(torch.nn.functional.cosine_similarity(gd, wd, dim=0)) if scale > 0 else 1.0
        report.append(f"{label} {k}: err {err:.3e} scale {scale:.3e} cos {cos:.5f}")
        if err > tol * scale + 1e-8 or cos < cos_min:
            problems.append(report[-1])
    print("\n".join(report))
    return problems


@pytest.mark.parametrize("name", G.names("train_mlp_"))
def test_forward_matches_reference_scores(dev, name):
    g = G.load(name)
    F = int(g["meta"][6])
    net = _net_from_golden(g, dev, F)
    batch = _batch(g, 0, dev, F)
    scale = float(np.abs(g["pos0"]).max())
    # train mode: this batch's BatchNorm statistics, running statistics updated once per pass
    net.train()
    with torch.no_grad():  # the tcgen05 path (with a graph being recorded forward runs torch ops: test_gpu_autograd.py)
        pos = net.forward(batch, "user_id", "pos_item_id", "pos_metadata_id")
        neg = net.forward(batch, "user_id", "neg_item_id", "neg_metadata_id")
    assert pos.shape == (64, 1)
    np.testing.assert_allclose(pos.cpu().numpy(), g["pos0"], rtol=2e-2, atol=2e-2 * scale)
    np.testing.assert_allclose(neg.cpu().numpy(), g["neg0"], rtol=2e-2, atol=2e-2 * scale)
    assert int(net.bns[0].num_batches_tracked) == 2
    # the running statistics after those two passes == the oracle's
    params = G.section(g, "init")
    bt = G.batch_at(g, 0)
    O.mlp_scores(params, bt["user"], bt["pos"], bt.get("pos_meta"), train=True)
    O.mlp_scores(params, bt["user"], bt["neg"], bt.get("neg_meta"), train=True)
    for l in range(2):
        for k in ("running_mean", "running_var"):
            got = getattr(net.bns[l], k).cpu().numpy()
            np.testing.assert_allclose(got, params[f"bns.{l}.{k}"], rtol=2e-2, atol=2e-3)
    # eval mode: running statistics
    net.eval()
    with torch.no_grad():
        got = net.forward(batch, "user_id", "pos_item_id", "pos_metadata_id").cpu().numpy()
    want = O.mlp_scores(params, bt["user"], bt["pos"], bt.get("pos_meta"), train=False)
    np.testing.assert_allclose(got, want, rtol=2e-2, atol=2e-2 * float(np.abs(want).max()))


def _golden_samples(g, dev):
    smp = {}
    for k in ("user", "pos", "neg", "pos_meta", "neg_meta"):
        if f"batch/{k}" in g:
            v = g[f"batch/{k}"]
            smp[k] = torch.from_numpy(v.reshape((-1,) + v.shape[2:])).contiguous()
    return smp, {k: v.to(dev) for k, v in smp.items()}


@pytest.mark.parametrize("name", G.names("train_mlp_"))
def test_sgd_steps_match_bf16_emulation_and_reference(dev, name):
    """Golden problems stepped with SGD: 1 and 5 steps against the bf16 emulation (tight), 1 step against
    the fp32 golden / oracle gradients (bf16 tolerance)."""
    from torchrecsys_b200.engine import MlpEpochRunner
    g = G.load(name)
    U, I, C, D, B, steps, F = (int(x) for x in g["meta"])
    lr = 0.5
    smp_cpu, smp = _golden_samples(g, dev)
    init = {k: torch.from_numpy(np.asarray(v)).clone() for k, v in G.section(g, "init").items()}
    keys = [k for k in init if not _skip(k)]
    problems = []
    for n_steps in (1, steps):
        net = _net_from_golden(g, dev, F).train()
        loss = MlpEpochRunner(net, torch.optim.SGD(net.parameters(), lr=lr)).run(
            {k: v[:n_steps * B] for k, v in smp.items()}, B).cpu().numpy()
        got = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        emu = {k: v.clone() for k, v in init.items()}
        want_loss = [E.sgd_step_bf16(emu, {k: v[s * B:(s + 1) * B] for k, v in smp_cpu.items()}, lr)
                     for s in range(n_steps)]
        np.testing.assert_allclose(loss, np.array(want_loss), rtol=2e-3, atol=1e-4)
        # trajectories of two bf16 computations drift apart slowly: 1e-2 per step
        problems += _compare_deltas(got, emu, init, keys, 1e-2 * n_steps, 0.9995 if n_steps == 1 else 0.995,
                                    f"[emul {n_steps} step]")
        assert int(got["bns.0.num_batches_tracked"]) == 2 * n_steps
    # one step against the fp32 oracle (closed-form backward pinned to the live reference in test_oracle_golden)
    net = _net_from_golden(g, dev, F).train()
    MlpEpochRunner(net, torch.optim.SGD(net.parameters(), lr=lr)).run({k: v[:B] for k, v in smp.items()}, B)
    got = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    params = G.section(g, "init")
    state = O.init_opt_state(params, O.OptSpec("sgd", lr=lr))
    O.mlp_train_step(params, state, G.batch_at(g, 0), O.OptSpec("sgd", lr=lr), 1)
    ref = {k: torch.from_numpy(np.asarray(v)) for k, v in params.items()}
    _, sparse0, _ = O.mlp_grads(G.section(g, "init"), G.batch_at(g, 0))
    emb_scale = {k: lr * float(np.abs(v).max()) for k, (_, v) in sparse0.items()}
    problems += _compare_deltas(got, ref, init, [k for k in keys if "running" not in k], 0.5, 0.97, "[fp32 1 step]",
                                emb_scale)
    assert not problems, "\n".join(problems)


@pytest.mark.parametrize("name", [n for n in G.names("train_mlp_") if n.endswith("adagrad")])
def test_adagrad_steps_track_reference(dev, name):
    """The recorded Adagrad run of the live reference.  Its first step is lr*g/(|g|+1e-10) = lr*sign(g): an
    entry whose gradient lies within the bf16 error of zero may flip, every other entry must take the same
    step; untouched rows must not move; over 5 steps the loss curve matches to 2e-2 and the parameters stay
    within the per-step bound."""
    from torchrecsys_b200.engine import MlpEpochRunner
    g = G.load(name)
    U, I, C, D, B, steps, F = (int(x) for x in g["meta"])
    lr = float(g["lr"])
    _, smp = _golden_samples(g, dev)
    init = G.section(g, "init")
    problems = []
    for n_steps, section in ((1, "after1"), (steps, "final")):
        net = _net_from_golden(g, dev, F).train()
        opt = torch.optim.Adagrad(net.parameters(), lr=lr)
        loss = MlpEpochRunner(net, opt).run({k: v[:n_steps * B] for k, v in smp.items()}, B).cpu().numpy()
        np.testing.assert_allclose(loss, g["loss"][:n_steps], rtol=2e-2, atol=1e-3)
        got = {k: v.detach().cpu().numpy() for k, v in net.state_dict().items()}
        for k, w in G.section(g, section).items():
            if _skip(k):
                continue
            if "running" in k:
                # after 5 Adagrad steps the flipped +-lr entries have moved the activations themselves
                if not np.allclose(got[k], w, rtol=3e-2, atol=3e-3 if n_steps == 1 else 0.3):
                    problems.append(f"{k} ({section}): running statistics differ by {np.abs(got[k] - w).max():.3e}")
                continue
            delta = w - init[k]
            touched = delta != 0
            if not (got[k][~touched] == init[k][~touched]).all():
                problems.append(f"{k} ({section}): entries the reference did not touch have moved")
            if n_steps == 1 and touched.any():
                same = np.abs((got[k] - init[k]) - delta) <= 0.05 * lr
                flips = int((~same[touched]).sum())
                if flips > max(2, 0.15 * touched.sum()):
                    problems.append(f"{k} ({section}): {flips} of {int(touched.sum())} first steps differ")
            elif touched.any():
                if np.abs(got[k] - w).max() > 2 * lr * n_steps:
                    problems.append(f"{k} ({section}): drift {np.abs(got[k] - w).max():.3e} beyond the step bound")
        # optimizer state stays torch-compatible
        assert float(opt.state[net.user.weight]["step"]) == n_steps
        opt.load_state_dict(opt.state_dict())
    assert not problems, "\n".join(problems)


def test_gradients_at_config3_shape(dev):
    """C3's tower ([512,256,128] + BN, D=64) on a 1000-sample step (padding rows in play): dense and embedding
    gradients, read back as SGD lr=1 parameter deltas, against the bf16 emulation (tight) and the fp32 torch
    port of the reference (bf16 tolerance)."""
    from oracle import torch_port as TP
    from torchrecsys_b200.collaborative.mlp import MLP
    from torchrecsys_b200.engine import MlpEpochRunner
    torch.manual_seed(7)
    U, I, D, B, hidden = 5000, 3000, 64, 1000, [512, 256, 128]
    ref = TP.MLPPort(U, I, [], D, hidden=hidden, batch_norm=True)
    with torch.no_grad():
        ref.user.weight.normal_(0, 0.5)
        ref.item.weight.normal_(0, 0.5)
    net = MLP(U, I, {}, D, use_metadata=False, use_batch_norm=True, hidden_layers=hidden, use_cuda=True)
    net.load_state_dict(ref.state_dict())
    net = net.to(dev).train()
    init = {k: v.clone() for k, v in ref.state_dict().items()}
    rng = np.random.default_rng(11)
    batch = {"user": torch.from_numpy(rng.integers(0, U, B)), "pos": torch.from_numpy(rng.integers(0, I, B)),
             "neg": torch.from_numpy(rng.integers(0, I, B))}
    emu = {k: v.clone() for k, v in init.items()}
    emu_loss = E.sgd_step_bf16(emu, batch, 1.0)
    ref.train()
    ref_loss = TP.train_step(ref, torch.optim.SGD(ref.parameters(), lr=1.0), batch)
    loss = MlpEpochRunner(net, torch.optim.SGD(net.parameters(), lr=1.0)).run(
        {k: v.to(dev) for k, v in batch.items()}, B)
    assert abs(float(loss[0]) - emu_loss) < 1e-3 and abs(float(loss[0]) - ref_loss) < 2e-2
    got = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    keys = [k for k in init if not _skip(k)]
    # single entries differ by up to ~15 % of the scale even between two bf16 computations (a 1-ulp bf16
    # difference in Z flips relu masks and rounding boundaries, and the pos/neg cancellation amplifies it);
    # the tensors as a whole agree to cosine 0.9999
    problems = _compare_deltas(got, emu, init, keys, 0.25, 0.9995, "[emul]")
    item_scale = float((ref.state_dict()["item.weight"] - init["item.weight"]).abs().max())
    problems += _compare_deltas(got, ref.state_dict(), init, [k for k in keys if "running" not in k], 0.5, 0.97,
                                "[fp32]", {"user.weight": item_scale})
    assert not problems, "\n".join(problems)


def test_mlp_without_batch_norm_and_ragged_last_batch(dev):
    """use_batch_norm=False, a batch size that is not a multiple of 128 and a short last batch."""
    from torchrecsys_b200.collaborative.mlp import MLP
    from torchrecsys_b200.engine import MlpEpochRunner
    from oracle import torch_port as TP
    torch.manual_seed(3)
    U, I, D, B, n, hidden = 300, 200, 16, 200, 520, [64, 32]
    ref = TP.MLPPort(U, I, [], D, hidden=hidden, batch_norm=False)
    with torch.no_grad():
        ref.user.weight.normal_(0, 0.7)
        ref.item.weight.normal_(0, 0.7)
    net = MLP(U, I, {}, D, use_metadata=False, use_batch_norm=False, hidden_layers=hidden, use_cuda=True)
    net.load_state_dict(ref.state_dict())
    net = net.to(dev).train()
    init = {k: v.clone() for k, v in ref.state_dict().items()}
    rng = np.random.default_rng(5)
    smp = {"user": torch.from_numpy(rng.integers(0, U, n)), "pos": torch.from_numpy(rng.integers(0, I, n)),
           "neg": torch.from_numpy(rng.integers(0, I, n))}
    emu = {k: v.clone() for k, v in init.items()}
    want = [E.sgd_step_bf16(emu, {k: v[lo:lo + B] for k, v in smp.items()}, 0.1, use_bn=False) for lo in range(0, n, B)]
    loss = MlpEpochRunner(net, torch.optim.SGD(net.parameters(), lr=0.1)).run({k: v.to(dev) for k, v in smp.items()}, B)
    np.testing.assert_allclose(loss.cpu().numpy(), np.array(want), rtol=2e-3, atol=1e-4)
    got = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    problems = _compare_deltas(got, emu, init, list(init), 3e-2, 0.995, "[emul 3 steps]")
    assert not problems, "\n".join(problems)


def test_gradient_error_is_measured_against_the_references_own_amp_run(dev):
    """The bound on the bf16 tower's gradients is not self-defined: the reference's OWN reduced-precision path --
    ``use_amp=True``: ``torch.cuda.amp.autocast`` (fp16) + GradScaler, model.py:86-88, 192-195 -- run here on the same
    B200, the same C3-shaped tower, initial weights and batch, deviates from its fp32 gradients by a measurable
    relative error; so does a bf16 autocast of it (what the reference's AMP would be with the dtype north_star names).
    The fused tcgen05 step must be as close to the fp32 gradients as that bf16 autocast run, tensor by tensor (factor
    1.5 for the different accumulation order), and never further than 2x the fp16 run's error where fp16 does not
    overflow."""
    from oracle import torch_port as TP
    from torchrecsys_b200.collaborative.mlp import MLP
    from torchrecsys_b200.engine import MlpEpochRunner
    torch.manual_seed(7)
    U, I, D, B, hidden = 5000, 3000, 64, 4096, [512, 256, 128]
    ref = TP.MLPPort(U, I, [], D, hidden=hidden, batch_norm=True)
    with torch.no_grad():
        ref.user.weight.normal_(0, 0.5)
        ref.item.weight.normal_(0, 0.5)
    init = {k: v.clone() for k, v in ref.state_dict().items()}
    rng = np.random.default_rng(11)
    batch = {k: torch.from_numpy(rng.integers(0, n, B)).to(dev) for k, n in (("user", U), ("pos", I), ("neg", I))}

    def port_grads(autocast_dtype):
        net = TP.MLPPort(U, I, [], D, hidden=hidden, batch_norm=True)
        net.load_state_dict(init)
        net = net.to(dev).train()
        ctx = torch.autocast("cuda", dtype=autocast_dtype) if autocast_dtype is not None else torch.autocast("cuda", enabled=False)
        with ctx:
            loss = TP.hinge(net(batch["user"], batch["pos"]), net(batch["user"], batch["neg"]))
        scale = 1024.0 if autocast_dtype is torch.float16 else 1.0     # GradScaler: scale, backward, unscale
        (loss * scale).backward()
        return {k: (p.grad.to_dense() if p.grad.is_sparse else p.grad).float() / scale
                for k, p in net.named_parameters()}

    g32, g16, gbf = port_grads(None), port_grads(torch.float16), port_grads(torch.bfloat16)
    net = MLP(U, I, {}, D, use_metadata=False, use_batch_norm=True, hidden_layers=hidden, use_cuda=True)
    net.load_state_dict(init)
    net = net.to(dev).train()
    MlpEpochRunner(net, torch.optim.SGD(net.parameters(), lr=1.0)).run(batch, B)
    ours = {k: (init[k].to(dev) - p.detach()) for k, p in net.named_parameters()}      # lr = 1: the delta IS the gradient

    def rel(a, b):
        return float((a - b).norm() / b.norm().clamp_min(1e-30))

    lines, bad = [], []
    for k in g32:
        if _skip(k):
            continue
        e16, ebf, eus = rel(g16[k], g32[k]), rel(gbf[k], g32[k]), rel(ours[k], g32[k])
        cos = float(torch.nn.functional.cosine_similarity(ours[k].flatten(), g32[k].flatten(), dim=0))
        lines.append(f"{k:28s} rel. error vs fp32: reference fp16 AMP {e16:.3e}, bf16 autocast {ebf:.3e}, this kernel {eus:.3e} (cos {cos:.5f})")
        if not (eus <= 1.5 * ebf + 1e-4):
            bad.append(f"{k}: {eus:.3e} > 1.5 x bf16 autocast {ebf:.3e}")
    print("\n".join(lines))
    assert not bad, "\n".join(bad + lines)
