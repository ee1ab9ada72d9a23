"""Stage-wise CUDA-event timing of one epoch through the public host path (negatives -> gather -> plan -> train kernel),
first call at a new epoch shape vs repeated calls (run on the GPU box: PYTHONPATH=. python tools/host_path_times.py)."""
import time, torch, numpy as np
from torchrecsys_b200 import _lib
from torchrecsys_b200.collaborative.fm import FM
from torchrecsys_b200.engine import EpochRunner, step_scales
dev = torch.device("cuda:0")
U, I, C, D, B, K, W = 1_000_000, 200_000, 100, 64, 8192, 200, 20
net = FM(U, I, {"product_category": C}, D, use_metadata=True, use_cuda=True).to(dev)
opt = torch.optim.SparseAdam(list(net.parameters()), lr=1e-3)
runner = EpochRunner(net, opt)
rng = np.random.default_rng(1234)
user_all = torch.from_numpy(rng.integers(0, U, (K + W) * B)).to(dev)
pos_all = torch.from_numpy(rng.integers(0, I, (K + W) * B)).to(dev)
item_meta = (torch.arange(I, device=dev) % C).view(-1, 1).contiguous()
ev = lambda: torch.cuda.Event(enable_timing=True)
def block(user, pos, first, tag):
    es = [ev() for _ in range(6)]
    torch.cuda.synchronize()
    es[0].record()
    neg, neg_meta = _lib.philox_negatives(1234, first, pos, I, item_meta)
    es[1].record()
    smp = {"user": user, "pos": pos, "neg": neg, "pos_meta": item_meta[pos], "neg_meta": neg_meta}
    es[2].record()
    b = runner.binding
    model = net.abi_model(opt.state, b.keys)
    epoch = _lib.make_epoch(smp["user"], smp["pos"], smp["neg"], smp["pos_meta"], smp["neg_meta"], B)
    n_steps = user.shape[0] // B
    scales = torch.tensor(step_scales(b, n_steps), dtype=torch.float64).to(torch.float32).to(dev, non_blocking=True)
    optim = _lib.Optim(b.kind, 0, b.beta1, b.beta2, b.eps, scales.data_ptr())
    plan = _lib.plan_build(model, epoch, dev)
    es[3].record()
    ws = _lib.train_workspace(model, epoch, dev)
    loss = torch.empty(n_steps, dtype=torch.float32, device=dev)
    es[4].record()
    _lib.train_steps(model, epoch, optim, plan, ws, 0, n_steps, loss)
    es[5].record()
    torch.cuda.synchronize()
    t = [es[i].elapsed_time(es[i + 1]) * 1e3 for i in range(5)]
    print(tag, f"philox {t[0]:.0f} gather {t[1]:.0f} plan {t[2]:.0f} alloc {t[3]:.0f} train {t[4]:.0f} total {es[0].elapsed_time(es[5])*1e3:.0f} us -> {es[0].elapsed_time(es[5])*1e3/n_steps:.2f} us/step")
block(user_all[:W * B], pos_all[:W * B], 0, "warm W")
runner.reserve(K * B, B)
for i in range(4):
    block(user_all[W * B:], pos_all[W * B:], W * B, f"K run {i}")
