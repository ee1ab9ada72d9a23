"""Per-CTA phase timeline of the fused train kernel on the C2 workload (debug; run on the GPU box).
Stamps (globaltimer ns) per step and CTA: 0 step start, 1 arrive barrier 1, 2 leave barrier 1,
3 arrive barrier 2."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["TRS_DEBUG_SKIP"] = str(64 | int(os.environ.get("TRS_DEBUG_SKIP", "0")))
from torchrecsys_b200 import _lib  # noqa: E402
from torchrecsys_b200.collaborative.fm import FM  # noqa: E402
from torchrecsys_b200 import engine  # noqa: E402

dev = torch.device("cuda:0")
B, K = int(os.environ.get("B", 8192)), 40
torch.manual_seed(0)
net = FM(1_000_000, 200_000, {"c": 100}, 64, use_metadata=True, use_cuda=True).to(dev)
opt = torch.optim.SparseAdam(list(net.parameters()), lr=1e-3)
rng = np.random.default_rng(0)
user = torch.from_numpy(rng.integers(0, 1_000_000, K * B)).to(dev)
pos = torch.from_numpy(rng.integers(0, 200_000, K * B)).to(dev)
item_meta = (torch.arange(200_000, device=dev) % 100).view(-1, 1).contiguous()
neg, neg_meta = _lib.philox_negatives(1, 0, pos, 200_000, item_meta)
b = engine.bind_optimizer(opt, list(net.parameters()))
model = net.abi_model(opt.state, b.keys)
pos_meta = item_meta[pos].contiguous()  # keep alive: make_epoch only stores the pointer
epoch = _lib.make_epoch(user, pos, neg, pos_meta, neg_meta, B)
scales = torch.tensor(engine.step_scales(b, K), dtype=torch.float64).float().to(dev)
optim = _lib.Optim(b.kind, 0, b.beta1, b.beta2, b.eps, scales.data_ptr())
plan = _lib.plan_build(model, epoch, dev)
ws = _lib.train_workspace(model, epoch, dev)
loss = torch.empty(K, device=dev)
for _ in range(3):
    _lib.train_steps(model, epoch, optim, plan, ws, 0, K, loss)
torch.cuda.synchronize()
L = _lib.lib()
L.trs_debug_trace_offset.restype = C.c_size_t
off = L.trs_debug_trace_offset(C.byref(model), C.byref(epoch))
sm, grid, _ = _lib.device_info()
tr = ws[off:off + K * grid * 128].cpu().numpy().view(np.uint64).reshape(K, grid, 16).astype(np.int64)
t0 = tr[:, :, 0].min(axis=1, keepdims=True)
rel = (tr - t0[:, :, None]) / 1e3  # us
names = {0: "step start", 8: "A: early fetch landed, meta issued", 9: "A: sample 0 loads done", 10: "A: sample 0 math done",
         11: "A: sample 0 updates issued", 12: "A: both samples done", 13: "A: long-seg preload done", 1: "arrive barrier1",
         2: "leave barrier1", 4: "B: next-step fetch issued", 5: "B: lin loads issued", 6: "B: long segs + lin done",
         7: "B: short segs done", 3: "arrive barrier2"}
order = [0, 8, 9, 10, 11, 12, 13, 1, 2, 4, 5, 6, 7, 3]
prev = None
for k in order:
    a = rel[5:, :, k]
    d = "" if prev is None else "  delta p50 %6.2f  mean-of-max-delta %6.2f" % (np.median(a - prev), (a - prev).max(axis=1).mean())
    print(f"  {names[k]:36s} p50 {np.median(a):6.2f}  mean-of-max {a.max(axis=1).mean():6.2f}{d}")
    prev = a
step = np.diff(tr[:, :, 0].min(axis=1)) / 1e3
print("step time us: p50 %.2f" % np.median(step[5:]))
