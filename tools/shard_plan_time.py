"""Time trs_shard_plan_build for ONE rank of a group of W ranks on the C4 shape (global batch 16384 * W), on one GPU:
what the routing / sort / dirty-flag plan costs per step as the group grows.
    python tools/shard_plan_time.py [--steps 20]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torchrecsys_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--users", type=int, default=50_000_000)
ap.add_argument("--items", type=int, default=5_000_000)
a = ap.parse_args()
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
for W in (1, 2, 4, 8):
    B = 16384 * W
    n = a.steps * B
    ids = [torch.from_numpy(rng.integers(0, m, n)).to(dev) for m in (a.users, a.items, a.items)]
    sh = _lib.Shard()
    sh.rank, sh.world, sh.dim, sh.n_users, sh.n_items = 0, W, 128, a.users, a.items
    ep = _lib.make_epoch(*ids, None, None, B)
    plan = _lib.shard_plan_build(sh, ep, dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record()
        plan = _lib.shard_plan_build(sh, ep, dev, plan)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"world {W}: global batch {B}, {a.steps} steps: plan {best * 1e3:.0f} us = {best * 1e3 / a.steps:.1f} us per step", flush=True)
