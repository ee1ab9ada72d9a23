"""How much of the host-to-device id copy ShardedLinearTrainer.train_epoch_host hides, against the chunk size (one GPU,
C4 shape): resident epoch vs host-fed epoch in chunks of --chunks steps.
    python tools/host_feed_time.py [--steps 1400] [--chunks 1400,700,350,175,88,44]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torchrecsys_b200 import _lib  # noqa: E402
from torchrecsys_b200.sharded import ShardedLinearTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--users", type=int, default=50_000_000)
ap.add_argument("--items", type=int, default=5_000_000)
ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--batch", type=int, default=16384)
ap.add_argument("--steps", type=int, default=1400)
ap.add_argument("--chunks", default="1400,700,350,175,88,44")
a = ap.parse_args()
dev = torch.device("cuda", 0)
tr = ShardedLinearTrainer(a.users, a.items, a.dim, global_batch=a.batch, device=dev, emulate_world=1)
rng = np.random.default_rng(0)
ids = torch.from_numpy(np.stack([rng.integers(0, a.users, (a.steps, a.batch)),
                                 rng.integers(0, a.items, (a.steps, a.batch))], 1)).pin_memory()
draw = lambda pos, first: _lib.philox_negatives(7, first, pos, a.items)[0]


def timed(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    fn()
    e1.record()
    host_ms = (time.perf_counter() - t0) * 1e3     # how long the host needed to QUEUE the epoch
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), host_ms


ids_d = ids.to(dev)
u, p = ids_d[:, 0].reshape(-1), ids_d[:, 1].reshape(-1)
ms, host = timed(lambda: tr.train_epoch(u, p, draw(p, 0), a.batch, check=False))
print(f"resident, one launch: {ms:.2f} ms = {ms / a.steps * 1e3:.2f} us / step (host queued it in {host:.2f} ms)")
for c in (int(x) for x in a.chunks.split(",")):
    ms, host = timed(lambda: tr.train_epoch_host(ids, draw, None, chunk_steps=c))
    print(f"host-fed, chunks of up to {c} steps: {ms:.2f} ms = {ms / a.steps * 1e3:.2f} us / step "
          f"(host queued it in {host:.2f} ms)", flush=True)
tr.check_status()
if os.environ.get("TRS_HOST_PROFILE"):    # which host call waits for the device?
    import cProfile
    import pstats
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    tr.train_epoch_host(ids, draw, None, chunk_steps=350)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(14)
