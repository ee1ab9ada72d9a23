#!/usr/bin/env python
"""bench.py -- headline benchmark of the TorchRecSys hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2_fm|c1_linear|c4_linear]

A "step" is one training step of the named workload: forward x2 + hinge + backward + sparse
optimizer update over one batch (reference model.py:274-284).  Default workload = BASELINE.json
configs[1] ("C2"): FM + 1 metadata feature (100 categories), dynamic negatives, 1M users x 200k
items, dim 64, batch 8192, SparseAdam, synthetic uniform ids.

  value      whole-job samples/s over K steps, every input already resident in HBM, timed with CUDA
             events; the timed region contains EVERYTHING a step needs: Philox negatives, the sort
             plan (coalesce's sort), and the persistent fused kernel.
  e2e        the same K steps with the ids starting in pinned HOST memory (H2D inside the timed
             region) and the per-step losses read back to the host (D2H inside).
  roofline   the fused train kernel alone: algorithmic HBM bytes (ids + each unique
             touched row's param+state read once and written once) / its CUDA-event duration,
             against the measured HBM peak of MEASURED_PEAKS.json.
  cpu_baseline  the reference's CPU op stream (oracle/torch_port.py, kind "port": /root/reference
             is not on the GPU box) on a bounded sample of the same workload on the host cores.

--impl reference runs only that CPU arm and prints it as the main line.
Multi-GPU (N>1, under torchrun): table rows are sharded by `row mod N`... see DESIGN.md (e);
each rank trains its own sample shard, weak scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: net, users, items, dim, n_categories (0 = no metadata), batch, optimizer, dynamic negs
    "c2_fm": dict(net="fm", n_users=1_000_000, n_items=200_000, dim=64, n_cat=100, batch=8192,
                  opt="sparse_adam", lr=1e-3, desc="BASELINE configs[1]: FM + product_category(100), "
                  "dynamic negatives, 1M users x 200k items, dim 64, batch 8192, SparseAdam"),
    "c1_linear": dict(net="linear", n_users=3000, n_items=1000, dim=80, n_cat=0, batch=1024,
                      opt="sparse_adam", lr=1e-3, desc="BASELINE configs[0] shape: linear 3k x 1k, dim 80, batch 1024"),
    "c3_mlp": dict(net="mlp", n_users=1_000_000, n_items=200_000, dim=64, n_cat=0, batch=16384, opt="adagrad",
                   lr=1e-2, hidden=[512, 256, 128], desc="BASELINE configs[2]: MLP [512,256,128] + batch norm, bf16 "
                   "tensor-core GEMMs, 1M users x 200k items, dim 64, batch 16384, Adagrad"),
    "c5_predict": dict(net="linear", n_users=1_000_000, n_items=5_000_000, dim=128, n_cat=0, batch=148 * 128, k=100,
                       opt=None, lr=0.0, predict=True, desc="BASELINE configs[4]: batched predict top-100 against a "
                       "5M-item table (linear scorer, dim 128); a step = 18944 users (148 user tiles)"),
    "c4_fused": dict(net="linear", n_users=50_000_000, n_items=5_000_000, dim=128, n_cat=0, batch=16384,
                     opt="sparse_adam", lr=1e-3, desc="BASELINE configs[3] tables (linear 50M x 5M, dim 128, batch 16384, "
                     "SparseAdam) on ONE GPU through the fused persistent kernel (84.5 GB of tables + state)"),
    "c4_linear": dict(net="linear", n_users=50_000_000, n_items=5_000_000, dim=128, n_cat=0, batch=16384,
                      opt="sparse_adam", lr=1e-3, sharded=True,
                      desc="BASELINE configs[3]: linear 50M x 5M, dim 128, batch 16384/GPU, tables row-sharded"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2_fm", choices=sorted(WORKLOADS))
    ap.add_argument("--users", type=int, default=0, help="override the workload's user count (memory-bound hosts)")
    ap.add_argument("--batch", type=int, default=0, help="override the workload's batch size")
    ap.add_argument("--dim", type=int, default=0, help="override the workload's n_factors")
    ap.add_argument("--zipf", type=float, default=0.0, help="draw training ids from Zipf(A), A > 1, instead of uniformly")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic data: uniform ids, rng seed 1234, category(item) = item mod n_cat
# ------------------------------------------------------------------------------------------------
def synth_ids(wl, n, seed=1234):
    """Uniform i.i.d. ids, or with --zipf A: rank-frequency Zipf(A) ids (hot rows, many duplicates)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    a = wl.get("zipf", 0.0)
    if a > 1.0:
        user = ((rng.zipf(a, n) - 1) % wl["n_users"]).astype(np.int64)
        pos = ((rng.zipf(a, n) - 1) % wl["n_items"]).astype(np.int64)
        return user, pos
    user = rng.integers(0, wl["n_users"], n, dtype=np.int64)
    pos = rng.integers(0, wl["n_items"], n, dtype=np.int64)
    return user, pos


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's op stream on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference(wl, steps, warmup, budget_s):
    """Times oracle/torch_port.py (same ATen CPU ops as the reference's use_cuda=False path) on a
    bounded sample: `steps` steps of the workload's batch size, thread count swept, best kept."""
    import numpy as np
    import torch
    from oracle import cf_oracle as O
    from oracle import torch_port as TP
    B = wl["batch"]
    cap_note = ""
    if wl["n_users"] > 5_000_000:  # host RAM: 50M x 128 fp32 + SparseAdam state would be 77 GB
        wl = dict(wl, n_users=5_000_000)
        cap_note = " (user table cut to 5M rows on the host: 25.6 GB + optimizer state does not fit; row access stays random)"
    n = (steps + warmup) * B
    user, pos = synth_ids(wl, n)
    neg = O.philox_negatives(1234, 0, pos, wl["n_items"])
    batch_of = lambda s: {"user": torch.from_numpy(user[s * B:(s + 1) * B]),
                          "pos": torch.from_numpy(pos[s * B:(s + 1) * B]),
                          "neg": torch.from_numpy(neg[s * B:(s + 1) * B])}
    metas = [wl["n_cat"]] if wl["n_cat"] else []
    ncores = len(os.sched_getaffinity(0))
    best = None
    t_start = time.time()
    threads = sorted({1, 2, 4, 8, 16, 32, ncores} & set(range(1, ncores + 1)))
    tried = {}
    for nt in threads:
        if time.time() - t_start > budget_s:
            break
        torch.set_num_threads(nt)
        torch.manual_seed(1234)
        kw = dict(hidden=wl["hidden"], batch_norm=True) if wl["net"] == "mlp" else {}
        net = TP.make_net(wl["net"], wl["n_users"], wl["n_items"], metas, wl["dim"], **kw)
        net.train()
        opt = TP.make_optimizer(wl["opt"], net, wl["lr"])

        def run(s):
            b = batch_of(s)
            if metas:
                b["pos_meta"] = (b["pos"] % wl["n_cat"]).view(-1, 1)
                b["neg_meta"] = (b["neg"] % wl["n_cat"]).view(-1, 1)
            return TP.train_step(net, opt, b)

        for s in range(warmup):
            run(s)
        t0 = time.perf_counter()
        done = 0
        for s in range(warmup, warmup + steps):
            run(s)
            done += 1
            if time.perf_counter() - t0 > budget_s / max(len(threads), 1) and done >= 2:
                break
        dt = time.perf_counter() - t0
        tried[nt] = done * B / dt
        if best is None or tried[nt] > best[0]:
            best = (tried[nt], nt, done, dt)
        del net, opt
    value, nt, done, dt = best
    return {"value": value, "unit": "samples/s", "cores": nt, "kind": "port",
            "sample": f"{done} steps of batch {B} ({done * B} samples) of the same workload after "
                      f"{warmup} warm-up steps, torch {torch.__version__} CPU, threads swept "
                      f"{ {k: round(v) for k, v in tried.items()} } on {ncores} cores, best kept" + cap_note,
            "ms_per_step": dt / done * 1e3, "steps": done}


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region (NVML, ~2 ms period)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))

                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            # NVML queries take a driver lock: dense at first (short timed regions still get a few samples), sparse
            # afterwards (a long region with many launches is not disturbed)
            time.sleep(0.002 if len(self.samples) < 6 else 0.01)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s), "sm_mhz_min": s[0] if s else None}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def algorithmic_bytes(wl, user, pos, neg, n_steps, B, S):
    """8 B per id consumed + per unique touched row (param + S state tensors) read once
    and written once, embedding row (dim floats) and its width-1 companion.  From the actual batches."""
    import torch
    D, F = wl["dim"], 1 if wl["n_cat"] else 0
    row_bytes = 4 * (D + (0 if wl["net"] == "mlp" else 1)) * (2 + 2 * S)  # MLP tables have no width-1 companion
    total = 0
    for s in range(n_steps):
        sl = slice(s * B, (s + 1) * B)
        uu = torch.unique(user[sl]).numel()
        items = torch.cat([pos[sl], neg[sl]])
        ui = torch.unique(items).numel()
        um = torch.unique(items % wl["n_cat"]).numel() if F else 0
        meta_row = row_bytes if wl["net"] == "fm" else 4 * D * (2 + 2 * S)
        total += 8 * (3 * B + 2 * B * F) + (uu + ui) * row_bytes + um * meta_row
    return total


def gpu_bench(args, wl):
    import torch
    import torch.distributed as dist
    from torchrecsys_b200 import _lib
    from torchrecsys_b200.collaborative.fm import FM
    from torchrecsys_b200.collaborative.linear import Linear
    from torchrecsys_b200.collaborative.mlp import MLP
    from torchrecsys_b200.engine import EpochRunner, MlpEpochRunner

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, W, B = args.steps, max(args.warmup, 3), args.batch or wl["batch"]
    F = 1 if wl["n_cat"] else 0

    torch.manual_seed(1234 + rank)
    is_mlp = wl["net"] == "mlp"
    if is_mlp:
        net = MLP(wl["n_users"], wl["n_items"], {}, wl["dim"], use_metadata=False, use_batch_norm=True,
                  hidden_layers=wl["hidden"], use_cuda=True).to(dev).train()
    else:
        cls = FM if wl["net"] == "fm" else Linear
        net = cls(wl["n_users"], wl["n_items"], {"product_category": wl["n_cat"]} if F else {}, wl["dim"],
                  use_metadata=bool(F), use_cuda=True).to(dev)
    if wl["opt"] == "sparse_adam":
        opt = torch.optim.SparseAdam(list(net.parameters()), lr=wl["lr"])
        S = 2
    else:
        opt = torch.optim.Adagrad(net.parameters(), lr=wl["lr"])
        S = 1
    runner = (MlpEpochRunner if is_mlp else EpochRunner)(net, opt)

    n = (K + W) * B
    user_h, pos_h = synth_ids(wl, n, seed=1234 + rank)
    user_h, pos_h = torch.from_numpy(user_h).pin_memory(), torch.from_numpy(pos_h).pin_memory()
    user, pos = user_h.to(dev), pos_h.to(dev)
    item_meta = (torch.arange(wl["n_items"], device=dev) % wl["n_cat"]).view(-1, 1).contiguous() if F else None

    def step_block(u, p, first):
        """Everything the steps need, device side: negatives -> plan -> fused kernel."""
        neg, neg_meta = _lib.philox_negatives(1234, first, p, wl["n_items"], item_meta)
        smp = {"user": u, "pos": p, "neg": neg}
        if F:
            smp["pos_meta"] = item_meta[p]  # category of the positive (torch gather: loader-side plumbing)
            smp["neg_meta"] = neg_meta
        return runner.run(smp, B)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    sync_all = (lambda: (torch.cuda.synchronize(), dist.barrier(), torch.cuda.synchronize())) if world > 1 \
        else torch.cuda.synchronize

    # clock ramp (not steps): keep the GPU busy ~0.3 s so the timed region does not start at idle clocks
    scratch = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    t0 = time.time()
    while time.time() - t0 < 0.3:
        scratch.copy_(scratch.flip(0))
        torch.cuda.synchronize()
    del scratch

    # W warm-up steps
    step_block(user[:W * B], pos[:W * B], 0)
    sync_all()

    # memory pool: an epoch of K steps needs other buffer sizes than the W warm-up steps did.  Reserve them now, so
    # that the timed region measures the steps and not torch's first cudaMalloc of each size (in a real fit() every
    # epoch has the same shape and reuses the blocks of the one before)
    if hasattr(runner, "reserve"):
        runner.reserve(K * B, B)
    pool = [torch.empty(K * B, dtype=torch.int64, device=dev) for _ in range(1 + 2 * F)] + [torch.empty(K, device=dev)]
    del pool
    sync_all()

    # ---- timed: K steps, inputs resident in HBM ----
    uK, pK = user[W * B:], pos[W * B:]
    launches0 = runner.launches
    e0, e1 = ev(), ev()
    with ClockSampler(local) as clocks:
        sync_all()
        e0.record()
        loss = step_block(uK, pK, W * B)
        e1.record()
        sync_all()
    ms = e0.elapsed_time(e1)
    launches = runner.launches - launches0 + 1  # + the Philox kernel
    mean_loss = float(loss.mean().item())

    # ---- the fused kernel alone (roofline): same K steps again on the now-further-trained model ----
    neg, neg_meta = _lib.philox_negatives(1234, W * B, pK, wl["n_items"], item_meta)
    smp = {"user": uK, "pos": pK, "neg": neg}
    if F:
        smp["pos_meta"], smp["neg_meta"] = item_meta[pK], neg_meta
    b = runner.binding
    model = net.abi_model(opt.state, b.keys)
    epoch = _lib.make_epoch(smp["user"], smp["pos"], smp["neg"], smp.get("pos_meta"), smp.get("neg_meta"), B)
    from torchrecsys_b200.engine import step_scales, advance_steps
    scales = torch.tensor(step_scales(b, K), dtype=torch.float64).float().to(dev)
    optim_c = _lib.Optim(b.kind, 0, b.beta1, b.beta2, b.eps, scales.data_ptr())
    p0, p1, k0, k1 = ev(), ev(), ev(), ev()
    torch.cuda.synchronize()
    p0.record()
    plan = _lib.plan_build(model, epoch, dev)
    p1.record()
    loss2 = torch.empty(K, device=dev)
    if is_mlp:
        mlp_c = net.abi_mlp(runner.grads, opt.state if b.keys[0] else None, b.keys[0])
        ws = _lib.mlp_train_workspace(model, mlp_c, epoch, dev)
        k0.record()
        _lib.mlp_train_steps(model, mlp_c, epoch, optim_c, plan, ws, 0, K, loss2)
        k1.record()
    else:
        ws = _lib.train_workspace(model, epoch, dev)
        k0.record()
        _lib.train_steps(model, epoch, optim_c, plan, ws, 0, K, loss2)
        k1.record()
    torch.cuda.synchronize()
    advance_steps(opt, runner.params, b, K)
    kernel_ms, plan_ms = k0.elapsed_time(k1), p0.elapsed_time(p1)
    alg_bytes = algorithmic_bytes(wl, uK, pK, neg, K, B, S)

    # ---- e2e: ids start in pinned host memory, losses end on the host ----
    loss_host = torch.empty(K, dtype=torch.float32).pin_memory()
    e2e_ms = None
    for attempt in range(2):  # first pass untimed: it warms the allocator for this region's buffer sizes
        x0, x1 = ev(), ev()
        sync_all()
        x0.record()
        u_d = user_h[W * B:].to(dev, non_blocking=True)
        p_d = pos_h[W * B:].to(dev, non_blocking=True)
        loss3 = step_block(u_d, p_d, (W + K * (1 + attempt)) * B)
        loss_host.copy_(loss3, non_blocking=True)
        x1.record()
        sync_all()
        e2e_ms = x0.elapsed_time(x1)
    assert loss_host.numel() == K

    times = torch.tensor([ms, e2e_ms, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms, kernel_ms = (float(x) for x in times.cpu())

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "kernel": "trs::train_kernel (one persistent launch, K steps)",
                "kernel_ms_per_step": kernel_ms / K, "plan_ms_per_step": plan_ms / K,
                "algorithmic_bytes_per_step": alg_bytes / K,
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650"}
    if is_mlp:
        # the tower's GEMMs bound this workload: 3 x 2 x MACs per row, 2 rows per sample
        width, macs = wl["dim"] * 2, 0
        for h in wl["hidden"]:
            macs += width * h
            width = h
        flops_step = 3 * 2 * (macs + width) * 2 * B
        tpeak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        tach = flops_step * K / (kernel_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": tach, "peak": tpeak, "unit": "TFLOP/s", "frac": tach / tpeak,
                    "kernel": "trs::gemm_tn_kernel (9 tcgen05 GEMMs per step) timed inside the whole fused step: "
                              "every kernel of trs_mlp_train_steps is in the denominator",
                    "kernel_ms_per_step": kernel_ms / K, "plan_ms_per_step": plan_ms / K,
                    "algorithmic_flops_per_step": flops_step,
                    "embedding_bytes_per_step": alg_bytes / K,
                    "peak_source": "measured sustained bf16 (MEASURED_PEAKS.json)" if peaks else "fallback 1400"}
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
    except Exception:
        pass
    out = {
        "metric": "train samples/sec (fwd+bwd+sparse update)", "value": world * K * B / (ms * 1e-3),
        "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if is_mlp else "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "name": args.workload, "batch_per_gpu": B, "global_batch": B * world,
                   "optimizer": wl["opt"], "l2": "inputs larger than L2: tables+optimizer state "
                   f"{(wl['n_users'] + wl['n_items']) * (wl['dim'] + 1) * 4 * (1 + S) / 1e9:.2f} GB, rows hit at random",
                   "timed_region": ("Philox negatives + sort plan + K fused MLP steps (one C call, ~40 kernels per step)"
                                    if is_mlp else
                                    "Philox negatives + sort plan + persistent fused train kernel, K steps in one launch"),
                   "parallelism": f"dp{world} (independent replicas)" if world > 1 else "single GPU",
                   "clock_ramp": "0.3 s of device copies before the warm-up steps",
                   "memory_pool": "buffer sizes of a K-step epoch reserved before the timed region (no cudaMalloc inside)",
                   "mean_loss": mean_loss},
        "e2e": {"value": world * K * B / (e2e_ms * 1e-3), "unit": "samples/s",
                "h2d_bytes_per_step": 16 * B, "d2h_bytes_per_step": 4,
                "note": "user+positive ids from pinned host memory; metadata ids and negatives are derived on the device"},
        "gpu_launches": launches,
        # traffic: dram bytes of the dominant kernel per launch, from the committed ncu --set full capture
        # (profiles/traffic.json holds bytes per step; one launch = K steps)
        "roofline": dict(roofline, traffic=(traffic * K if traffic else None)),
        "clocks": clocks.summary(),
    }
    if world > 1:
        dist.destroy_process_group()
    return out, rank


# ------------------------------------------------------------------------------------------------
# predict / top-k (BASELINE configs[4])
# ------------------------------------------------------------------------------------------------
def cpu_predict(wl, budget_s):
    """The reference's predict (model.py:341-452: score every item for one user, sort, slice) on the host
    cores through oracle/torch_port.py (same ATen ops, without the per-chunk pandas frame)."""
    import torch
    from oracle import torch_port as TP
    ncores = len(os.sched_getaffinity(0))
    torch.set_num_threads(ncores)
    torch.manual_seed(1234)
    net = TP.make_net("linear", 1024, wl["n_items"], [], wl["dim"]).eval()
    t0, done = time.perf_counter(), 0
    while done < 3 or (time.perf_counter() - t0 < budget_s and done < 64):
        TP.predict_topk(net, done % 1024, wl["n_items"], wl["k"], chunk=1 << 20)
        done += 1
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": "users/s", "cores": ncores, "kind": "port",
            "sample": f"{done} users x all {wl['n_items']} items, top-{wl['k']}, torch {torch.__version__} CPU, "
                      f"{ncores} threads (user table cut to 1024 rows: it is not on the path)",
            "ms_per_step": dt / done * 1e3, "steps": done}


def _traffic(workload):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(workload)
    except Exception:
        return None


def predict_bench(args, wl):
    import torch
    from torchrecsys_b200 import _lib
    from torchrecsys_b200.collaborative.linear import Linear
    import torch.distributed as dist
    from torchrecsys_b200 import sharded as S
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, W, B, k = args.steps, max(args.warmup, 3), args.batch or wl["batch"], wl["k"]
    torch.manual_seed(1234)
    # N > 1: the item table is cut into contiguous blocks, one per rank; user rows are replicated
    lo_item, hi_item = S.item_block(wl["n_items"], rank, world)
    net = Linear(wl["n_users"], hi_item - lo_item, {}, wl["dim"], use_metadata=False, use_cuda=True).to(dev).eval()
    with torch.no_grad():
        net.item_bias.weight.normal_(0, 0.01)
    model = net.abi_model()
    if world > 1:
        single = _lib.predict_topk

        def local_topk(u, kk, offset):
            idx, score, over = single(model, u, kk, item_offset=offset)
            return idx, score

        class _Sharded:  # same call shape as _lib.predict_topk
            @staticmethod
            def predict_topk(_model, u, kk):
                idx, score = S.sharded_predict_topk(local_topk, u, kk, wl["n_items"])
                return idx, score, torch.zeros(1, dtype=torch.int32, device=dev)
        _lib = _Sharded
    import numpy as np
    rng = np.random.default_rng(1234)
    users_h = torch.from_numpy(rng.integers(0, wl["n_users"], (K + W) * B)).pin_memory()
    users = users_h.to(dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    # clock ramp (not steps): keep the GPU busy ~0.3 s so the timed region does not start at idle clocks
    scratch = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    t0 = time.time()
    while time.time() - t0 < 0.3:
        scratch.copy_(scratch.flip(0))
        torch.cuda.synchronize()
    del scratch
    for s in range(W):
        _lib.predict_topk(model, users[s * B:(s + 1) * B], k)
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    with ClockSampler(dev.index) as clocks:
        e0.record()
        for s in range(W, W + K):
            idx, score, over = _lib.predict_topk(model, users[s * B:(s + 1) * B], k)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    n_over = int(over.sum())
    x0, x1 = ev(), ev()
    out_h = torch.empty((B, k), dtype=torch.int64).pin_memory()
    x0.record()
    for s in range(W, W + K):
        u = users_h[s * B:(s + 1) * B].to(dev, non_blocking=True)
        idx, score, over = _lib.predict_topk(model, u, k)
        out_h.copy_(idx, non_blocking=True)
    x1.record()
    torch.cuda.synchronize()
    e2e_ms = x0.elapsed_time(x1)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tpeak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    flops = 2.0 * wl["n_items"] * wl["dim"] * B * K
    tach = flops / (ms * 1e-3) / 1e12 / world   # per GPU
    if world > 1:
        t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = (float(x) for x in t.cpu())
        dist.destroy_process_group()
    return {
        "metric": "predict top-k users/sec", "value": K * B / (ms * 1e-3), "unit": "users/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": wl["desc"], "name": args.workload, "users_per_step": B, "top_k": k,
                   "parallelism": (f"item table in {world} contiguous blocks, local top-k per rank, all-gather + "
                                   "k-way merge (the same users on every rank)") if world > 1 else "single GPU",
                   "l2": "inputs larger than L2: bf16 item operand 1.44 GB streamed per user wave",
                   "timed_region": "operand preparation (fp32 tables -> bf16 [w,c] rows) + tcgen05 score/top-k kernel "
                                   "+ exact fp32 re-scoring, every step", "overflow_users_last_step": n_over},
        "e2e": {"value": K * B / (e2e_ms * 1e-3), "unit": "users/s", "h2d_bytes_per_step": 8 * B,
                "d2h_bytes_per_step": 8 * B * k},
        "gpu_launches": 7 * K,
        "roofline": {"bound": "tensor", "achieved": tach, "peak": tpeak, "unit": "TFLOP/s", "frac": tach / tpeak,
                     "traffic": _traffic(args.workload), "kernel": "trs::topk_score_kernel, timed inside the whole predict call "
                     "(preparation and re-scoring are in the denominator)",
                     "algorithmic_flops_per_step": flops / K,
                     "peak_source": "measured sustained bf16 (MEASURED_PEAKS.json)" if peaks else "fallback 1400"},
        "clocks": clocks.summary(),
    }


# ------------------------------------------------------------------------------------------------
# row-sharded large-table training (BASELINE configs[3]) -- torchrun, one rank per GPU
# ------------------------------------------------------------------------------------------------
def sharded_bench(args, wl):
    import numpy as np
    import torch
    import torch.distributed as dist
    from torchrecsys_b200 import sharded as S
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if not dist.is_initialized():
        if world == 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    K, W, B = args.steps, max(args.warmup, 3), args.batch or wl["batch"]
    tr = S.ShardedLinearTrainer(wl["n_users"], wl["n_items"], wl["dim"], optimizer=wl["opt"], lr=wl["lr"], device=dev)
    rng = np.random.default_rng(1234 + rank)
    ids_h = [torch.from_numpy(rng.integers(0, n, (K + W) * B)).pin_memory()
             for n in (wl["n_users"], wl["n_items"], wl["n_items"])]
    ids = [t.to(dev) for t in ids_h]
    step = lambda src, s: tr.train_step(*(t[s * B:(s + 1) * B] for t in src))
    ev = lambda: torch.cuda.Event(enable_timing=True)
    sync_all = lambda: (torch.cuda.synchronize(), dist.barrier(), torch.cuda.synchronize())
    for s in range(W):
        step(ids, s)
    e0, e1 = ev(), ev()
    hs = torch.zeros(1, device=dev)
    with ClockSampler(local) as clocks:
        sync_all()
        e0.record()
        for s in range(W, W + K):
            hs += step(ids, s)
        e1.record()
        sync_all()
    ms = e0.elapsed_time(e1)
    x0, x1 = ev(), ev()
    hs2 = torch.zeros(1, device=dev)
    sync_all()
    x0.record()
    for s in range(W, W + K):
        hs2 += step([t[s * B:(s + 1) * B].to(dev, non_blocking=True) for t in ids_h], 0)
    loss_h = hs2.cpu()
    x1.record()
    sync_all()
    e2e_ms = x0.elapsed_time(x1)
    times = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dist.all_reduce(hs)
    ms, e2e_ms = (float(x) for x in times.cpu())
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    S_ = {"sgd": 0, "adagrad": 1, "sparse_adam": 2}[wl["opt"]]
    # no-duplicate bound (uniform ids over 50M / 5M rows): 3 rows of (dim+1) floats, param + S states, read + write
    alg = K * B * (24 + 3 * (wl["dim"] + 1) * 4 * (2 + 2 * S_))
    ach = alg / (ms * 1e-3) / 1e9
    nvl = K * B * 3 * (8 + 2 * (wl["dim"] + 1) * 4) * (world - 1) / world  # ids out, rows back, gradient rows out
    out = {
        "metric": "train samples/sec (fwd+bwd+sparse update)", "value": world * K * B / (ms * 1e-3), "unit": "samples/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "name": args.workload, "batch_per_gpu": B, "global_batch": B * world,
                   "optimizer": wl["opt"], "parallelism": f"tables row-sharded over {world} ranks (row % G), "
                   "all_to_all of ids / rows / gradient rows per step, owner-side coalesce + update",
                   "l2": f"inputs larger than L2: {(wl['n_users'] + wl['n_items']) * (wl['dim'] + 1) * 4 * (1 + S_) / world / 1e9:.1f} "
                         "GB of tables + optimizer state per rank, rows hit at random",
                   "mean_loss": float(hs) / (world * K * B)},
        "e2e": {"value": world * K * B / (e2e_ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": 24 * B,
                "d2h_bytes_per_step": 4},
        "gpu_launches": K * 14,
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                     "kernel": "per rank: gather + trs_linear_rows_step + trs_sparse_row_update (the whole step incl. "
                               "routing and NCCL is in the denominator)",
                     "algorithmic_bytes_per_step": alg / K, "nvlink_bytes_per_step_per_rank": nvl / K,
                     "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650"},
        "clocks": clocks.summary(),
    }
    dist.destroy_process_group()
    return out, rank


def main():
    args = parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.batch:
        wl["batch"] = args.batch
    if args.users:
        wl["n_users"] = args.users
    if args.dim:
        wl["dim"] = args.dim
        wl["desc"] += f" [n_factors overridden: {args.dim}]"
    if args.zipf:
        wl["zipf"] = args.zipf
        wl["desc"] += f" [ids drawn from Zipf({args.zipf})]"
    rank = int(os.environ.get("RANK", "0"))
    if wl.get("predict"):
        if rank != 0 and args.impl == "reference":
            return
        if args.impl == "reference":
            r = cpu_predict(wl, max(args.cpu_seconds, 20.0) * 3)
            print(json.dumps({"impl": "reference", "metric": "predict top-k users/sec", "value": r["value"],
                              "unit": "users/s", "n_gpus": args.gpus, "steps": r["steps"], "warmup": 0,
                              "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                              "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                              "config": {"workload": wl["desc"], "name": args.workload},
                              "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                              "e2e": {"value": r["value"], "unit": "users/s", "h2d_bytes_per_step": 0,
                                      "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
            return
        out = predict_bench(args, wl)
        if rank != 0:
            return
        if not args.no_cpu_baseline and args.gpus == 1:
            r = cpu_predict(wl, args.cpu_seconds)
            out["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(out))
        return
    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, args.steps)
        r = cpu_reference(wl, steps, max(1, min(args.warmup, 3)), budget_s=max(args.cpu_seconds, 20.0) * 3)
        line = {"impl": "reference", "metric": "train samples/sec (fwd+bwd+sparse update)", "value": r["value"],
                "unit": "samples/s", "n_gpus": args.gpus, "steps": r["steps"], "warmup": max(1, min(args.warmup, 3)),
                "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": wl["desc"], "name": args.workload, "batch_per_gpu": wl["batch"]},
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return
    out, rank = sharded_bench(args, wl) if wl.get("sharded") else gpu_bench(args, wl)
    if rank == 0:
        if args.gpus == 1 and not args.no_cpu_baseline:
            r = cpu_reference(wl, steps=8, warmup=1, budget_s=args.cpu_seconds)
            out["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(out))


if __name__ == "__main__":
    main()
