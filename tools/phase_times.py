"""Attribute the fused train kernel's time to its phases on the C2 workload (timing only: the
TRS_DEBUG_SKIP variants compute wrong results on purpose).  Run on the GPU box."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
variants = [("full", {}), ("no_phaseA", {"TRS_DEBUG_SKIP": "1"}), ("no_long", {"TRS_DEBUG_SKIP": "2"}),
            ("no_short", {"TRS_DEBUG_SKIP": "4"}), ("no_lin", {"TRS_DEBUG_SKIP": "32"}),
            ("no_phaseB", {"TRS_DEBUG_SKIP": "38"}), ("no_sync", {"TRS_DEBUG_SKIP": "8"}),
            ("only_sync", {"TRS_DEBUG_SKIP": "55"}), ("no_prefetch", {"TRS_DEBUG_SKIP": "16"}),
            ("no_fused_update", {"TRS_DEBUG_SKIP": "128"}), ("no_fused_lin", {"TRS_DEBUG_SKIP": "256"}),
            ("A_only_no_fused", {"TRS_DEBUG_SKIP": str(2 + 4 + 32 + 128)}), ("A_only", {"TRS_DEBUG_SKIP": str(2 + 4 + 32)})]
extra = sys.argv[1:]
for name, env in variants:
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "200", "--warmup", "20",
                        "--no-cpu-baseline", *extra], env=dict(os.environ, **env), capture_output=True, text=True)
    try:
        j = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"{name:10s} kernel {j['roofline']['kernel_ms_per_step'] * 1e3:8.2f} us/step   "
              f"step {j['ms_per_step'] * 1e3:8.2f} us   frac {j['roofline']['frac']:.3f}", flush=True)
    except Exception as e:
        print(name, "failed", e, r.stderr[-500:])
