#!/bin/bash
# Multi-GPU measurements of one box (run through `gpurun --gpus N -- bash tools/scale_run.sh N`).
# For n in 1,2,4,..,N: the default workload (replicas), the row-sharded C4 trainer, the item-sharded C5 predict.
N=${1:-2}
O=gpurun_out
cd "$(dirname "$0")/.."
run() {  # n, extra bench args, output tag
  local n=$1; shift; local tag=$1; shift
  if [ "$n" = 1 ]; then python bench.py --gpus 1 "$@" > $O/scale_${tag}_n$n.json 2> $O/scale_${tag}_n$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
       bench.py --gpus $n "$@" > $O/scale_${tag}_n$n.json 2> $O/scale_${tag}_n$n.err; fi
  python - <<PY
import json
try:
    j = json.loads(open("$O/scale_${tag}_n$n.json").read().strip().splitlines()[-1])
    print("$tag", "n=$n", round(j["value"]), j["unit"], "ms/step", round(j["ms_per_step"], 4), "e2e", round(j["e2e"]["value"]))
except Exception as e:
    print("$tag n=$n FAILED", e)
PY
}
n=1
while [ $n -le $N ]; do
  run $n c2 --steps 200 --warmup 20 --no-cpu-baseline
  run $n c4 --workload c4_linear --steps 30 --warmup 5 --no-cpu-baseline
  run $n c5 --workload c5_predict --steps 3 --warmup 3 --no-cpu-baseline
  n=$((n * 2))
done
