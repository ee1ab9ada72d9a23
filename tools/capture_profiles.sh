#!/bin/bash
# Round profile capture (run on the GPU box through gpurun): plain runs first, then ncu on the same command lines.
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
python bench.py --steps 200 --warmup 20 > $O/bench_r1_c2.json 2> $O/bench_r1_c2.err
python bench.py --steps 60 --warmup 3 --no-cpu-baseline > $O/plain_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c2.csv \
    python bench.py --steps 60 --warmup 3 --no-cpu-baseline > $O/ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:train_kernel -c 1 -f -o $O/prof_train_r1 \
    python bench.py --steps 60 --warmup 3 --no-cpu-baseline > $O/ncu_c2_full.log 2>&1
python bench.py --workload c3_mlp --steps 10 --warmup 3 --no-cpu-baseline > $O/plain_c3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv --log-file $O/launches_c3.csv \
    python bench.py --workload c3_mlp --steps 10 --warmup 3 --no-cpu-baseline > $O/ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tn -s 30 -c 3 -f -o $O/prof_gemm_r1 \
    python bench.py --workload c3_mlp --steps 10 --warmup 3 --no-cpu-baseline > $O/ncu_c3_full.log 2>&1
python bench.py --workload c5_predict --steps 1 --warmup 3 --no-cpu-baseline > $O/plain_c5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:topk_score -s 3 -c 1 -f -o $O/prof_topk_r1 \
    python bench.py --workload c5_predict --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_c5_full.log 2>&1
python bench.py --workload c3_mlp --steps 50 --warmup 5 > $O/bench_r1_c3.json 2> $O/bench_r1_c3.err
python bench.py --workload c5_predict --steps 5 --warmup 3 > $O/bench_r1_c5.json 2> $O/bench_r1_c5.err
tail -c 400 $O/bench_r1_c2.json
