#!/bin/bash
# 1 GPU: launch list (ncu, durations only) of one configs[3]-shaped shard (125000 rows x 125000 columns, depth 1 and 4)
O=gpurun_out/r2t
mkdir -p $O
t0=$(date +%s)
ARGS="--gpus 1 --steps 1 --warmup 3 --events 1e8 --e2e-events 4194304 --no-cosine --big on --c4-items 125000 --c4-events 2.5e8 --c4-check-rows 256 --c4-check-rows-d4 32 --c4-d4 1 --c5-events 0 --c5-items 0"
timeout 300 python bench.py $ARGS > $O/bench_n1_c4shard.json 2> $O/bench_n1_c4shard.err
echo "bench rc=$? t=$(( $(date +%s) - t0 ))" | tee -a $O/summary.txt
grep bench_big $O/bench_n1_c4shard.err | tail
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_c4shard.csv python bench.py $ARGS > $O/ncu_run.json 2> $O/ncu_run.err
echo "ncu rc=$? t=$(( $(date +%s) - t0 ))" | tee -a $O/summary.txt
wc -l $O/launches_c4shard.csv
