#!/bin/bash
# one GPU visit that produces the ncu evidence of a round: launch list of the bench command, then one
# --set full capture each of K3 (config 3), K1 (config 2 shape) and the ingest kernels.
set -x
python bench.py --steps 2 --warmup 3 --events 2.5e8 > gpurun_out/prof_bench.json 2> gpurun_out/prof_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1c.csv \
    python bench.py --steps 2 --warmup 3 --events 2.5e8 > gpurun_out/ncu_bench.log 2>&1
python tools/cosine_perf.py --reps 1 > gpurun_out/prof_cp.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_cosine -c 1 -f -o gpurun_out/prof_cosine_r1c \
    python tools/cosine_perf.py --reps 1 > gpurun_out/ncu_cos.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_update_single -s 2 -c 1 -f -o gpurun_out/prof_update_r1c \
    python bench.py --steps 1 --warmup 3 --events 2.5e8 --no-cosine > gpurun_out/ncu_upd.log 2>&1
python tools/ingest_perf.py --events 4e6 > gpurun_out/prof_ingest.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'k_parse_lines|k_prep_insert|k_count_lines' -c 3 -f -o gpurun_out/prof_ingest_r1c \
    python tools/ingest_perf.py --events 4e6 > gpurun_out/ncu_ingest.log 2>&1
ls -la gpurun_out/*.ncu-rep
