#!/bin/bash
# round-2 GPU session C: partition kernels with L2 prefetch / 4 CTAs per SM; narrow wire format; RED peak
mkdir -p gpurun_out/r2c
timeout 900 python -m pytest tests/test_sketch_gpu.py tests/test_ingest_gpu.py -x -q -m gpu > gpurun_out/r2c/pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2c/summary.txt
tail -3 gpurun_out/r2c/pytest.log
for cfg in "26744 2e7 1 grouped" "125000 2.5e8 1 grouped" "125000 1e9 1 grouped" "125000 1e9 1 csr"; do
  set -- $cfg
  timeout 600 python tools/k1_bank_bench.py --items $1 --events $2 --calls $3 --mode $4 --parity-events 0 >> gpurun_out/r2c/k1_bank.jsonl 2>> gpurun_out/r2c/k1_bank.err
  echo "k1 $cfg rc=$?" | tee -a gpurun_out/r2c/summary.txt
done
cat gpurun_out/r2c/k1_bank.jsonl
python - <<'PY' 2>&1 | tee gpurun_out/r2c/red_peak.txt
import mahout_b200 as mb
from mahout_b200 import synth
ctx = mb.Context(0)
for lg in (22, 24, 28):
    print("RED.ADD.64 peak, 2^%d cells: %.1f G/s" % (lg, synth.red64_peak(ctx, lg) / 1e9))
PY
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_group|k_update_grouped|k_scan" -c 16 --csv \
  --log-file gpurun_out/r2c/launches_k1_grouped.csv python tools/k1_bank_bench.py --items 125000 --events 2.5e8 --mode grouped --reps 1 --parity-events 0 \
  > gpurun_out/r2c/ncu_k1.log 2>&1
echo "ncu rc=$?" | tee -a gpurun_out/r2c/summary.txt
