#!/bin/bash
# round-2 GPU session A: grouped K1 parity + stage timings + launch list
mkdir -p gpurun_out/r2a
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_sketch_gpu.py -x -q -m gpu > gpurun_out/r2a/pytest_sketch.log 2>&1
echo "pytest_sketch rc=$?" | tee -a gpurun_out/r2a/summary.txt
tail -5 gpurun_out/r2a/pytest_sketch.log
for cfg in "26744 2e7 1" "125000 2.5e8 1" "125000 2.5e8 4"; do
  set -- $cfg
  for mode in direct grouped csr; do
    timeout 600 python tools/k1_bank_bench.py --items $1 --events $2 --calls $3 --mode $mode >> gpurun_out/r2a/k1_bank.jsonl 2>> gpurun_out/r2a/k1_bank.err
    echo "k1 $cfg $mode rc=$?" | tee -a gpurun_out/r2a/summary.txt
  done
done
cat gpurun_out/r2a/k1_bank.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
  --log-file gpurun_out/r2a/launches_k1_grouped.csv python tools/k1_bank_bench.py --items 125000 --events 2.5e8 --mode grouped --reps 1 --parity-events 0 \
  > gpurun_out/r2a/ncu_k1.log 2>&1
echo "ncu rc=$?" | tee -a gpurun_out/r2a/summary.txt
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_sketch_gpu.py > gpurun_out/r2a/pytest_rest.log 2>&1
echo "pytest_rest rc=$?" | tee -a gpurun_out/r2a/summary.txt
tail -5 gpurun_out/r2a/pytest_rest.log
