#!/bin/bash
mkdir -p gpurun_out/r2f
timeout 1200 python -m pytest tests/test_cosine_gpu.py tests/test_itemsimilarity_gpu.py tests/test_sketch_gpu.py -x -q -m gpu > gpurun_out/r2f/pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2f/summary.txt
tail -15 gpurun_out/r2f/pytest.log
for prec in tensor certified rescored; do
  timeout 300 python tools/cosine_perf.py --precision $prec --reps 4 >> gpurun_out/r2f/cosine_perf.jsonl 2>> gpurun_out/r2f/cosine_perf.err
done
cat gpurun_out/r2f/cosine_perf.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2f/launches_certified.csv \
  python tools/cosine_perf.py --precision certified --reps 2 > gpurun_out/r2f/ncu_cert.log 2>&1
echo "ncu rc=$?" | tee -a gpurun_out/r2f/summary.txt
