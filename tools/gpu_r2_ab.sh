#!/bin/bash
# 1 GPU: the whole GPU suite and the default bench line with the final code
O=gpurun_out/r2ab
mkdir -p $O
t0=$(date +%s)
timeout 600 python -m pytest tests -q -m gpu --tb=short > $O/pytest.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - t0 ))" | tee -a $O/summary.txt
tail -12 $O/pytest.log
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$? t=$(( $(date +%s) - t0 ))" | tee -a $O/summary.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
