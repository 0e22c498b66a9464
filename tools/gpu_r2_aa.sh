#!/bin/bash
O=gpurun_out/r2aa
mkdir -p $O
MB200_TRACE=1 timeout 300 python tools/k3_k_sweep.py --items 500000 --events 1e9 --depth 1 --ks 100 --chunks 0,4,8,16 --reps 1 > $O/k3_chunks_500k.jsonl 2> $O/err1.log; echo rc=$?
cat $O/k3_chunks_500k.jsonl; grep "plan:" $O/err1.log | sort | uniq -c; tail -2 $O/err1.log
