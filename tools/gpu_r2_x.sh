#!/bin/bash
O=gpurun_out/r2x
mkdir -p $O
timeout 200 python tools/k3_k_sweep.py --depth 1 --ks 10,50,100,150 > $O/k3_k_sweep.jsonl 2> $O/err1.log; echo rc=$?
timeout 200 python tools/k3_k_sweep.py --depth 4 --ks 10,50,100 >> $O/k3_k_sweep.jsonl 2> $O/err2.log; echo rc=$?
timeout 200 python tools/k3_k_sweep.py --depth 1 --ks 100 --precision certified >> $O/k3_k_sweep.jsonl 2> $O/err3.log; echo rc=$?
cat $O/k3_k_sweep.jsonl; tail -3 $O/err1.log
