#!/bin/bash
mkdir -p gpurun_out/r2l
timeout 1200 python -m pytest tests/test_cosine_gpu.py tests/test_itemsimilarity_gpu.py tests/test_sketch_gpu.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2l/pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/pipelined_debug2.py 200000 4e8 2>&1 | grep -v "OMP_NUM\|^\*\*\*\|^$" | tee gpurun_out/r2l/pipelined_debug2.log
timeout 900 python tools/stress.py 2>&1 | tee gpurun_out/r2l/stress.jsonl | cut -c1-330
