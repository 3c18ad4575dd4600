#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_rdb_v6.log
: > $L
timeout 200 python -m pytest tests/test_gpu_forward.py -x -q -k "fused_dense_block" 2>&1 | tail -3 >> $L
timeout 60 python tools/rdb_probe.py 259 64 5 >> $L 2>&1 || echo "FAILED rc=$?" >> $L
timeout 60 python tools/rdb_probe.py 3 64 5 >> $L 2>&1
timeout 60 python tools/rdb_probe.py 259 16 5 >> $L 2>&1
timeout 60 python tools/rdb_probe.py 259 1 20 >> $L 2>&1
timeout 200 python -m pytest tests/test_gpu_bench_dispatch.py -x -q -k "loss_and_gradient" -s 2>&1 | grep -E "rel-L2|passed|failed" >> $L
cat $L
