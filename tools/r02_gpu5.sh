#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_rdb_prof_v5.log
: > $L
XMM_RDB_PROF=1 timeout 60 python tools/rdb_probe.py 259 64 1 >> $L 2>&1
timeout 200 python -m pytest tests/test_gpu_bench_dispatch.py -x -q -k "loss_and_gradient" -s 2>&1 | grep -E "rel-L2|passed|failed" >> $L
cat $L
