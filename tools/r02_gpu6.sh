#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_rdb_v5.log
: > $L
timeout 200 python -m pytest tests/test_gpu_forward.py -x -q -k "fused_dense_block" 2>&1 | tail -15 >> $L
timeout 60 python tools/rdb_probe.py 259 64 5 >> $L 2>&1 || echo "FAILED rc=$?" >> $L
timeout 60 python tools/rdb_probe.py 3 64 5 >> $L 2>&1
XMM_RDB_PREFETCH_ROWS=3 timeout 60 python tools/rdb_probe.py 259 64 5 >> $L 2>&1
XMM_RDB_BACKOFF_NS=32 timeout 60 python tools/rdb_probe.py 259 64 5 >> $L 2>&1
timeout 60 python tools/rdb_probe.py 259 16 5 >> $L 2>&1
timeout 60 python tools/rdb_probe.py 259 1 20 >> $L 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv3x3_rdb_kernel --launch-skip 2 -c 2 -f -o gpurun_out/r02_ncu_rdb_v5 python tools/rdb_probe.py 259 64 2 > gpurun_out/r02_ncu_rdb_v5.log 2>&1
cat $L
