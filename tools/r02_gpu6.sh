#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_rdb_v8.log
: > $L
timeout 200 python -m pytest tests/test_gpu_forward.py -x -q -k "fused_dense_block" 2>&1 | tail -3 >> $L
timeout 60 python tools/rdb_probe.py 259 64 5 >> $L 2>&1 || echo "FAILED rc=$?" >> $L
XMM_RDB_SPLIT_PRODUCERS=0 timeout 60 python tools/rdb_probe.py 259 64 5 >> $L 2>&1 || echo "FAILED rc=$?" >> $L
timeout 60 python tools/rdb_probe.py 259 64 5 >> $L 2>&1 || echo "FAILED rc=$?" >> $L
timeout 60 python tools/rdb_probe.py 259 16 5 >> $L 2>&1
timeout 100 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:conv3x3_rdb_kernel --launch-skip 2 -c 2 --csv --log-file gpurun_out/r02_rdb_v8_times.csv python tools/rdb_probe.py 259 64 3 > /dev/null 2>&1
grep -E "rdb_kernel" gpurun_out/r02_rdb_v8_times.csv | awk -F'","' '{print $5, $13, $15}' >> $L
cat $L
