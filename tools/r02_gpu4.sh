#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_rdb_kernel --launch-skip 2 -c 2 -f -o gpurun_out/r02_ncu_rdb_v1 python tools/rdb_probe.py 259 64 2 > gpurun_out/r02_ncu_rdb_v1.log 2>&1
tail -3 gpurun_out/r02_ncu_rdb_v1.log
ls -la gpurun_out/*.ncu-rep
