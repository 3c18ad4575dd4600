#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_rdb_prof.log
: > $L
XMM_RDB_PROF=1 timeout 60 python tools/rdb_probe.py 259 64 1 >> $L 2>&1
cat $L
