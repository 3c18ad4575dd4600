#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_rdb_v5.log
: > $L
timeout 200 python -m pytest tests/test_gpu_forward.py -x -q -k "fused_dense_block" 2>&1 | tail -3 >> $L
run() { echo "== $*" >> $L; env "$@" timeout 60 python tools/rdb_probe.py 259 64 5 >> $L 2>&1 || echo "FAILED rc=$?" >> $L; }
run A=0
run XMM_RDB_MULTI_ISSUE=0
run XMM_RDB_BACKOFF_NS=32
run XMM_RDB_BACKOFF_NS=100
run XMM_RDB_PREFETCH_ROWS=3
run XMM_RDB_PREFETCH_ROWS=6
run XMM_RDB_PREFETCH_ROWS=4 XMM_RDB_BACKOFF_NS=32
run XMM_RDB_PREFETCH_ROWS=4 XMM_RDB_BACKOFF_NS=32 XMM_RDB_MULTI_ISSUE=0
run XMM_RDB_MAX_CTAS=74
cat $L
