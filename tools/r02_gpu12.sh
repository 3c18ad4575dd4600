#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_bench_dispatch.py -q -k "loss_and_gradient" -s --timeout=280 --timeout-method=thread 2>&1 | grep -E "rel-L2|passed|failed|^E " | head -20
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 400 -c 700 --csv --log-file gpurun_out/r02_ncu_launches_train_dn.csv python bench.py --workload train_dn --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train.log 2>&1
wc -l gpurun_out/r02_ncu_launches_train_dn.csv
