#!/bin/bash
# round-2 single-GPU evidence: GPU test-suite, smoke, headline bench (+ training extras), reference arm, launch lists
# (each step under its own timeout; a wedged kernel must not eat the GPU budget)
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -q -m gpu --timeout=300 --timeout-method=thread -s 2>&1 | grep -v "^$" | tail -150) > gpurun_out/r02_pytest_gpu_final.log 2>&1
grep -E "passed|failed|FAILED|Timeout" gpurun_out/r02_pytest_gpu_final.log | tail -8
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
XMM_BENCH_WATCHDOG=400 timeout 500 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_full_1gpu_final.json 2> gpurun_out/r02_bench_final.err
tail -c 300 gpurun_out/r02_bench_final.err
if [ "$1" = "all" ]; then
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>> gpurun_out/r02_bench_final.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 170 -c 80 --csv --log-file gpurun_out/r02_ncu_launches_infer_b64.csv python bench.py --steps 1 --warmup 3 --no-train-extra --no-cpu-baseline --no-parity > gpurun_out/ncu_infer.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 400 -c 700 --csv --log-file gpurun_out/r02_ncu_launches_train_dn.csv python bench.py --workload train_dn --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train.log 2>&1
fi
python - <<'PY'
import json
for f in ('gpurun_out/r02_bench_full_1gpu_final.json',):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k:d[k] for k in ('value','ms_per_step','parity_rel_l2','gpu_launches') if k in d}, d.get('e2e',{}).get('value'), d.get('cpu_baseline'))
        if 'roofline' in d: print({k:d['roofline'][k] for k in ('achieved','frac','frac_of_burst','share_of_step','launches_timed','whole_step_tflops')})
        print({k:(round(v['value'],1),round(v['ms_per_step'],2),round(v['roofline']['frac_of_burst'],3)) for k,v in d.get('extra',{}).items()})
    except Exception as e: print(f, "no line", e)
PY
