#!/bin/bash
# full GPU test-suite (each test under a hard 300 s limit), then the headline bench (watchdog), then the launch list
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -q -m gpu --timeout=300 --timeout-method=thread -s 2>&1 | grep -v "^$" | tail -120) > gpurun_out/r02_pytest_gpu.log 2>&1
grep -E "passed|failed|FAILED|Timeout|rel-L2|dPSNR|gradient" gpurun_out/r02_pytest_gpu.log | tail -60
XMM_BENCH_WATCHDOG=300 timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
tail -c 400 gpurun_out/r02_bench_a.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_a.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','e2e','parity','gpu_launches','cpu_baseline') if k in d})
    print(d['roofline'])
    print({k:(v['value'],v['ms_per_step']) for k,v in d.get('extra',{}).items()})
except Exception as e: print("no bench line", e)
PY
