for p in 1 0 1 0; do XMM_DX_PAIR=$p python bench.py --steps 10 --warmup 3 --no-train-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pair=$p', round(d['value'],1), 'img/s', d['ms_per_step'], d['clocks'], d['roofline']['frac'])"; done 2>&1 | tee gpurun_out/bench_pair_ab.log
