python -m pytest tests -m gpu -x -q 2>&1 | tail -3
build/conv_prof 8 pair 2>&1 | grep "^dx-rr\|mma-loop total" | head -8
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('infer', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz']); e=d['extra']; print('train_dn', round(e['train_dn']['value'],1), 'train_sr', round(e['train_sr']['value'],1))" | tee gpurun_out/bench_hybrid.log
