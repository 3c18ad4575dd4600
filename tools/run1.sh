for p in 0 1 2 3; do echo "== L2PROMO=$p"; XMM_TMAP_L2PROMO=$p build/probe time 2>&1 | grep "mode=1"; done > gpurun_out/probe3_promo.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1200 -c 1300 --csv --log-file gpurun_out/launches_train_dn.csv python bench.py --workload train_dn --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_train.log 2>&1
echo rc=$?
