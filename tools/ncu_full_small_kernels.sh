mkdir -p gpurun_out/r2z
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2z/tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2z/tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-live-peaks > gpurun_out/r2z/bench_small_s1.json 2> gpurun_out/r2z/bench_small_s1.err; echo "bench small rc=$?"; cut -c1-200 gpurun_out/r2z/bench_small_s1.json
timeout 900 ncu --set full --clock-control none -k regex:'colnorm|rank_kernel|prox_kernel|admm_|mixup|im2col|assemble|cvt_f16|pos_cls|distill|grad_scale' --launch-skip 60 --launch-count 22 -o /tmp/kern_e python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-live-peaks > gpurun_out/r2z/ncu_e.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/kern_e.ncu-rep --page raw --csv > gpurun_out/r2z/kern_e_raw.csv 2>/dev/null
