mkdir -p gpurun_out/r2y
timeout 900 ncu --set full --clock-control none -k regex:'colnorm|rank_kernel|prox_kernel|admm_|mixup|im2col|assemble|cvt_f16|pos_cls|distill|grad_scale' --launch-skip 60 --launch-count 22 -o /tmp/kern_d python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-live-peaks > gpurun_out/r2y/ncu_d.log 2>&1; echo "ncu rc=$?"; grep -c Profiling gpurun_out/r2y/ncu_d.log
ncu -i /tmp/kern_d.ncu-rep --page raw --csv > gpurun_out/r2y/kern_d_raw.csv 2>/dev/null; ls -la gpurun_out/r2y
