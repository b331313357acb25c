mkdir -p gpurun_out/r2l
timeout 900 python -m pytest tests/test_compact_train_gpu.py tests/test_t2t_gpu.py tests/test_cli_gpu.py -q -x > gpurun_out/r2l/tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2l/tests.log
for c in small_eval base_s2_eval; do
timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-live-peaks > gpurun_out/r2l/bench_$c.json 2> gpurun_out/r2l/bench_$c.err; echo "bench $c rc=$?"; cut -c1-250 gpurun_out/r2l/bench_$c.json; tail -3 gpurun_out/r2l/bench_$c.err
done
UVC_STAGE2=dense timeout 600 python bench.py --config base_s2_eval --steps 20 --warmup 5 --no-live-peaks --no-cpu-baseline > gpurun_out/r2l/bench_base_s2_eval_dense.json 2> gpurun_out/r2l/bench_dense.err; echo "bench dense rc=$?"; cut -c1-250 gpurun_out/r2l/bench_base_s2_eval_dense.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2l/launches_t2t.csv python bench.py --config t2t_s1 --steps 1 --warmup 3 --no-cpu-baseline --no-live-peaks > gpurun_out/r2l/ncu_t2t.log 2>&1; echo "ncu rc=$?"
