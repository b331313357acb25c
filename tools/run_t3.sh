mkdir -p gpurun_out/r2t3
timeout 600 python -m pytest tests/test_compact_train_gpu.py tests/test_stage2_gpu.py tests/test_cli_gpu.py -q -x 2>&1 | tail -2
for c in base_s2 small_eval base_s2_eval; do
timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline --no-live-peaks > gpurun_out/r2t3/bench_$c.json 2> gpurun_out/r2t3/bench_$c.err; echo "bench $c rc=$?"; cut -c1-220 gpurun_out/r2t3/bench_$c.json
done
UVC_STAGE2=dense timeout 600 python bench.py --config base_s2 --steps 20 --warmup 5 --no-cpu-baseline --no-live-peaks > gpurun_out/r2t3/bench_base_s2_dense.json 2> gpurun_out/r2t3/e1.err; cut -c1-220 gpurun_out/r2t3/bench_base_s2_dense.json
UVC_STAGE2=dense timeout 600 python bench.py --config base_s2_eval --steps 20 --warmup 5 --no-cpu-baseline --no-live-peaks > gpurun_out/r2t3/bench_base_s2_eval_dense.json 2> gpurun_out/r2t3/e2.err; cut -c1-220 gpurun_out/r2t3/bench_base_s2_eval_dense.json
