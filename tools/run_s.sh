mkdir -p gpurun_out/r2s
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2s/tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2s/tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2s/bench_small_s1.json 2> gpurun_out/r2s/bench_small_s1.err; echo "bench small rc=$?"; cut -c1-200 gpurun_out/r2s/bench_small_s1.json
for c in tiny_s1 base_s2 t2t_s1 small_eval base_s2_eval; do
timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2s/bench_$c.json 2> gpurun_out/r2s/bench_$c.err; echo "bench $c rc=$?"; cut -c1-200 gpurun_out/r2s/bench_$c.json
done
UVC_STAGE2=dense timeout 600 python bench.py --config base_s2 --steps 20 --warmup 5 --no-cpu-baseline --no-live-peaks > gpurun_out/r2s/bench_base_s2_dense.json 2> gpurun_out/r2s/bench_base_s2_dense.err; cut -c1-200 gpurun_out/r2s/bench_base_s2_dense.json
UVC_STAGE2=dense timeout 600 python bench.py --config base_s2_eval --steps 20 --warmup 5 --no-cpu-baseline --no-live-peaks > gpurun_out/r2s/bench_base_s2_eval_dense.json 2> gpurun_out/r2s/bench_base_s2_eval_dense.err; cut -c1-200 gpurun_out/r2s/bench_base_s2_eval_dense.json
