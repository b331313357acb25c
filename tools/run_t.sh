mkdir -p gpurun_out/r2t
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2t/tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2t/tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-live-peaks > gpurun_out/r2t/bench_small_s1.json 2> gpurun_out/r2t/bench_small_s1.err; echo "bench small rc=$?"; cut -c1-200 gpurun_out/r2t/bench_small_s1.json
for c in t2t_s1 small_eval base_s2_eval base_s2; do
timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline --no-live-peaks > gpurun_out/r2t/bench_$c.json 2> gpurun_out/r2t/bench_$c.err; echo "bench $c rc=$?"; cut -c1-200 gpurun_out/r2t/bench_$c.json
done
