mkdir -p gpurun_out/r2k
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2k/tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2k/tests.log
timeout 300 python -m pytest tests/test_t2t_gpu.py -q -s -k front_end 2>&1 | grep -E "rel err|passed|failed"
timeout 600 python bench.py --config t2t_s1 --steps 10 --warmup 3 --no-cpu-baseline --no-live-peaks > gpurun_out/r2k/bench_t2t.json 2> gpurun_out/r2k/bench_t2t.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2k/bench_t2t.json; tail -3 gpurun_out/r2k/bench_t2t.err
