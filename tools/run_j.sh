mkdir -p gpurun_out/r2j
timeout 600 python -m pytest tests/test_t2t_gpu.py -x -q -s > gpurun_out/r2j/tests.log 2>&1; echo "tests rc=$?"; grep -E "rel err|passed|failed|Error|error" gpurun_out/r2j/tests.log | head -30; tail -15 gpurun_out/r2j/tests.log
timeout 600 python bench.py --config t2t_s1 --steps 10 --warmup 3 --no-cpu-baseline --no-live-peaks > gpurun_out/r2j/bench_t2t.json 2> gpurun_out/r2j/bench_t2t.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2j/bench_t2t.json; tail -5 gpurun_out/r2j/bench_t2t.err
