mkdir -p gpurun_out/r2c
timeout 600 python -m pytest tests/test_stage2_gpu.py tests/test_ops_f16_gpu.py -q -x > gpurun_out/r2c/tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2c/tests.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 1300 --csv --log-file gpurun_out/r2c/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2c/ncu_bench.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c/bench_f16.json 2> gpurun_out/r2c/bench_f16.err; echo "bench rc=$?"; cat gpurun_out/r2c/bench_f16.json
