mkdir -p gpurun_out/r2r
timeout 900 python -m pytest tests/test_ops_f16_gpu.py tests/test_model_gpu.py tests/test_stage2_gpu.py -q -x > gpurun_out/r2r/tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2r/tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-live-peaks > gpurun_out/r2r/bench_stream.json 2> gpurun_out/r2r/bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2r/bench_stream.json
UVC_LN_BWD_REG=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-live-peaks > gpurun_out/r2r/bench_reg.json 2> gpurun_out/r2r/bench2.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2r/bench_reg.json
timeout 300 python tests/bringup/ln16_perf.py > gpurun_out/r2r/ln16.log 2>&1; cat gpurun_out/r2r/ln16.log
UVC_LN_BWD_REG=1 timeout 300 python tests/bringup/ln16_perf.py > gpurun_out/r2r/ln16_reg.log 2>&1; cat gpurun_out/r2r/ln16_reg.log
