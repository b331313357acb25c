mkdir -p gpurun_out/r2g
timeout 300 python tests/bringup/gemm16_perf.py > gpurun_out/r2g/gemm16.log 2>&1; echo "perf rc=$?"; cat gpurun_out/r2g/gemm16.log
timeout 900 python -m pytest tests/test_ops_f16_gpu.py tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_input_pipeline_gpu.py -q > gpurun_out/r2g/tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2g/tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2g/bench.json 2> gpurun_out/r2g/bench.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/r2g/bench.json
