mkdir -p gpurun_out/r2d
timeout 300 python tests/bringup/gemm16_perf.py > gpurun_out/r2d/gemm16_ew12.log 2>&1; echo "perf rc=$?"; cat gpurun_out/r2d/gemm16_ew12.log
UVC_LIB_PATH=$PWD/uvc_b200/libuvc_sm100_ew8.so timeout 300 python tests/bringup/gemm16_perf.py > gpurun_out/r2d/gemm16_ew8.log 2>&1; echo "perf ew8 rc=$?"; grep -E "fc1 fwd|fc2 dgrad|sum" gpurun_out/r2d/gemm16_ew8.log
timeout 900 python -m pytest tests/test_token_gate_gpu.py tests/test_admm_gpu.py -q > gpurun_out/r2d/new_tests.log 2>&1; echo "new tests rc=$?"; tail -15 gpurun_out/r2d/new_tests.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2d/all_tests.log 2>&1; echo "all tests rc=$?"; tail -5 gpurun_out/r2d/all_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2d/bench.json 2> gpurun_out/r2d/bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2d/bench.json
