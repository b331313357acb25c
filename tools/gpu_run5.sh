mkdir -p gpurun_out/r2e
timeout 900 python -m pytest tests/test_token_gate_gpu.py tests/test_stage2_gpu.py -q > gpurun_out/r2e/tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2e/tests.log
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2e/kern_full python tests/bringup/kern_cases.py > gpurun_out/r2e/ncu_full.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2e/ncu_full.log
