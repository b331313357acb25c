mkdir -p gpurun_out/r2a
for k in "gemm" "layernorm or cvt" "f16_forward" "f16_backward" "nan"; do
  name=$(echo $k | tr ' ' '_')
  timeout 300 python -m pytest tests/test_ops_f16_gpu.py -q -k "$k" > gpurun_out/r2a/$name.log 2>&1
  echo "$k rc=$?"; tail -5 gpurun_out/r2a/$name.log
done
timeout 600 python -m pytest tests/test_ops_gpu.py -q -x > gpurun_out/r2a/ops_old.log 2>&1; echo "old ops rc=$?"; tail -3 gpurun_out/r2a/ops_old.log
