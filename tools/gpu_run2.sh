mkdir -p gpurun_out/r2b
timeout 300 python -m pytest tests/test_ops_f16_gpu.py -q -k gemm > gpurun_out/r2b/gemm.log 2>&1; echo "gemm rc=$?"; tail -3 gpurun_out/r2b/gemm.log
for f in test_model_gpu test_stage2_gpu test_t2t_gpu; do
  timeout 900 python -m pytest tests/$f.py -q -s > gpurun_out/r2b/$f.log 2>&1; echo "$f rc=$?"; grep -E "rel err|passed|failed|FAILED|Error" gpurun_out/r2b/$f.log | tail -40
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b/bench_f16.json 2> gpurun_out/r2b/bench_f16.err; echo "bench rc=$?"; cat gpurun_out/r2b/bench_f16.json; tail -5 gpurun_out/r2b/bench_f16.err
UVC_PRECISION=tf32 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b/bench_tf32.json 2> gpurun_out/r2b/bench_tf32.err; echo "bench tf32 rc=$?"; cat gpurun_out/r2b/bench_tf32.json
