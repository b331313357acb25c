mkdir -p gpurun_out/r2m
timeout 600 python -m pytest tests/test_t2t_gpu.py tests/test_compact_train_gpu.py -q -x > gpurun_out/r2m/tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2m/tests.log
timeout 300 python tests/bringup/t2t_frontend_prof.py > gpurun_out/r2m/t2t_front.log 2>&1; grep -E "front end|performer|unfold|gemm|layernorm|dropout|cvt|Self CUDA time" gpurun_out/r2m/t2t_front.log | cut -c1-200 | head -40
timeout 600 python bench.py --config t2t_s1 --steps 10 --warmup 3 --no-cpu-baseline --no-live-peaks > gpurun_out/r2m/bench_t2t.json 2> gpurun_out/r2m/bench_t2t.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r2m/bench_t2t.json
