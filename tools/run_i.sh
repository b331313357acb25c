mkdir -p gpurun_out/r2i
timeout 600 python -m pytest tests/test_compact_train_gpu.py tests/test_abi.py -x -q > gpurun_out/r2i/tests.log 2>&1; echo "tests rc=$?"; tail -30 gpurun_out/r2i/tests.log
for m in dense compact exact; do
UVC_STAGE2=$m timeout 600 python bench.py --config base_s2 --steps 10 --warmup 3 --no-cpu-baseline --no-live-peaks > gpurun_out/r2i/bench_base_s2_$m.json 2> gpurun_out/r2i/bench_$m.err; echo "bench $m rc=$?"; cut -c1-260 gpurun_out/r2i/bench_base_s2_$m.json; tail -3 gpurun_out/r2i/bench_$m.err
done
