mkdir -p gpurun_out/r2z2
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2z2/tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2z2/tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-live-peaks > gpurun_out/r2z2/bench_small_s1.json 2> gpurun_out/r2z2/bench_small_s1.err; echo "bench small rc=$?"; cut -c1-200 gpurun_out/r2z2/bench_small_s1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'rank_kernel|assemble_tokens_kernel' --launch-skip 12 --launch-count 6 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-live-peaks 2>&1 | grep -E "rank_kernel|assemble_tokens_kernel|gpu__time" | head -12
