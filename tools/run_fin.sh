mkdir -p gpurun_out/r2fin
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2fin/tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2fin/tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2fin/bench_small_s1.json 2> gpurun_out/r2fin/bench_small_s1.err; echo "bench small rc=$?"; cut -c1-200 gpurun_out/r2fin/bench_small_s1.json
for c in tiny_s1 base_s2 t2t_s1; do
timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2fin/bench_$c.json 2> gpurun_out/r2fin/bench_$c.err; echo "bench $c rc=$?"; cut -c1-200 gpurun_out/r2fin/bench_$c.json
done
for c in small_eval base_s2_eval; do
timeout 600 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r2fin/bench_$c.json 2> gpurun_out/r2fin/bench_$c.err; echo "bench $c rc=$?"; cut -c1-200 gpurun_out/r2fin/bench_$c.json
done
UVC_STAGE2=dense timeout 600 python bench.py --config base_s2 --steps 20 --warmup 5 --no-cpu-baseline --no-live-peaks > gpurun_out/r2fin/bench_base_s2_dense.json 2> gpurun_out/r2fin/e1.err; cut -c1-200 gpurun_out/r2fin/bench_base_s2_dense.json
UVC_STAGE2=dense timeout 600 python bench.py --config base_s2_eval --steps 20 --warmup 5 --no-cpu-baseline --no-live-peaks > gpurun_out/r2fin/bench_base_s2_eval_dense.json 2> gpurun_out/r2fin/e2.err; cut -c1-200 gpurun_out/r2fin/bench_base_s2_eval_dense.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2330 --launch-count 800 --csv --log-file gpurun_out/r2fin/launches_small_s1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-live-peaks > gpurun_out/r2fin/ncu_bench.log 2>&1; echo "ncu rc=$?"
