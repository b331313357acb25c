mkdir -p gpurun_out/measure
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/measure/tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/measure/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/measure/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/measure/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/measure/bench_small_s1.json 2> gpurun_out/measure/bench_small_s1.err; echo "bench small rc=$?"; cut -c1-200 gpurun_out/measure/bench_small_s1.json
for c in tiny_s1 base_s2 t2t_s1; do
timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/measure/bench_$c.json 2> gpurun_out/measure/bench_$c.err; echo "bench $c rc=$?"; cut -c1-200 gpurun_out/measure/bench_$c.json
done
for c in small_eval base_s2_eval; do
timeout 600 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/measure/bench_$c.json 2> gpurun_out/measure/bench_$c.err; echo "bench $c rc=$?"; cut -c1-200 gpurun_out/measure/bench_$c.json
done
UVC_STAGE2=dense timeout 600 python bench.py --config base_s2 --steps 20 --warmup 5 --no-cpu-baseline --no-live-peaks > gpurun_out/measure/bench_base_s2_dense.json 2> gpurun_out/measure/e1.err; cut -c1-200 gpurun_out/measure/bench_base_s2_dense.json
UVC_STAGE2=dense timeout 600 python bench.py --config base_s2_eval --steps 20 --warmup 5 --no-cpu-baseline --no-live-peaks > gpurun_out/measure/bench_base_s2_eval_dense.json 2> gpurun_out/measure/e2.err; cut -c1-200 gpurun_out/measure/bench_base_s2_eval_dense.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/measure/bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/measure/bench_reference.json
