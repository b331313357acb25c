mkdir -p gpurun_out/r2v
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2v/tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2v/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2v/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2v/bench_small_s1.json 2> gpurun_out/r2v/bench_small_s1.err; echo "bench small rc=$?"; cut -c1-200 gpurun_out/r2v/bench_small_s1.json
for c in tiny_s1 base_s2 t2t_s1; do
timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2v/bench_$c.json 2> gpurun_out/r2v/bench_$c.err; echo "bench $c rc=$?"; cut -c1-200 gpurun_out/r2v/bench_$c.json
done
for c in small_eval base_s2_eval; do
timeout 600 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r2v/bench_$c.json 2> gpurun_out/r2v/bench_$c.err; echo "bench $c rc=$?"; cut -c1-200 gpurun_out/r2v/bench_$c.json
done
UVC_STAGE2=dense timeout 600 python bench.py --config base_s2 --steps 20 --warmup 5 --no-cpu-baseline --no-live-peaks > gpurun_out/r2v/bench_base_s2_dense.json 2> gpurun_out/r2v/e1.err; cut -c1-200 gpurun_out/r2v/bench_base_s2_dense.json
UVC_STAGE2=dense timeout 600 python bench.py --config base_s2_eval --steps 20 --warmup 5 --no-cpu-baseline --no-live-peaks > gpurun_out/r2v/bench_base_s2_eval_dense.json 2> gpurun_out/r2v/e2.err; cut -c1-200 gpurun_out/r2v/bench_base_s2_eval_dense.json
CASES=ln,fc1 timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:'layernorm|gemm2' -o /tmp/kern_c python tests/bringup/kern_cases.py > gpurun_out/r2v/ncu_c.log 2>&1; echo "ncu c rc=$?"
ncu -i /tmp/kern_c.ncu-rep --page raw --csv > gpurun_out/r2v/kern_c_raw.csv 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2330 --launch-count 800 --csv --log-file gpurun_out/r2v/launches_small_s1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-live-peaks > gpurun_out/r2v/ncu_bench.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm2 --launch-skip 480 --launch-count 288 --csv --log-file gpurun_out/r2v/gemm2_traffic_small_s1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-live-peaks > gpurun_out/r2v/ncu2.log 2>&1; echo "ncu2 rc=$?"
