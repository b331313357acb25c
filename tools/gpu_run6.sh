mkdir -p gpurun_out/r2f
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2f/all_tests.log 2>&1; echo "all tests rc=$?"; tail -4 gpurun_out/r2f/all_tests.log
for pdl in 0 1; do
  UVC_PDL=$pdl timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f/bench_pdl$pdl.json 2> gpurun_out/r2f/bench_pdl$pdl.err; echo "bench pdl=$pdl rc=$?"; cut -c1-330 gpurun_out/r2f/bench_pdl$pdl.json
done
for c in tiny_s1 base_s2 t2t_s1; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2f/bench_$c.json 2> gpurun_out/r2f/bench_$c.err; echo "bench $c rc=$?"; cut -c1-330 gpurun_out/r2f/bench_$c.json; tail -3 gpurun_out/r2f/bench_$c.err
done
