mkdir -p gpurun_out/r2h
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2h/tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2h/tests.log
python tests/bringup/ln16_perf.py > gpurun_out/r2h/ln16.log 2>&1; cat gpurun_out/r2h/ln16.log
UVC_LIB_PATH=$PWD/uvc_b200/libuvc_sm100_alt.so python tests/bringup/ln16_perf.py > gpurun_out/r2h/ln16_alt.log 2>&1; cat gpurun_out/r2h/ln16_alt.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2h/bench.json 2> gpurun_out/r2h/bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2h/bench.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2h/ncu_bench.log 2>&1; echo "ncu rc=$?"
