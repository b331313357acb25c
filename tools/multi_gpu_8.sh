mkdir -p gpurun_out/r2x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-live-peaks > gpurun_out/r2x/bench_8gpu.json 2> gpurun_out/r2x/bench_8gpu.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r2x/bench_8gpu.json; tail -2 gpurun_out/r2x/bench_8gpu.err
