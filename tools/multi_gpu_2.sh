mkdir -p gpurun_out/r2q
timeout 600 python -m pytest tests/test_ddp_nccl_gpu.py -q -x > gpurun_out/r2q/tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2q/tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-live-peaks > gpurun_out/r2q/bench_2gpu.json 2> gpurun_out/r2q/bench_2gpu.err; echo "bench rc=$?"; cut -c1-220 gpurun_out/r2q/bench_2gpu.json; tail -3 gpurun_out/r2q/bench_2gpu.err
