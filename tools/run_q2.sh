for v in 0 1; do
UVC_TEACHER_STREAM=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$v bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-live-peaks 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('teacher_stream=$v', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
done
