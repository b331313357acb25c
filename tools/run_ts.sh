mkdir -p gpurun_out/r2ts
for v in 1 0 1 0; do
UVC_TEACHER_STREAM=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-live-peaks 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('teacher_stream=$v small_s1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'loss', d['e2e']['last_loss'])"
done
for v in 1 0; do
UVC_TEACHER_STREAM=$v timeout 600 python bench.py --config base_s2 --steps 20 --warmup 5 --no-cpu-baseline --no-live-peaks 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('teacher_stream=$v base_s2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
UVC_TEACHER_STREAM=$v timeout 600 python bench.py --config t2t_s1 --steps 20 --warmup 5 --no-cpu-baseline --no-live-peaks 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('teacher_stream=$v t2t_s1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
done
