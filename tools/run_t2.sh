timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
UVC_STEP_TIMING=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-live-peaks 2>&1 >/dev/null | grep -v "Initial FLOP" | tail -15
python - <<'PY'
import sys, time, torch
sys.path.insert(0, ".")
import bench, types
cfg = bench.BENCH_CONFIGS["small_s1"]
step, model, info = bench.build_gpu_step(cfg, torch.device("cuda", 0), 1)
x = torch.randn(128, 3, 224, 224, device="cuda"); y = torch.randint(0, 1000, (128,), device="cuda")
for _ in range(5): step(x.clone(), y)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): step(x.clone(), y)
t1 = time.perf_counter()          # CPU enqueue time (GPU runs behind)
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0)/20:.2f} ms per step; wall incl. drain {1e3*(t2-t0)/20:.2f} ms per step", file=sys.stderr)
PY
