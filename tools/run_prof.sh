python - <<'PY' 2>&1 | tail -12
import sys, time, torch
sys.path.insert(0, ".")
import bench
for name in ("small_s1", "small_eval"):
    cfg = bench.BENCH_CONFIGS[name]
    step, model, info = bench.build_gpu_step(cfg, torch.device("cuda", 0), 1)
    B = cfg["batch"]
    x = torch.randn(B, 3, 224, 224, device="cuda"); y = torch.randint(0, 1000, (B,), device="cuda")
    for _ in range(5): step(x.clone(), y)
    ts = []
    for _ in range(10):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); step(x.clone(), y); t1 = time.perf_counter()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        ts.append((t1 - t0, t2 - t0))
    ts.sort()
    print(f"{name}: host enqueue of one step from an idle GPU: median {1e3*ts[5][0]:.2f} ms (min {1e3*ts[0][0]:.2f}); step wall from idle {1e3*sorted(t[1] for t in ts)[5]:.2f} ms", file=sys.stderr)
PY
