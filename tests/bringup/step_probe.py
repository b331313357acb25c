"""GPU bring-up probe: time student fwd+bwd and teacher fwd of DeiT-Small at B=128 through the engine."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from uvc_b200.models import deit_small_patch16_224, deit_tiny_patch16_224

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
mk = deit_small_patch16_224 if (len(sys.argv) < 3 or sys.argv[2] == "small") else deit_tiny_patch16_224
torch.manual_seed(0)
m = mk(gumbel_hard=False).cuda().train()
t = mk().cuda().eval()
m.enable_block_gating, m.enable_warmup = 1, 1
x = torch.randn(B, 3, 224, 224, device="cuda")


N_IT, W_IT = int(os.environ.get('PROBE_N', 10)), int(os.environ.get('PROBE_W', 3))


def timeit(fn, n=None, w=None):
    n = N_IT if n is None else n; w = W_IT if w is None else w
    for _ in range(w): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n


def fb():
    (o, _), _ = m(x)
    o.backward(torch.ones_like(o) * 1e-3)
    m.zero_grad(set_to_none=True)


def fwd_only():
    with torch.no_grad():
        t(x)


gpu, wall = timeit(fb)
F = 9.198e9 if mk is deit_small_patch16_224 else 2.507e9
print(f"student fwd+bwd  B={B}: {gpu:.2f} ms gpu ({wall:.2f} ms wall)  {3*F*B/gpu/1e9:.1f} TFLOP/s")
gpu2, wall2 = timeit(fwd_only)
print(f"teacher fwd      B={B}: {gpu2:.2f} ms gpu ({wall2:.2f} ms wall)  {F*B/gpu2/1e9:.1f} TFLOP/s")
print(f"=> {B/(gpu+gpu2)*1e3:.0f} img/s (fwd+bwd+teacher only)")
print("mem GB", torch.cuda.max_memory_allocated() / 2**30)
