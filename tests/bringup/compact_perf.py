"""Bring-up: inference images/sec of DeiT-Base under the fixed ~50 % Stage-2 layout of BASELINE.json configs[3] (blocks 8 and 10 skipped; in the live
blocks 3 heads, 16 dims of every surviving head and 1417 neurons pruned): whole-model engine on the masked-dense checkpoint vs the compact runner."""
import os, sys
from functools import partial
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import fixtures as fx
from uvc_b200 import compact as cp
from uvc_b200.models.model_distilled import DistilledVisionTransformer

B = int(os.environ.get("B", 128))
sd, dims = fx.make_state_dict("deit_base_patch16_224", 12, seed=1)
H, C_ = dims["num_heads"], dims["embed_dim"]
Fh = 4 * C_
g = torch.Generator().manual_seed(0)
for l in range(12):
    pre = f"blocks.{l}."
    m1, m3, m2 = torch.ones(C_, C_), torch.ones(C_, Fh), torch.ones(Fh, C_)
    m1[:, : 3 * 64] = 0
    for h in range(3, H):
        m1[:, h * 64 + torch.randperm(64, generator=g)[:16]] = 0
    dead = torch.randperm(Fh, generator=g)[:1417]
    m3[:, dead] = 0; m2[dead, :] = 0
    sd[pre + "attn.proj.mask"], sd[pre + "mlp.fc2.mask"], sd[pre + "mlp.fc1.mask"] = m1, m3, m2
sd["block_skip_gating"][8] = torch.tensor([1.0, -1.0]); sd["block_skip_gating"][10] = torch.tensor([1.0, -1.0])
dense = dict(sd)
for k in list(sd):
    if k.endswith(".mask"):
        dense[k[:-4] + "weight"] = sd[k[:-4] + "weight"] * sd[k]
lay = cp.compile_layout(sd, H)
print("MACs ratio compact / dense: %.3f" % cp.macs(lay)["ratio"])
comp = cp.CompactViT(cp.compact_state_dict(sd, lay)).cuda()
m = DistilledVisionTransformer(enable_dist=0, patch_size=16, mlp_ratio=4, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), drop_rate=0,
                               embed_dim=C_, depth=12, num_heads=H)
m.load_state_dict({k: v for k, v in dense.items() if not k.endswith(".mask")}, strict=False)
m = m.cuda().eval()
x = torch.randn(B, 3, 224, 224, device="cuda")


def timeit(fn, n=15):
    for _ in range(4): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    a, _ = m(x); b = comp(x)
    print("logits rel diff compact vs engine: %.2e" % float((a - b).abs().max() / a.abs().max()))
    t_e = timeit(lambda: m(x)); t_c = timeit(lambda: comp(x))
print(f"B={B}: engine masked-dense {t_e:.2f} ms = {B / t_e * 1e3:.0f} images/sec ; compact runner {t_c:.2f} ms = {B / t_c * 1e3:.0f} images/sec ; x{t_e / t_c:.2f}")
