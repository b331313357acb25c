"""Bring-up: where the tokens_to_token front end (uvc_t2t_forward / uvc_t2t_backward) spends its time (forward + backward, B = 128, training mode).
Round 1 ran this on the eager-torch front end: 26.7 ms (profiles/r01_t2t_frontend_torch_profile.txt)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from torch.profiler import profile, ProfilerActivity
from uvc_b200.T2TViT.models import T2T_module
torch.manual_seed(0)
m = T2T_module(embed_dim=384).cuda().train()
x = torch.randn(128, 3, 224, 224, device="cuda")
r = torch.randn(128, 196, 384, device="cuda") * 0.01
def step():
    tok, _ = m(x)
    (tok * r).sum().backward()
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); torch.cuda.synchronize()
print("front end fwd+bwd: %.2f ms" % (e0.elapsed_time(e1) / 5))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
