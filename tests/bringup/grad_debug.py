import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import fixtures as fx, vit_oracle as vo
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from test_model_gpu import build
from uvc_b200.models.model_distilled import _VitFunction, _engine_param_list

for depth, use_blend in [(1, False), (2, False), (1, True), (2, True)]:
    sd, dims = fx.make_state_dict("deit_tiny_patch16_224", depth, seed=21)
    x, _ = fx.make_batch(2, seed=731)
    m = build("deit_tiny_patch16_224", depth, sd).train()
    blend = torch.tensor([[0.3, 0.7]] * depth).cuda().requires_grad_(True) if use_blend else None
    params = [p for _, p in _engine_param_list(m)]
    logits = _VitFunction.apply(m, x.cuda(), blend, None, None, None, *params)
    dl = torch.randn_like(logits) * 1e-2
    logits.backward(dl)
    torch.cuda.synchronize()
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    br = blend.detach().cpu().requires_grad_(True) if use_blend else None
    lo = vo.forward(sdr, x, depth, dims["num_heads"], blend=br)
    lo.backward(dl.cpu())
    print(f"--- depth={depth} blend={use_blend} arena absmax={m.flat_grad.abs().max().item():.3e}")
    for k, p in m.named_parameters():
        if p.grad is None: continue
        e = ((p.grad.cpu() - sdr[k].grad).abs().max() / sdr[k].grad.abs().max().clamp_min(1e-30)).item()
        if e > 5e-3 or "cls" in k or "head.w" in k:
            print(f"   {k:32s} rel={e:.3e} got={p.grad.abs().max().item():.3e} ref={sdr[k].grad.abs().max().item():.3e} inarena={p.grad.data_ptr() >= m.flat_grad.data_ptr()}")
    if use_blend:
        print("   d_blend", blend.grad.tolist(), br.grad.tolist())
