"""Small end-to-end cases for `compute-sanitizer` (memcheck / racecheck): one DeiT-Tiny Stage-1-like step at a batch large enough for the streamed
LayerNorm backward (1576 rows), one compacted Stage-2 step, one T2T front-end forward + backward with dropout."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import fixtures as fx
from test_model_gpu import build
from test_compact import pruned_checkpoint
from test_compact_train_gpu import _stage2_model, _run
from uvc_b200 import compact as cp
from uvc_b200.models.model_distilled import _VitFunction, _engine_param_list

which = set((os.environ.get("CASES") or "gated,compact,t2t").split(","))
if "gated" in which:
    sd, dims = fx.make_state_dict("deit_tiny_patch16_224", 2, seed=8)
    x, _ = fx.make_batch(8, seed=14)
    m = build("deit_tiny_patch16_224", 2, sd).train()
    blend = torch.tensor([[0.3, 0.7], [0.55, 0.45]]).cuda().requires_grad_(True)
    logits = _VitFunction.apply(m, x.cuda(), blend, None, None, None, *[p for _, p in _engine_param_list(m)])
    logits.square().mean().backward()
    with torch.no_grad():
        m.eval()(x.cuda())
    print("gated step ok", float(blend.grad.abs().sum()))
if "compact" in which:
    sd, dims = pruned_checkpoint("deit_tiny_patch16_224", 4, seed=3)
    m = _stage2_model(sd, "deit_tiny_patch16_224", 4)
    m.compact_layout = cp.engine_layout_for(m)
    x, _ = fx.make_batch(6, seed=11)
    logits, grads = _run(m, x, fx.soft_targets(6, seed=11))
    print("compact step ok", float(logits.abs().sum()))
if "t2t" in which:
    from uvc_b200.T2TViT.models import T2T_module
    torch.manual_seed(0)
    t = T2T_module(embed_dim=384).cuda().train()
    x = torch.randn(2, 3, 224, 224, device="cuda")
    tok, _ = t(x)
    (tok * torch.randn_like(tok) * 0.01).sum().backward()
    print("t2t step ok", float(tok.abs().sum()))
torch.cuda.synchronize()
