import sys; sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from oracle import fixtures as fx, vit_oracle as vo
import test_stage2_gpu as T
from uvc_b200 import ops
mt, depth, B = "deit_tiny_patch16_224", 3, 4
sd, dims = fx.make_state_dict(mt, depth, seed=21)
H = dims["num_heads"]
m = T.build(mt, depth, sd, gumbel_hard=True).train()
x, _ = fx.make_batch(B, seed=5); tgt = fx.soft_targets(B, seed=5); t_logits = torch.zeros(B, 1000)
sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
lo = vo.forward(sdr, x, depth, H, skip=[False]*3)
loss_o, _, _ = vo.distillation_loss(lo, t_logits, tgt, 0.1, 1.0); loss_o.backward()
m.enable_block_gating = 0
(logits, _), macs = m(x.cuda())
parts, dl = ops.distill_loss(logits.detach(), t_logits.cuda(), tgt.cuda(), 0.1, 1.0)
logits.backward(dl)
for k in ["pos_embed", "cls_token", "patch_embed.proj.weight", "blocks.0.attn.qkv.weight", "head.bias"]:
    g = dict(m.named_parameters())[k].grad.cpu(); r = sdr[k].grad
    print(k, "grad max", float(g.abs().max()), float(r.abs().max()), "rel", float((g - r).abs().max() / r.abs().max()), "median |g|", float(r.abs().median()))
print(m.no_weight_decay() if hasattr(m, "no_weight_decay") else None)
