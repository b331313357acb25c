"""Bring-up: images/sec of the Stage-1 joint_train step on T2T-ViT-14 (BASELINE.json configs[4]: budget 0.6, token gate on, 128 images per GPU,
soft distillation from a copy of the initial student), same step object and flags as bench.py's DeiT-Small arm.  One GPU, device-resident inputs."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from uvc_b200.joint_train import Stage1Step, get_uvc_layers, make_optimizer
from uvc_b200.T2TViT.models import t2t_vit_14
from uvc_b200.utils.losses import DistillationLoss
from uvc_b200.utils.mixup import Mixup, SoftTargetCrossEntropy
from uvc_b200.utils.scheduler import WarmupCosineSchedule
from uvc_b200.uvc_optimizer import build_minimax_model
from uvc_b200.uvc_utils import prune_w_mask

device = torch.device("cuda", 0)
B = int(os.environ.get("B", 128))
args = bench.uvc_args_namespace(6, device=device, local_rank=-1, budget=0.6, enable_patch_gating=2)
torch.manual_seed(730); np.random.seed(730)
model = t2t_vit_14(gumbel_hard=False).to(device)
teacher = t2t_vit_14(gumbel_hard=True).to(device).eval()
teacher.load_state_dict(model.state_dict(), strict=False)
for _, m in model.named_modules():
    if hasattr(m, "weight"):
        m.register_buffer("mask", torch.ones_like(m.weight))
layer_names, uvc_layers, uvc_dict = get_uvc_layers(model)
model.eval()
with torch.no_grad():
    _, flops_list = model(torch.ones(1, 3, 224, 224, device=device))
uvc = list(build_minimax_model(model, layer_names, uvc_layers, uvc_dict, args, flops_list))
mm = uvc[0]
with torch.no_grad():
    mm.s[:, 0] = 1.3; mm.s[:, 1] = 300.5; mm.r.fill_(9.2)
prune_w_mask(mm, None)
model.train(); model.enable_warmup = 0
model.block_skip_gating.requires_grad = True
optimizer = make_optimizer(args, model, args.learning_rate, args.weight_decay)
scheduler = WarmupCosineSchedule(optimizer, warmup_steps=500, t_total=100000)
mixup = Mixup(mixup_alpha=args.mixup, cutmix_alpha=args.cutmix, prob=args.mixup_prob, switch_prob=args.mixup_switch_prob,
              label_smoothing=args.smoothing, num_classes=1000)
crit = DistillationLoss(SoftTargetCrossEntropy(), teacher, "soft", args.distillation_alpha, args.distillation_tau)
step = Stage1Step(args, model, model, optimizer, scheduler, crit, mixup, uvc)
x = torch.randn(B, 3, 224, 224, device=device); y = torch.randint(0, 1000, (B,), device=device)
for _ in range(4):
    out = step(x.clone(), y)
torch.cuda.synchronize()
n = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    out = step(x.clone(), y)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"T2T-ViT-14 Stage-1 step, B={B}: {ms:.2f} ms/step, {B / ms * 1e3:.0f} images/sec, loss {float(out['loss']):.4f}")
if os.environ.get("UVC_STEP_TIMING"):
    print(step.timing_report())
