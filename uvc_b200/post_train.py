"""Stage 2 of UVC — post-training of the weights under the FIXED layout found by Stage 1 — on the sm_100a engine.

Mirror of the reference's `UVC/post_train.py`: load a Stage-1 checkpoint (weights + `.mask` buffers + gates, strict), hard-skip
blocks whose gate prefers "skip" (models/model_distilled.py:496-500), keep every masked weight at zero, AdamW + per-epoch cosine
schedule (timm `create_optimizer` / `create_scheduler`, re-stated: timm is an un-vendored dependency), soft distillation from the
dense teacher, evaluate every epoch and keep the best checkpoint.

The reference re-multiplies every weight by its mask before every step (:357-360, ~150 launches); here the mask rides inside the
fused clip+AdamW sweep (the update is multiplied by the mask), so masked weights are exactly 0 at every forward, identically.
"""
import math
import os
import time

import torch

from . import joint_train as jt
from .utils.data_utils import get_loader
from .utils.ddp import DistributedDataParallel as DDP
from .utils.dist_util import get_world_size
from .utils.optim import FusedClipAdamW


def apply_masks(model):
    """weight *= mask for every module that carries one (post_train.py:228-231,357-360)"""
    with torch.no_grad():
        for _, m in model.named_modules():
            if hasattr(m, "mask"):
                m.weight.mul_(m.mask)


def param_groups_weight_decay(model, weight_decay):
    """timm.optim.optim_factory.add_weight_decay: no decay for 1-d tensors, biases and model.no_weight_decay() names"""
    skip = model.no_weight_decay() if hasattr(model, "no_weight_decay") else set()
    decay, no_decay = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        (no_decay if (p.ndim <= 1 or name.endswith(".bias") or name in skip) else decay).append(p)
    return [{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": weight_decay}]


class CosineEpochSchedule:
    """timm CosineLRScheduler stepped per epoch with linear warm-up (create_scheduler defaults of the DeiT recipe:
    warmup_epochs 5 from warmup_lr 1e-6, min_lr 1e-5)."""

    def __init__(self, optimizer, epochs, base_lr, warmup_epochs=5, warmup_lr=1e-6, min_lr=1e-5):
        self.opt, self.epochs, self.base_lr = optimizer, epochs, base_lr
        self.warmup_epochs, self.warmup_lr, self.min_lr = warmup_epochs, warmup_lr, min_lr

    def get_epoch_values(self, epoch):
        if epoch < self.warmup_epochs:
            lr = self.warmup_lr + epoch * (self.base_lr - self.warmup_lr) / self.warmup_epochs
        else:
            lr = self.min_lr + 0.5 * (self.base_lr - self.min_lr) * (1 + math.cos(math.pi * epoch / self.epochs))
        return [lr]

    def step(self, epoch):
        lr = self.get_epoch_values(epoch)[0]
        for g in self.opt.param_groups:
            g["lr"] = lr


class Stage2Step:
    """The body of Stage 2's hot loop (post_train.py:351-383) as one callable, shared by `post_training()` and `bench.py`: the weights were
    masked once (`apply_masks`) and stay masked through the optimiser's update (uvc_clip_adamw_flags), which is the reference's
    `weight *= mask` before every step without ~150 elementwise launches."""

    def __init__(self, args, model, ddp_model, optimizer, criterion, mixup_fn):
        self.args, self.model, self.ddp_model, self.optimizer, self.criterion, self.mixup_fn = args, model, ddp_model, optimizer, criterion, mixup_fn
        self.global_step = 0

    def __call__(self, x, y):
        if len(x) % 2 != 0:
            x, y = x[:-1], y[:-1]
        if self.mixup_fn is not None:
            x, y = self.mixup_fn(x, y)
        if hasattr(self.criterion, "prefetch_teacher"):
            self.criterion.prefetch_teacher(x)
        outputs, _ = self.ddp_model(x)
        loss = self.criterion(x, outputs, y)
        loss.backward()
        self.optimizer.step()
        self.global_step += 1
        self.optimizer.zero_grad()
        return {"loss": loss}


def set_compact_training(model, mode):
    """Stage-2 training on the physically compacted model (uvc_b200/compact.py:EngineLayout; include/uvc_b200.h uvc_vit_layout).
    mode 0: masked-dense, the reference's arithmetic (post_train.py:357-360 multiplies by the zeros);
    mode 1: skipped blocks, fully pruned heads and pruned neurons are not computed -- same logits, same gradients at every live position; the
            attn.proj columns of fully pruned heads receive no gradient (they are re-masked to zero anyway, but the reference's clip norm counts them);
    mode 2: as 1 but pruned heads stay computed, which reproduces the reference's gradient -- and clip norm -- everywhere."""
    mode = int(mode or 0)
    if mode == 0:
        model.compact_layout = None
        return None
    from .compact import engine_layout_for
    model.compact_layout = engine_layout_for(model, keep_pruned_heads=(mode == 2))
    return model.compact_layout


def post_training(args, model, mixup_fn=None, criterion=None, lr=None, weight_decay=None, epochs=None):
    """post_train.py:270-402"""
    epochs = epochs if epochs is not None else args.epochs
    lr = lr if lr is not None else args.learning_rate
    weight_decay = weight_decay if weight_decay is not None else args.weight_decay
    if args.local_rank in [-1, 0]:
        os.makedirs(os.path.join(args.output_dir, args.name), exist_ok=True)
    print("Starting post training")
    train_loader, test_loader = get_loader(args)
    model.enable_block_gating = 0            # hard skip by gate comparison, as the Stage-2 constructor default does
    model.block_skip_gating.requires_grad = False
    apply_masks(model)
    set_compact_training(model, getattr(args, "compact_train", 0))
    ddp_model = DDP(model, message_size=250000000, gradient_predivide_factor=get_world_size(), delay_allreduce=True) \
        if args.local_rank != -1 and get_world_size() > 1 else model
    args.lr = lr * args.train_batch_size * get_world_size() / 512.0
    masks = {m.weight: m.mask for _, m in model.named_modules() if hasattr(m, "mask")}
    optimizer = FusedClipAdamW(param_groups_weight_decay(model, weight_decay), lr=args.lr, weight_decay=weight_decay,
                               max_grad_norm=args.max_grad_norm, model=model, masks=masks)
    scheduler = CosineEpochSchedule(optimizer, epochs, args.lr)
    print("***** [Stage 2] Post Training *****")
    print("  Instantaneous batch size per GPU = %d" % args.train_batch_size)
    model.zero_grad()
    jt.set_seed(args)
    losses = jt.AverageMeter()
    global_step, best_acc = 0, 0
    run_step = Stage2Step(args, model, ddp_model, optimizer, criterion, mixup_fn)
    for epoch in range(epochs):
        model.train()
        print("=" * 60)
        print(f"Start training [Epoch {epoch}]")
        scheduler.step(epoch)
        t0 = time.time()
        for step, (x, y) in enumerate(train_loader):
            x, y = x.to(args.device, non_blocking=True), y.to(args.device, non_blocking=True)
            loss = run_step(x, y)["loss"]
            global_step += 1
            if (step + 1) % max(1, getattr(args, "print_every", 50)) == 0 and args.local_rank in [-1, 0]:
                losses.update(loss.item())
                print(f"Training [{global_step} Steps] [LR: {scheduler.get_epoch_values(epoch)[0]:.6f} | Loss: {losses.val:.3f}] "
                      f"{(step + 1) * x.shape[0] * get_world_size() / (time.time() - t0):.1f} img/s")
        if args.local_rank in [-1, 0]:
            accuracy = valid(args, model, None, test_loader, global_step)
            if best_acc < accuracy:
                jt.save_model(args, model, None, global_step)
                best_acc = accuracy
    return best_acc


class CompactEval(torch.nn.Module):
    """The physically compacted model (uvc_b200/compact.py) behind the `(x, tau, ratio) -> (logits, macs)` call the validation loop makes:
    skipped blocks, fully pruned heads and pruned neurons are not computed at all instead of being multiplied by zeros."""

    def __init__(self, model):
        super().__init__()
        from .compact import CompactViT, compact_state_dict, compile_layout
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        layout = compile_layout(sd, model.blocks[0].attn.num_heads)
        self.runner = CompactViT(compact_state_dict(sd, layout), eps=model.norm.eps).to(model.cls_token.device)

    def forward(self, x, tau=-1, number=0.9):
        return self.runner(x), None


def valid(args, model, writer, test_loader, global_step):
    apply_masks(model)           # post_train.py:228-231
    if getattr(args, "compact_eval", 0):
        from ._lib import operand_f16_for
        C_, H = model.embed_dim, model.blocks[0].attn.num_heads
        if operand_f16_for(model, C_ // H, model.patch_embed.num_patches + 1, C_, model.blocks[0].mlp.fc1.out_features):
            # the whole-model engine at the live widths (uvc_vit_layout): token / patch gates and jumping connections work unchanged
            prev = getattr(model, "compact_layout", None)
            if prev is None:
                set_compact_training(model, 1)
            try:
                return jt.valid(args, model, writer, test_loader, global_step)
            finally:
                model.compact_layout = prev
        if args.enable_patch_gating == 2 or getattr(model, "enable_patch_gating", 0) or getattr(model, "enable_jumping", 0):
            print("--compact_eval: token / patch gates and jumping connections are not in the per-operator compact runner; validating the masked-dense model")
        else:
            return jt.valid(args, CompactEval(model), writer, test_loader, global_step)
    return jt.valid(args, model, writer, test_loader, global_step)


def build_parser():
    p = jt.build_parser()
    p.add_argument("--checkpoint_dir", default=None, type=str, help="Stage-1 checkpoint (state dict with masks and gates)")
    p.add_argument("--epochs", default=120, type=int)
    p.add_argument("--compact_train", default=0, type=int, choices=[0, 1, 2],
                   help="train Stage 2 on the physically compacted model: 1 = blocks + pruned heads + pruned neurons removed, 2 = blocks + neurons only "
                        "(exact reference clip norm); 0 = masked-dense as the reference")
    p.add_argument("--compact_eval", default=0, type=int,
                   help="validate through the physically compacted model (skipped blocks / pruned heads / pruned neurons removed; same logits)")
    return p


def main(argv=None):
    import torch.distributed as dist
    from datetime import timedelta
    from .models import CONFIGS
    from .utils.losses import DistillationLoss
    from .utils.mixup import Mixup, SoftTargetCrossEntropy, LabelSmoothingCrossEntropy
    args = build_parser().parse_args(argv)
    config = CONFIGS[args.model_type]
    args.local_rank = int(os.environ.get("LOCAL_RANK", args.local_rank))
    if "WORLD_SIZE" in os.environ and int(os.environ["WORLD_SIZE"]) > 1:
        torch.cuda.set_device(args.local_rank)
        dist.init_process_group(backend='nccl', timeout=timedelta(minutes=60))
    else:
        args.local_rank = -1 if "LOCAL_RANK" not in os.environ else args.local_rank
    args.n_gpu = 1
    args.device = torch.device("cuda", max(args.local_rank, 0))
    jt.set_seed(args)
    args.num_classes = {"cifar10": 10, "cifar100": 100}.get(args.dataset, 1000)
    model = jt.make_model(args, config, gumbel_hard=True)
    for _, m in model.named_modules():      # masks registered BEFORE the strict load so the layout is restored (post_train.py:176-186)
        if hasattr(m, "weight"):
            m.register_buffer("mask", torch.ones_like(m.weight))
    if args.checkpoint_dir:
        # Strict, as the reference (post_train.py:676-683) -- except for `patch_gating`: Stage 1 with `--enable_patch_gating 1` (the shipped
        # run_uvc_train.sh) attaches that Parameter to the model after construction (uvc_utils.py:287-288), so it lands in the checkpoint,
        # while the Stage-2 model is built without it (post_train.py:150-155) and the reference's own strict load would reject the file.
        # Stage 2 never applies the patch gate, so the key is dropped with a notice instead of failing.
        ck = torch.load(args.checkpoint_dir, map_location='cpu')
        for key in ("model", "state_dict_ema", "state_dict"):
            if isinstance(ck, dict) and key in ck:
                ck = ck[key]
                break
        if "patch_gating" in ck and model.patch_gating is None:
            print("[post_train] dropping `patch_gating` from the Stage-1 checkpoint: the Stage-2 model has no patch gate (reference post_train.py:150-155)")
            ck = {k: v for k, v in ck.items() if k != "patch_gating"}
        model.load_state_dict(ck, strict=True)
    model.to(args.device)
    model.flatten_parameters()
    mixup_fn = Mixup(mixup_alpha=args.mixup, cutmix_alpha=args.cutmix, prob=args.mixup_prob, switch_prob=args.mixup_switch_prob,
                     label_smoothing=args.smoothing, num_classes=args.num_classes) if (args.mixup > 0 or args.cutmix > 0) else None
    base = SoftTargetCrossEntropy() if args.mixup > 0 else (LabelSmoothingCrossEntropy(args.smoothing) if args.smoothing else torch.nn.CrossEntropyLoss())
    teacher = None
    if args.distillation_type != 'none':
        teacher = jt.make_model(args, config, gumbel_hard=True)
        path = args.teacher_path or args.model_path
        if path is not None:
            jt.load_checkpoint(teacher, path)
        teacher.to(args.device).eval()
    criterion = DistillationLoss(base, teacher, args.distillation_type, args.distillation_alpha, args.distillation_tau)
    post_training(args, model, mixup_fn, criterion)


if __name__ == "__main__":
    main()
