"""Stage 1 of UVC — joint training of weights, block gates and the ADMM sparsity variables — on the sm_100a engine.

Mirror of the reference's `UVC/joint_train.py`: same command-line flags (:684-879), same phases
(warm-up with fixed [.5,.5] gates -> UVC/ADMM training -> inline post-training), same per-step order
(:395-450): mixup -> student forward -> DistillationLoss (teacher forward inside) -> backward (+ gradient all-reduce) ->
clip + AdamW -> LR schedule -> ADMM step (`uvc_optimizer`) -> zero_grad; same prints / JSON traces / checkpoints.

    torchrun --nproc-per-node 8 -m uvc_b200.joint_train --uvc_train --model_type deit_small_patch16_224 \
        --distillation-type soft --distillation-alpha 0.1 --train_batch_size 128 --budget 0.5 --dataset synthetic ...

What differs from the reference is only where the work runs: model forward/backward, loss, clip+AdamW and the ADMM update
are calls into libuvc_sm100.so (see include/uvc_b200.h); apex DDP/amp are replaced by `utils.ddp.DistributedDataParallel`
(one flat NCCL all-reduce).  `--dataset synthetic` feeds ImageNet-shaped random tensors (no dataset on the benchmark box).
"""
import argparse
import json
import os
import random
import time
from datetime import timedelta
from functools import partial

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from .models import CONFIGS, DistilledVisionTransformer
from .utils.data_utils import get_loader
from .utils.ddp import DistributedDataParallel as DDP, broadcast_seed
from .utils.dist_util import get_world_size
from .utils.losses import DistillationLoss
from .utils.mixup import LabelSmoothingCrossEntropy, Mixup, SoftTargetCrossEntropy
from .utils.optim import FusedClipAdamW
from .utils.scheduler import WarmupCosineSchedule, WarmupLinearSchedule
from .uvc_optimizer import build_minimax_model, uvc_optimizer, uvc_optimizer_gating
from .uvc_utils import PresetLRScheduler, prune_w_mask

DEIT_FAMILY = ["deit_tiny_patch16_224", "deit_small_patch16_224", "deit_base_patch16_224"]


class AverageMeter(object):
    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def get_tau(max_tau, min_tau, ite, total):
    """linear ramp from min_tau to max_tau over training (joint_train.py:83-85; the token-gate temperature RISES 0.1 -> 10)"""
    return min_tau + (max_tau - min_tau) * ite / total


def complex_accuracy(output, target, topk=(1,)):
    maxk = max(topk)
    _, pred = output.topk(maxk, 1, True, True)
    correct = pred.t().eq(target.view(1, -1).expand_as(pred.t()))
    return [correct[:k].reshape(-1).float().sum(0) * (100.0 / target.size(0)) for k in topk]


def save_model(args, model, minimax_model, global_step):
    m = model.module if hasattr(model, 'module') else model
    os.makedirs(os.path.join(args.output_dir, args.name), exist_ok=True)
    path = os.path.join(args.output_dir, args.name, f"{args.model_type}_{global_step}.pth.tar")
    torch.save(m.state_dict(), path)       # weights + every `.mask` buffer + gates: the layout travels to Stage 2 this way
    print("Saved model checkpoint to [DIR: %s]" % os.path.join(args.output_dir, args.name))


def count_mask(model):
    total = 0
    for _, p in model.named_modules():
        if hasattr(p, "mask"):
            total += p.mask.sum()
    return total / 1e6


def count_parameters(model):
    return sum(p.numel() for p in model.parameters() if p.requires_grad) / 1000000


def make_model(args, config, gumbel_hard):
    if "t2t" in args.model_type:
        # joint_train.py:143-148: t2t_vit_14() with its defaults; the backbone Blocks run on the engine, tokens_to_token feeds `pe_in`
        from .T2TViT.models import t2t_vit_14
        return t2t_vit_14(gumbel_hard=gumbel_hard, num_classes=args.num_classes)
    if "deit" not in args.model_type:
        raise NotImplementedError(f"--model_type {args.model_type}: the sm_100a hot path covers the DeiT family ({', '.join(DEIT_FAMILY)}) and t2t_vit_14")
    return DistilledVisionTransformer(enable_dist=args.enable_deit, patch_size=config.patch_size, embed_dim=config.embed_dim, depth=config.depth,
                                      num_heads=config.num_heads, mlp_ratio=4, qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                                      drop_rate=0, gumbel_hard=gumbel_hard, num_classes=args.num_classes)


def load_checkpoint(model, path, strict=False):
    ck = torch.hub.load_state_dict_from_url(path, map_location='cpu', check_hash=True) if path.startswith('https') \
        else torch.load(path, map_location='cpu')
    for key in ("model", "state_dict_ema", "state_dict"):
        if isinstance(ck, dict) and key in ck:
            ck = ck[key]
            break
    return model.load_state_dict(ck, strict=strict)


def setup(args):
    """model + `.mask` buffers (joint_train.py:122-175)"""
    config = CONFIGS[args.model_type]
    args.num_classes = {"cifar10": 10, "cifar100": 100}.get(args.dataset, 1000)
    model = make_model(args, config, gumbel_hard=False)
    if args.pretrained and args.model_path is not None:
        if args.local_rank in [-1, 0]:
            print(f"Loading checkpoint for model from ====> {args.model_path}")
        load_checkpoint(model, args.model_path)
    model.to(args.device)
    for _, p in model.named_modules():
        if hasattr(p, "weight"):
            p.register_buffer("mask", torch.ones_like(p.weight))
    args.total_param = count_mask(model)
    return args, model


def set_seed(args):
    random.seed(args.seed)
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    if args.n_gpu > 0:
        torch.cuda.manual_seed_all(args.seed)


def get_uvc_layers(model, args=None):
    """W1 = attn.proj, W2 = mlp.fc1, W3 = mlp.fc2 of every block + index tables (joint_train.py:530-564)."""
    layer_names = {None: None}
    uvc_layers = {"W1": [], "W2": [], "W3": []}
    for name, m in model.named_modules():
        if not hasattr(m, "in_features"):
            continue
        if "attn.proj" in name:
            layer_names[m] = name; uvc_layers["W1"].append(m); m.uvc_s = 0
        elif "mlp.fc2" in name:
            layer_names[m] = name; uvc_layers["W3"].append(m); m.uvc_s = 0
        if "mlp.fc1" in name:
            layer_names[m] = name; uvc_layers["W2"].append(m); m.uvc_s = 0
    d = {"s_dict": {}, "r_dict": {}}
    for i, m in enumerate(uvc_layers["W1"]):
        d["s_dict"][m] = [i, 0]; d["r_dict"][m] = i
    for i, m in enumerate(uvc_layers["W3"]):
        d["s_dict"][m] = [i, 1]
    return layer_names, uvc_layers, d


class EvalStep:
    """The body of the validation loop (joint_train.py:199-246, post_train.py:209-265) as one callable, shared by `valid()` and `bench.py`:
    eval forward -> CrossEntropyLoss -> top-1.  The reference reads loss and accuracy back after every batch (two host syncs per batch); here the
    sums stay on the device and `result()` reads them once, giving the same two averages (top-1 weighted by batch size, loss per batch)."""

    def __init__(self, args, model):
        self.args, self.model = args, model
        self.loss_fct = torch.nn.CrossEntropyLoss()
        self.tau = 1 if args.enable_patch_gating == 2 else -1
        self.acc = None
        self.images = self.batches = 0

    def __call__(self, x, y):
        with torch.no_grad():
            logits, _ = self.model(x, self.tau, self.args.patch_ratio)
            loss = self.loss_fct(logits, y)
            correct = (logits.argmax(dim=1) == y).sum()
            if self.acc is None:
                self.acc = torch.zeros(2, device=logits.device, dtype=torch.float64)    # correct predictions, sum of batch losses
            self.acc += torch.stack([correct.double(), loss.double()])
        self.images += int(x.size(0)); self.batches += 1         # host-known counts stay on the host (a device scalar built from a Python number is a synchronous copy)
        return {"loss": loss}

    def result(self):
        """(top-1 in percent, mean loss) -- the one device->host read of the loop"""
        if self.acc is None:
            return 0.0, 0.0
        c, ls = self.acc.tolist()
        return 100.0 * c / max(self.images, 1), ls / max(self.batches, 1)


def valid(args, model, writer, test_loader, global_step):
    if args.local_rank in [-1, 0]:
        print("***** Running Validation *****")
        print("  Num steps = %d" % len(test_loader))
        print("  Batch size = %d" % args.eval_batch_size)
    model.eval()
    step = EvalStep(args, model)
    for x, y in test_loader:
        step(x.to(args.device, non_blocking=True), y.to(args.device, non_blocking=True))
    top1, loss = step.result()
    if args.local_rank in [-1, 0]:
        print("\nValidation Results")
        print("Global Steps: %d" % global_step)
        print("Valid Loss: %2.5f" % loss)
        print("Valid Accuracy: %2.5f" % top1)
    return top1


class Stage1Step:
    """The body of the hot loop (joint_train.py:395-450) as one callable, shared by `train()` and `bench.py`."""

    def __init__(self, args, model, ddp_model, optimizer, scheduler, criterion, mixup_fn, uvc_args, zlr_scheduler=None):
        self.args, self.model, self.ddp_model = args, model, ddp_model
        self.optimizer, self.scheduler, self.criterion, self.mixup_fn = optimizer, scheduler, criterion, mixup_fn
        self.uvc_args = uvc_args
        self.zlr_scheduler = zlr_scheduler
        self.global_step = 0
        self.gating_grad_list = []
        self.uvc_fn = uvc_optimizer if args.enable_pruning else uvc_optimizer_gating
        # the ADMM step's host return values are only printed every log_interval steps: fetch them lazily (no device sync per step)
        self.uvc_kw = {"lazy": True} if args.enable_pruning else {}
        self.last = {}
        self._timing = [] if os.environ.get("UVC_STEP_TIMING") else None

    def _mark(self, name):
        """UVC_STEP_TIMING=1: CPU wall clock + CUDA event at every phase boundary (bring-up aid; off by default)"""
        if self._timing is None:
            return
        ev = torch.cuda.Event(enable_timing=True); ev.record()
        self._timing.append((name, time.perf_counter(), ev))

    def timing_report(self):
        if not self._timing:
            return ""
        torch.cuda.synchronize()
        agg, prev = {}, None
        for name, t, ev in self._timing:
            if prev is not None and name != "start":
                a = agg.setdefault(name, [0.0, 0.0, 0])
                a[0] += (t - prev[1]) * 1e3; a[1] += prev[2].elapsed_time(ev); a[2] += 1
            prev = (name, t, ev)
        return "\n".join(f"  {k:12s} cpu {v[0]/v[2]:7.3f} ms   gpu-timeline {v[1]/v[2]:7.3f} ms" for k, v in agg.items())

    def __call__(self, x, y, epoch=0, total_steps=1):
        args = self.args
        self._mark("start")
        if len(x) % 2 != 0:
            x, y = x[:-1], y[:-1]
        tau = get_tau(10, 0.1, self.global_step, total_steps) if args.enable_patch_gating == 2 else -1
        if self.mixup_fn is not None:
            x, y = self.mixup_fn(x, y)
        self._mark("mixup")
        if hasattr(self.criterion, "prefetch_teacher"):
            self.criterion.prefetch_teacher(x)        # the dense teacher forward runs on a side stream next to the student forward
        outputs, flops_list = self.ddp_model(x, tau, args.patch_ratio)
        self._mark("student_fwd")
        loss = self.criterion(x, outputs, y)
        self._mark("teacher+loss")
        loss.backward()
        self._mark("backward")
        self.optimizer.step()                 # global-norm clip (max_grad_norm) + AdamW, fused
        self.scheduler.step()
        self._mark("optimizer")
        self.global_step += 1
        out = {"loss": loss}
        if args.uvc_train:
            minimax_model, dual_optimizer, s_optimizer, r_optimizer, gating_optimizer = self.uvc_args
            if not minimax_model.model.enable_warmup and self.zlr_scheduler is not None:
                self.zlr_scheduler(dual_optimizer, epoch, "zlr")
            minimax_model.update_gating()
            cur_resource, s_data, r_data, gating_data, self.gating_grad_list = self.uvc_fn(
                self.optimizer, minimax_model, s_optimizer, r_optimizer, gating_optimizer, dual_optimizer, args, {"global_step": self.global_step},
                [], flops_list, args.z_grad_clip, self.global_step, args.gating_interval, self.gating_grad_list, **self.uvc_kw)
            out.update(cur_resource=cur_resource, s=s_data, r=r_data, gating=gating_data)
        self._mark("admm")
        self.optimizer.zero_grad()
        self._mark("zero_grad")
        self.last = out
        return out


def make_optimizer(args, model, lr, weight_decay):
    m = model.module if hasattr(model, "module") else model
    return FusedClipAdamW(m.parameters(), lr=lr, weight_decay=weight_decay, max_grad_norm=args.max_grad_norm, model=m)


def train(args, model, uvc_args=None, mixup_fn=None, criterion=None):
    """Stage 1 (joint_train.py:249-528)"""
    args.train_batch_size = args.train_batch_size // args.gradient_accumulation_steps
    train_loader, test_loader = get_loader(args)
    model.flatten_parameters()
    optimizer = make_optimizer(args, model, args.learning_rate, args.weight_decay)
    t_total = len(train_loader) * args.num_epochs
    Sched = WarmupCosineSchedule if args.decay_type == "cosine" else WarmupLinearSchedule
    scheduler = Sched(optimizer, warmup_steps=args.warmup_steps, t_total=t_total)
    zlr_scheduler = PresetLRScheduler(args.zlr_schedule)
    ddp_model = DDP(model, message_size=250000000, gradient_predivide_factor=get_world_size(), delay_allreduce=True) \
        if args.local_rank != -1 and get_world_size() > 1 else model
    if args.local_rank in [-1, 0]:
        print("***** [Stage 1] Training with ADMM *****")
        print(f"  Total optimization steps = {args.num_steps}")
        print(f"  Instantaneous batch size per GPU = {args.train_batch_size}")
    model.zero_grad()
    set_seed(args)
    losses = AverageMeter()
    best_acc = 0
    s_list, r_list, gating_list = [], [], []
    minimax_model = uvc_args[0]
    step_fn = Stage1Step(args, model, ddp_model, optimizer, scheduler, criterion, mixup_fn, uvc_args, zlr_scheduler)
    model.train()
    epoch = 0
    while epoch <= args.num_epochs:            # (sic) num_epochs + 1 epochs, as the reference (:335)
        epoch += 1
        step_fn.gating_grad_list = []
        if epoch <= args.warmup_epochs:
            stage = "Warm Up"
            args.gumbel_hard = 1
            minimax_model.model.enable_warmup = 1
            minimax_model.model.block_skip_gating.requires_grad = False
            for params in optimizer.param_groups:
                params['lr'] = args.warmup_lr
        else:
            stage = "UVC Train"
            minimax_model.model.enable_warmup = 0
            args.enable_warmup = 0
            args.gumbel_hard = 0
            minimax_model.model.block_skip_gating.requires_grad = True
            if epoch == args.warmup_epochs + 1 and args.warmup_reset:
                optimizer = make_optimizer(args, model, args.learning_rate, args.weight_decay)
                scheduler = Sched(optimizer, warmup_steps=args.warmup_steps, t_total=t_total)
                step_fn.optimizer, step_fn.scheduler = optimizer, scheduler
        prune_w_mask(minimax_model, optimizer)
        remained_param = count_mask(minimax_model.model)
        if args.local_rank in [-1, 0]:
            print("=" * 60)
            print(f"Start [Epoch {epoch}] at Stage {stage}")
            print(f"[Initial Sparsity|Epoch {epoch}] Parameter size: {(remained_param):.2f}M / {args.total_param:.2f}M = {(remained_param)/args.total_param*100:.2f}%")
        if stage == "UVC Train":
            minimax_model.update_eps()
        t0 = time.time()
        for step, (x, y) in enumerate(train_loader):
            x, y = x.to(args.device, non_blocking=True), y.to(args.device, non_blocking=True)
            out = step_fn(x, y, epoch, args.num_epochs * len(train_loader))
            if (step + 1) % max(1, args.print_every) == 0:
                losses.update(out["loss"].item())
                if args.local_rank in [-1, 0]:
                    dt = time.time() - t0
                    print(f"Stage [{epoch} / {args.num_epochs} Epochs] [{step_fn.global_step} Steps] [LR: {scheduler.get_last_lr()[0]:.6f} | Loss: {losses.val:.3f} | "
                          f"Flops: {float(out.get('cur_resource', 1.0))*100:.2f}%] {(step + 1) * x.shape[0] * get_world_size() / dt:.1f} img/s")
            if args.uvc_train and step_fn.global_step % args.log_interval == 0 and args.local_rank in [-1, 0]:
                s_list.append(out["s"].tolist()); r_list.append(out["r"].tolist())
                if out.get("gating") is not None:
                    gating_list.append(out["gating"].tolist())
                os.makedirs(os.path.join(args.output_dir, args.name), exist_ok=True)
                for nm, lst in (("s", s_list), ("r", r_list), ("gating", gating_list)):
                    with open(os.path.join(args.output_dir, args.name, f"{nm}_{args.model_type}.json"), "w") as f:
                        json.dump(lst, f)
        accuracy = valid(args, model, None, test_loader, step_fn.global_step)
        best_acc = max(best_acc, accuracy)
        model.train()
        prune_w_mask(minimax_model, optimizer)
        remained_param = count_mask(minimax_model.model)
        if args.local_rank in [-1, 0]:
            save_model(args, minimax_model.model, minimax_model, epoch)
            print(f"[Epoch {epoch}] Parameter size: {(remained_param):.2f}M / {args.total_param:.2f}M = {(remained_param)/args.total_param*100:.2f}%")
            print(f"Expectation FLOPs: {float(minimax_model.run_resource_fn(args.gumbel_hard))*100}%",
                  f"Real FLOPs: {float(minimax_model.run_resource_fn(gumbel_hard=True))*100}%")
        losses.reset()
    if args.local_rank in [-1, 0]:
        print("Best Accuracy: \t%f" % best_acc)
        print("End Training!")
    return minimax_model


def post_training(args, model, mixup_fn=None, criterion=None):
    """Inline Stage 2 at the end of Stage 1 (joint_train.py:567-678): fixed layout, weights only."""
    from .post_train import post_training as run
    return run(args, model, mixup_fn, criterion, lr=args.post_learning_rate, weight_decay=args.post_weight_decay, epochs=args.post_num_epochs)


def build_parser():
    p = argparse.ArgumentParser()
    a = p.add_argument
    a("--name", default="debug"); a("--dataset", choices=["cifar10", "cifar100", "imagenet", "synthetic"], default="imagenet")
    a("--data_dir", default="/ssd1/shixing/imagenet2012"); a("--num_workers", default=4, type=int)
    a("--model_type", choices=list(CONFIGS.keys()), default="deit_tiny_patch16_224")
    a("--model_path", default=None); a("--pretrained_dir", type=str, default=None); a("--pretrained", type=int, default=1)
    a("--output_dir", default="../result/output/uvc_train", type=str); a("--img_size", default=224, type=int)
    a("--train_batch_size", default=1024, type=int); a("--eval_batch_size", default=64, type=int); a("--eval_every", default=1000, type=int)
    a("--learning_rate", default=1e-4, type=float); a("--weight_decay", default=0.05, type=float)
    a("--num_steps", default=10000, type=int); a("--num_epochs", default=20, type=int)
    a("--decay_type", choices=["cosine", "linear"], default="cosine"); a("--warmup_steps", default=500, type=int)
    a("--max_grad_norm", default=1.0, type=float)
    a("--local_rank", "--local-rank", type=int, default=int(os.environ.get("LOCAL_RANK", 0)))
    a('--seed', type=int, default=42); a('--gradient_accumulation_steps', type=int, default=1)
    a('--fp16', action='store_true'); a('--fp16_opt_level', type=str, default='O2'); a('--loss_scale', type=float, default=0)
    a('--uvc_train', action='store_true', default=True)
    a('--soptim', default='sgd'); a('--roptim', default='sgd'); a('--zlr_schedule_list', default="10,20,30,40,50", type=str)
    a('--ylr', default=1e-4, type=float); a('--plr', default=1e-4, type=float); a('--slr', default=0.02, type=float)
    a('--rlr', default=0.02, type=float); a('--glr', default=1e-3, type=float); a('--log_interval', default=2000, type=int)
    a('--save_budgets', default='0.6, 0.5, 0.4'); a('--budget', default=0.5, type=float); a('--sl2wd', default=0.0, type=float)
    a('--verbose', default=True, action='store_true')
    a('--mixup', type=float, default=0.8); a('--cutmix', type=float, default=1.0); a('--cutmix-minmax', type=float, nargs='+', default=None)
    a('--mixup-prob', type=float, default=0.8); a('--mixup-switch-prob', type=float, default=0.5); a('--mixup-mode', type=str, default='batch')
    a('--teacher-model', default=None, type=str); a('--teacher-path', type=str, default=None)
    a('--distillation-type', default='hard', choices=['none', 'soft', 'hard'], type=str)
    a('--distillation-alpha', default=0.5, type=float); a('--distillation-tau', default=1.0, type=float); a('--smoothing', type=float, default=0.1)
    a("--post_learning_rate", default=1e-3, type=float); a("--post_weight_decay", default=0.05, type=float); a("--post_num_epochs", default=100, type=int)
    a("--use_distribute", default=1, type=int); a("--enable_writer", default=0, type=int); a("--flops_with_mhsa", type=int, default=1)
    a("--enable_block_gating", type=int, default=1); a("--enable_part_gating", type=int, default=0); a("--enable_jumping", type=int, default=0)
    a("--enable_deit", type=int, default=0); a("--enable_pruning", type=int, default=1); a("--enable_patch_gating", type=int, default=2)
    a("--patch_ratio", type=float, default=0.9); a('--z_grad_clip', default=0.5, type=float); a('--gating_interval', default=100, type=int)
    a('--gating_weight', default=5, type=float); a('--patch_weight', default=5, type=float); a('--patch_l1_weight', default=0.01, type=float)
    a('--patchlr', default=0.01, type=float); a('--patchloss', default="l1", type=str); a('--use_gumbel', default=1, type=int)
    a('--eps', default=0.1, type=float); a('--eps_decay', default=0.92, type=float); a('--enable_warmup', default=1, type=int)
    a('--warmup_epochs', default=5, type=int); a('--warmup_lr', default=1e-4, type=float); a('--warmup_reset', default=0, type=int)
    a("--gpu_num", type=str, default="0, 1")
    # additions of this implementation (not in the reference)
    a("--print_every", type=int, default=50, help="steps between progress lines (the reference uses a tqdm bar)")
    a("--synthetic_steps", type=int, default=100, help="--dataset synthetic: steps per epoch")
    a("--skip_post_training", type=int, default=0)
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.fp16:
        raise NotImplementedError("--fp16 (apex amp) is outside the sm_100a hot path: every shipped UVC run is fp32 (the engine computes in TF32)")
    config = CONFIGS[args.model_type]
    args.head_size = config.hidden_size // config.transformer["num_heads"]
    args.num_heads = config.transformer["num_heads"]
    args.local_rank = int(os.environ.get("LOCAL_RANK", args.local_rank))
    if "WORLD_SIZE" in os.environ and int(os.environ["WORLD_SIZE"]) > 1:
        torch.cuda.set_device(args.local_rank)
        dist.init_process_group(backend='nccl', timeout=timedelta(minutes=60))
        args.n_gpu = 1
    else:
        args.local_rank = -1 if "LOCAL_RANK" not in os.environ else args.local_rank
        args.n_gpu = 1
    device = torch.device("cuda", max(args.local_rank, 0))
    args.device = device
    args.seed = broadcast_seed(args.seed, device)
    set_seed(args)
    args, model = setup(args)

    mixup_fn = None
    if args.mixup > 0 or args.cutmix > 0.:
        mixup_fn = Mixup(mixup_alpha=args.mixup, cutmix_alpha=args.cutmix, cutmix_minmax=args.cutmix_minmax, prob=args.mixup_prob,
                         switch_prob=args.mixup_switch_prob, mode=args.mixup_mode, label_smoothing=args.smoothing, num_classes=args.num_classes)
    if args.mixup > 0.:
        criterion = SoftTargetCrossEntropy()
    elif args.smoothing:
        criterion = LabelSmoothingCrossEntropy(smoothing=args.smoothing)
    else:
        criterion = torch.nn.CrossEntropyLoss()

    teacher_model = None
    if args.distillation_type != 'none':
        teacher_model = make_model(args, config, gumbel_hard=True)
        path = args.teacher_path or args.model_path
        if args.pretrained and path is not None:
            load_checkpoint(teacher_model, path)
        else:
            teacher_model.load_state_dict({k: v for k, v in model.state_dict().items() if not k.endswith(".mask")}, strict=False)
        teacher_model.to(device)
        teacher_model.eval()
    criterion = DistillationLoss(criterion, teacher_model, args.distillation_type, args.distillation_alpha, args.distillation_tau)

    if args.uvc_train:
        zl = [int(v) for v in str(args.zlr_schedule_list).split(",")]
        gap = max(1, args.num_epochs // len(zl))
        args.zlr_schedule = {i * gap: zl[i] for i in range(len(zl))}
        args.zlr_schedule_list = zl
        layer_names, uvc_layers, uvc_layers_dict = get_uvc_layers(model, args)
        with torch.no_grad():
            model.eval()
            _, flops_list = model(torch.ones(1, 3, args.img_size, args.img_size, device=device), number=args.patch_ratio)
        uvc_args = list(build_minimax_model(model, layer_names, uvc_layers, uvc_layers_dict, args, flops_list))
        uvc_args[0] = uvc_args[0].to(device)
        prune_w_mask(uvc_args[0], None)
        minimax_model = train(args, model, uvc_args=uvc_args, mixup_fn=mixup_fn, criterion=criterion)
        prune_w_mask(minimax_model, None)
        if not args.skip_post_training:
            post_training(args, minimax_model.model, mixup_fn, criterion)
    else:
        raise NotImplementedError("plain fine-tuning without --uvc_train is not part of the UVC hot path")


if __name__ == "__main__":
    main()
