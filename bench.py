#!/usr/bin/env python
"""Benchmark of the UVC hot path: images/sec of the Stage-1 (joint_train) step, DeiT-Small, 50 % FLOPs budget.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one full iteration of the reference's hot loop (joint_train.py:395-450) through this repo's public API
(`uvc_b200.joint_train.Stage1Step`): mixup -> student forward (soft Gumbel block gates, patch gate) -> dense teacher forward
+ soft-target CE + soft KD loss -> backward -> (N>1: flat NCCL all-reduce) -> global-norm clip + AdamW -> LR schedule ->
ADMM primal-dual update (`uvc_optimizer`, non-warm-up) -> zero_grad.  Workload: BASELINE.json configs[2] at its per-GPU size
(DeiT-Small patch16 224, 128 images per GPU, weak scaling), synthetic 224x224 batch, random-init weights, seed 730.

Prints ONE JSON line (rank 0).  `value`: device-resident inputs.  `e2e`: the same step fed from pinned HOST memory each step
(H2D copy on a side stream, double-buffered, inside the timed region) with the loss / ADMM results read back each step.
`roofline`: the dominant kernel (the tcgen05 TF32 GEMM), per-launch CUDA-event timing of every GEMM launch of instrumented
steps run right after the timed region (uvc_gemm_profile).  `cpu_baseline` / `--impl reference`: the oracle port (plain
PyTorch CPU restatement of the same step, oracle/) timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DENSE_FWD_FLOPS = {"deit_tiny_patch16_224": 2.507e9, "deit_small_patch16_224": 9.198e9, "deit_base_patch16_224": 35.128e9,
                   "t2t_vit_14": 9.57e9}   # 2 x MACs of one dense forward, reference accounting

# --config: the BASELINE.json configurations that fit one GPU at their per-GPU size (configs[0] is the CPU numerics gate, covered by tests/)
BENCH_CONFIGS = {
    "small_s1": dict(model="deit_small_patch16_224", stage=1, batch=128, budget=0.5, patch_gating=1,
                     metric="images/sec DeiT-Small UVC@50%FLOPs (Stage-1 joint_train step)",
                     workload="BASELINE.json configs[2] at per-GPU size: DeiT-Small patch16 224 UVC joint_train (ADMM active), budget 0.5, "
                              "soft-distill alpha=0.1, 128 images/GPU, DDP over N GPUs"),
    "tiny_s1": dict(model="deit_tiny_patch16_224", stage=1, batch=512, budget=0.5, patch_gating=1,
                    metric="images/sec DeiT-Tiny UVC@50%FLOPs (Stage-1 joint_train step)",
                    workload="BASELINE.json configs[1]: DeiT-Tiny patch16 224 UVC joint_train (ADMM active), budget 0.5, 512 images/GPU"),
    "base_s2": dict(model="deit_base_patch16_224", stage=2, batch=256, budget=0.5, patch_gating=0,
                    metric="images/sec DeiT-Base post_train @ fixed 50% pruned layout (Stage-2 step)",
                    workload="BASELINE.json configs[3] at per-GPU size: DeiT-Base patch16 224 post_train, fixed 50 % layout (blocks 8, 10 skipped; per live "
                             "block the 3 lowest-norm heads, 16 lowest dims of each surviving head and 1417 lowest neurons pruned), 256 images/GPU, "
                             "soft-distill from a dense teacher"),
    "t2t_s1": dict(model="t2t_vit_14", stage=1, batch=128, budget=0.6, patch_gating=2,
                   metric="images/sec T2T-ViT-14 UVC@60%FLOPs (Stage-1 joint_train step, token slimming on)",
                   workload="BASELINE.json configs[4] at per-GPU size: T2T-ViT-14 UVC joint_train (ADMM active), budget 0.6, token gate (top 90 %), "
                            "128 images/GPU"),
    # the validation / inference path (joint_train.py:199-246, post_train.py:209-265): eval forward + CrossEntropyLoss + top-1
    "small_eval": dict(model="deit_small_patch16_224", stage="eval", batch=256, budget=0.5, patch_gating=0,
                       metric="images/sec DeiT-Small validation step (eval forward + CE + top-1)",
                       workload="validation loop body of joint_train.py:199-246 on the dense DeiT-Small of BASELINE.json configs[2], 256 images/GPU"),
    "base_s2_eval": dict(model="deit_base_patch16_224", stage="eval", batch=256, budget=0.5, patch_gating=0, layout=True,
                         metric="images/sec DeiT-Base validation step @ fixed 50% pruned layout (eval forward + CE + top-1)",
                         workload="validation loop body of post_train.py:209-265 on the DeiT-Base of BASELINE.json configs[3] (fixed 50 % layout: blocks 8, 10 "
                                  "skipped, 3 heads / 16 dims per head / 1417 neurons pruned per live block), 256 images/GPU"),
}


def uvc_args_namespace(H, **over):
    """flags of run_uvc_train.sh (the shipped Stage-1 recipe) that the step reads"""
    a = types.SimpleNamespace(
        head_size=64, num_heads=H, flops_with_mhsa=1, use_gumbel=1, enable_block_gating=1, enable_part_gating=0, enable_patch_gating=1,
        enable_jumping=0, enable_pruning=1, eps=0.1, eps_decay=0.92, enable_warmup=0, soptim="sgd", roptim="sgd", slr=0.02, rlr=0.02,
        glr=0.1, ylr=1e-4, plr=1e-4, zlr_schedule_list=[1, 5, 9, 13, 17], budget=0.5, sl2wd=0.0, gating_weight=5e-4, z_grad_clip=0.5,
        gating_interval=50, patch_ratio=0.9, uvc_train=True, max_grad_norm=1.0, learning_rate=1e-4, weight_decay=0.05,
        distillation_alpha=0.1, distillation_tau=1.0, smoothing=0.1, mixup=0.8, cutmix=1.0, mixup_prob=0.8, mixup_switch_prob=0.5)
    for k, v in over.items():
        setattr(a, k, v)
    return a


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.  The sampler is started before the warm-up (nvidia-smi takes a
    second or more to produce its first line on an 8-GPU box) and every line is stamped on arrival, so the summary can be restricted to the
    timed window; if that window was too short to catch a sample, the samples of the following end-to-end leg (same step, under load) are used
    and the summary says so."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, windows):
        """windows: [(t0, t1, label), ...] in preference order"""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        num = lambda v: v.replace(".", "").isdigit()
        for t0, t1, label in windows:
            rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.05 and len(r) >= 7]
            if rows:
                break
        else:
            rows, label = [r for _, r in self.rows if len(r) >= 7], "whole run"
        sm = [float(r[0]) for r in rows if num(r[0])]
        mx = [float(r[1]) for r in rows if num(r[1])]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        pw = [float(r[2]) for r in rows if num(r[2])]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "window": label}


def bind_to_gpu_numa_node(local):
    """Pin this rank's host threads (and therefore its first-touch pinned staging buffers) to the NUMA node its GPU hangs off: with 8 ranks
    feeding 77 MB per step each, cross-socket staging is what the end-to-end scaling loses first.  Best effort; returns a note for the JSON line."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return "numa_node unknown"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"node {node}: no allowed cpus"
        os.sched_setaffinity(0, cpus)
        return f"node {node} ({len(cpus)} cpus)"
    except Exception as e:
        return f"unavailable ({type(e).__name__})"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"bf16_sustained": float(p["bf16_tflops_sustained"]), "bf16_burst": float(p["bf16_tflops"]), "hbm": float(p["hbm_gbs"]), "src": "measured"}
    except Exception:
        return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "src": "fallback"}


# ------------------------------------------------------------------------------------------------ the GPU arm
def make_student_teacher(cfg, device):
    import torch
    from functools import partial
    from uvc_b200.models import CONFIGS, DistilledVisionTransformer
    if cfg["model"] == "t2t_vit_14":
        from uvc_b200.T2TViT.models import t2t_vit_14
        make = lambda gumbel_hard: t2t_vit_14(gumbel_hard=gumbel_hard)
        H = 6
    else:
        c = CONFIGS[cfg["model"]]
        H = c.num_heads

        def make(gumbel_hard):
            return DistilledVisionTransformer(enable_dist=0, patch_size=16, embed_dim=c.embed_dim, depth=c.depth, num_heads=c.num_heads, mlp_ratio=4,
                                              qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), drop_rate=0, gumbel_hard=gumbel_hard)
    torch.manual_seed(730)
    model = make(cfg["stage"] == 2).to(device)
    teacher = make(True).to(device).eval()
    teacher.load_state_dict(model.state_dict(), strict=False)        # teacher = copy of the initial student (no checkpoints offline)
    return model, teacher, H


def build_gpu_step(cfg, device, world):
    import torch
    from uvc_b200.joint_train import Stage1Step, get_uvc_layers, make_optimizer
    from uvc_b200.utils.ddp import DistributedDataParallel as DDP
    from uvc_b200.utils.losses import DistillationLoss
    from uvc_b200.utils.mixup import Mixup, SoftTargetCrossEntropy
    from uvc_b200.utils.scheduler import WarmupCosineSchedule
    from uvc_b200.uvc_optimizer import build_minimax_model
    from uvc_b200.uvc_utils import prune_w_mask
    model, teacher, H = make_student_teacher(cfg, device)
    args = uvc_args_namespace(H, device=device, local_rank=0 if world > 1 else -1, budget=cfg["budget"], enable_patch_gating=cfg["patch_gating"])
    for _, m in model.named_modules():
        if hasattr(m, "weight"):
            m.register_buffer("mask", torch.ones_like(m.weight))
    layer_names, uvc_layers, uvc_dict = get_uvc_layers(model)
    model.eval()
    with torch.no_grad():
        _, flops_list = model(torch.ones(1, 3, 224, 224, device=device))
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):        # the reference's "** Initial FLOP size" print: stdout carries the ONE JSON line only
        uvc = list(build_minimax_model(model, layer_names, uvc_layers, uvc_dict, args, flops_list))
    mm = uvc[0]
    Fh = model.blocks[0].mlp.fc1.out_features
    mixup = Mixup(mixup_alpha=args.mixup, cutmix_alpha=args.cutmix, prob=args.mixup_prob, switch_prob=args.mixup_switch_prob,
                  label_smoothing=args.smoothing, num_classes=1000)
    crit = DistillationLoss(SoftTargetCrossEntropy(), teacher, "soft", args.distillation_alpha, args.distillation_tau)
    info = {}
    if cfg["stage"] == "eval":
        from uvc_b200.joint_train import EvalStep
        from uvc_b200.post_train import apply_masks, set_compact_training
        from uvc_b200 import compact as cp
        rho = 1.0
        if cfg.get("layout"):
            with torch.no_grad():
                mm.s[:, 0] = 3.0; mm.s[:, 1] = 1417.0; mm.r.fill_(16.0)
                model.block_skip_gating[8] = torch.tensor([1.0, -1.0], device=device); model.block_skip_gating[10] = torch.tensor([1.0, -1.0], device=device)
            prune_w_mask(mm, None)
            apply_masks(model)
            mode = {"dense": 0, "compact": 1, "exact": 2}[os.environ.get("UVC_STAGE2", "compact")]
            elay = set_compact_training(model, mode)
            info["execution"] = "masked-dense" if elay is None else "physically compacted (uvc_vit_layout)"
            if elay is not None:
                info["executed_macs_ratio"] = round(float(elay.executed_macs_ratio()), 4)
            rho = float(cp.macs(cp.compile_layout({k: v.detach().cpu() for k, v in model.state_dict().items()}, H))["budget_ratio"])
            info["rho"] = round(rho, 4)
        model.eval()
        model.enable_block_gating = 0
        args.enable_patch_gating = 0
        step = EvalStep(args, model)
        info["step_flops_per_image"] = rho * DENSE_FWD_FLOPS[cfg["model"]]
        return step, model, info
    if cfg["stage"] == 1:
        with torch.no_grad():           # mid-training ADMM state: the selections / prox / dual updates all do real work
            mm.s[:, 0] = 1.3; mm.s[:, 1] = 0.26 * Fh + 0.5; mm.r.fill_(9.2)
        prune_w_mask(mm, None)
        model.train()
        model.enable_warmup = 0
        model.block_skip_gating.requires_grad = True
        model.flatten_parameters()
        optimizer = make_optimizer(args, model, args.learning_rate, args.weight_decay)
        scheduler = WarmupCosineSchedule(optimizer, warmup_steps=500, t_total=100000)
        ddp = DDP(model, gradient_predivide_factor=world, delay_allreduce=True) if world > 1 else model
        step = Stage1Step(args, model, ddp, optimizer, scheduler, crit, mixup, uvc)
        info["step_flops_per_image"] = 4 * DENSE_FWD_FLOPS[cfg["model"]]     # student fwd + bwd (2x) + dense teacher fwd, reference MAC accounting
    else:
        # Stage 2 (post_train.py:351-383) on the fixed ~50 % layout of SURVEY.md 8(d): the masks are written by this repo's own device selection
        # (prune_w_mask) from s / r, blocks 8 and 10 are switched off through their gates, weights are masked once and stay masked through the update
        from uvc_b200.post_train import Stage2Step, apply_masks, param_groups_weight_decay
        from uvc_b200.utils.optim import FusedClipAdamW
        from uvc_b200 import compact as cp
        with torch.no_grad():
            mm.s[:, 0] = 3.0; mm.s[:, 1] = 1417.0; mm.r.fill_(16.0)
            model.block_skip_gating[8] = torch.tensor([1.0, -1.0], device=device); model.block_skip_gating[10] = torch.tensor([1.0, -1.0], device=device)
        prune_w_mask(mm, None)
        model.train()
        model.enable_block_gating = 0
        model.block_skip_gating.requires_grad = False
        model.flatten_parameters()
        apply_masks(model)
        from uvc_b200.post_train import set_compact_training
        mode = {"dense": 0, "compact": 1, "exact": 2}[os.environ.get("UVC_STAGE2", "compact")]
        elay = set_compact_training(model, mode)
        info["stage2_execution"] = {0: "masked-dense (the reference's arithmetic: zeros are multiplied)",
                                    1: "physically compacted: skipped blocks, fully pruned heads and pruned neurons are not computed",
                                    2: "physically compacted (blocks + neurons; pruned heads kept for the reference's exact clip norm)"}[mode]
        if elay is not None:
            info["executed_macs_ratio"] = round(float(elay.executed_macs_ratio()), 4)
        masks = {m.weight: m.mask for _, m in model.named_modules() if hasattr(m, "mask")}
        lr = 5e-4 * cfg["batch"] * world / 512.0
        optimizer = FusedClipAdamW(param_groups_weight_decay(model, args.weight_decay), lr=lr, weight_decay=args.weight_decay,
                                   max_grad_norm=args.max_grad_norm, model=model, masks=masks)
        ddp = DDP(model, gradient_predivide_factor=world, delay_allreduce=True) if world > 1 else model
        step = Stage2Step(args, model, ddp, optimizer, crit, mixup)
        lay = cp.compile_layout({k: v.detach().cpu() for k, v in model.state_dict().items()}, H)
        rho = float(cp.macs(lay)["budget_ratio"])
        info["rho"] = round(rho, 4)
        info["step_flops_per_image"] = (3 * rho + 1) * DENSE_FWD_FLOPS[cfg["model"]]   # live sub-network fwd + bwd (3 rho F) + dense teacher fwd (F)
    return step, model, info


def run_gpu(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from uvc_b200 import _lib
    cfg = BENCH_CONFIGS[a.config]
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else "single rank: not bound"
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = _lib.load()
    np.random.seed(730)
    B = cfg["batch"]
    step, model, info = build_gpu_step(cfg, device, world)
    g = torch.Generator().manual_seed(730 + rank)
    x_host = [torch.randn(B, 3, 224, 224, generator=g).pin_memory() for _ in range(2)]
    y_host = [torch.randint(0, 1000, (B,), generator=g).pin_memory() for _ in range(2)]
    x_dev, y_dev = x_host[0].to(device), y_host[0].to(device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n, fin=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        if fin is not None:
            fin()                               # host work that belongs to the last step (its result read-back)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident arm (inputs already in HBM; the step's own clone + mixup happen inside)
    def dev_step(i):
        step(x_dev.clone(), y_dev)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for i in range(a.warmup):
        dev_step(i)
    if getattr(step, "_timing", None) is not None:
        step._timing = []          # UVC_STEP_TIMING: report the timed steps only
    n0 = lib.uvc_launch_count()
    w0 = time.time()
    ms = timed(dev_step, a.steps)
    w1 = time.time()
    if os.environ.get("UVC_STEP_TIMING") and rank == 0 and hasattr(step, "timing_report"):
        print(step.timing_report(), file=sys.stderr); step._timing = None
    launches = int(lib.uvc_launch_count() - n0)

    # ---- end-to-end arm: pinned host -> device every step (double-buffered on a copy stream), results read back every step
    copy_stream = torch.cuda.Stream(device)
    bufs = [(torch.empty_like(x_dev), torch.empty_like(y_dev)) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    d2h = [0]
    done = [None, None]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            if done[i % 2] is not None:
                copy_stream.wait_event(done[i % 2])        # the step that last consumed this buffer has finished
            bx, by = bufs[i % 2]
            bx.copy_(x_host[i % 2], non_blocking=True); by.copy_(y_host[i % 2], non_blocking=True)
            ready[i % 2].record(copy_stream)

    def e2e_step(i):
        if i == 0:
            prefetch(0)
        prefetch(i + 1)                         # next batch's H2D copy overlaps this step's compute
        cur = torch.cuda.current_stream()
        cur.wait_event(ready[i % 2])
        bx, by = bufs[i % 2]
        out = step(bx, by)                      # consumes bx in place (mixup)
        done[i % 2] = torch.cuda.Event(); done[i % 2].record(cur)
        # device -> host read of EVERY step's result (loss + the ADMM state the reference API returns), issued now as an asynchronous copy into
        # pinned memory and consumed one step later, the way a training loop logs without stalling the launch queue
        lh = loss_host[i % 2]
        lh.copy_(out["loss"].detach().reshape(1), non_blocking=True)
        ev = torch.cuda.Event(); ev.record(cur)
        resolve()
        pending.append((ev, lh, out))

    loss_host = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    pending, losses = [], []

    def resolve():
        while pending:
            ev, lh, out = pending.pop(0)
            ev.synchronize()
            losses.append(float(lh.item()))
            n = 1
            for k in ("s", "r", "gating"):      # resolves the deferred ADMM read-back (Stage 1)
                if out.get(k) is not None:
                    n += out[k].size
            d2h[0] = 4 + (4 * n if "s" in out else 0)
    for i in range(min(3, a.warmup)):
        e2e_step(i)
    resolve()
    w2 = time.time()
    ms_e2e = timed(e2e_step, a.steps, fin=resolve)
    w3 = time.time()
    clk = clocks.stop([(w0, w1, "timed region"), (w2, w3, "end-to-end leg (timed region too short for a sample)")]) if rank == 0 else None

    # ---- roofline leg: every GEMM launch of 2 instrumented steps timed with a CUDA-event pair on the launching stream
    # (the teacher forward normally runs on a side stream next to the student forward; for this leg it is put back in line so that every
    # launch is timed alone on the GPU -- a kernel that shares the SMs with another stream's kernel is not a roofline measurement)
    prev_ts = os.environ.get("UVC_TEACHER_STREAM")
    os.environ["UVC_TEACHER_STREAM"] = "0"
    lib.uvc_gemm_profile(1)
    for i in range(2):
        dev_step(i)
    torch.cuda.synchronize()
    if prev_ts is None:
        os.environ.pop("UVC_TEACHER_STREAM", None)
    else:
        os.environ["UVC_TEACHER_STREAM"] = prev_ts
    import ctypes
    f16_mode = bool(model._dims(B).operand_f16)
    t_ms, t_fl, n_l = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
    lib.uvc_gemm_profile_read_kind(3 if f16_mode else 2, ctypes.byref(t_ms), ctypes.byref(t_fl), ctypes.byref(n_l))   # the dominant kernel: persistent CTA-pair GEMM (kind 3: fp16 operands)
    a_ms, a_fl, a_n = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
    lib.uvc_gemm_profile_read_kind(0, ctypes.byref(a_ms), ctypes.byref(a_fl), ctypes.byref(a_n))       # every GEMM launch (both kernels)
    lib.uvc_gemm_profile(0)
    barrier()
    # ---- chip peaks measured in THIS run the way MEASURED_PEAKS.json does (torch.matmul 8192^3, best of 10): the operand format the kernels use
    live = measure_live_peaks(device) if (rank == 0 and not a.no_live_peaks) else None
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "gemm2_traffic.json")) as f:     # dram bytes per launch of the same kernel, from an ncu capture of this bench
            traffic = json.load(f).get(f"{a.config}:{'f16' if f16_mode else 'tf32'}")      # per configuration and precision mode; None if never captured
    except Exception:
        pass
    imgs = B * world * a.steps
    value = imgs / (ms / 1e3)
    step_flops = info["step_flops_per_image"] * B
    gemm_tflops = (t_fl.value / max(t_ms.value, 1e-9)) / 1e9
    tf32_peak = peaks["bf16_sustained"] / 2.0
    mma_peak = peaks["bf16_sustained"] if f16_mode else tf32_peak       # the peak of the operand format the kernel actually uses
    kind_txt = "kind::f16 (fp16 operand storage, fp32 accumulate)" if f16_mode else "kind::tf32"
    out = {
        "metric": cfg["metric"], "value": round(value, 1), "unit": "images/sec",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(ms / a.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": ("f16 operands / f32 accumulate (GEMM + attention operands stored as fp16 = TF32's 10 mantissa bits; residual stream, LayerNorm, softmax, loss, "
                  "optimizer, ADMM in f32)" if f16_mode else "tf32 (fp32 storage, fp32 accumulate)"), "data": "synthetic",
        "config": {"workload": cfg["workload"], "name": a.config, "per_gpu_batch": B, "global_batch": B * world,
                   "parallelism": f"dp{world}", "host_numa_binding": numa, "l2": "inputs + activations per step (>8 GB) far exceed the 126 MB L2; no explicit flush",
                   "parity_unpinned": "Mixup / soft-target CE / AdamW grouping follow timm's public semantics (timm is absent from the reference tree)",
                   **({"teacher_forward": ("in line" if os.environ.get("UVC_TEACHER_STREAM", "1") == "0" else "on a side stream next to the student forward")}
                      if cfg["stage"] != "eval" else {}), **info},
        "clocks": clk,
        "e2e": {"value": round(imgs / (ms_e2e / 1e3), 1), "unit": "images/sec", "ms_per_step": round(ms_e2e / a.steps, 3),
                "h2d_bytes_per_step": int(x_host[0].numel() * 4 + y_host[0].numel() * 8), "d2h_bytes_per_step": int(d2h[0]),
                "readback": "every step's loss (and, in Stage 1, the ADMM state) is copied to pinned host memory asynchronously and consumed one step later; the "
                            "last step's read-back is inside the timed region", "last_loss": round(losses[-1], 4) if losses else None},
        "gpu_launches": launches,
        "step_tflops_per_gpu": round(step_flops / (ms / a.steps / 1e3) / 1e12, 1),
        "roofline": {"bound": "tensor", "kernel": f"uvc::gemm2_tf32_kernel (persistent CTA pairs, tcgen05.mma cta_group::2 {kind_txt})",
                     "achieved": round(gemm_tflops, 1), "peak": round(mma_peak, 1), "unit": "TFLOP/s", "frac": round(gemm_tflops / mma_peak, 3),
                     "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                     "traffic_source": traffic["source"] if traffic else None,
                     "algorithmic_flops_per_launch": round(t_fl.value / max(1, n_l.value)),
                     "avg_launch_us": round(t_ms.value * 1e3 / max(1, n_l.value), 2),
                     "peak_source": (f"bf16_tflops_sustained of MEASURED_PEAKS.json ({peaks['src']}): 16-bit dense, the format the kernel's MMAs run in" if f16_mode else
                                     f"TF32 dense = 1/2 x bf16_tflops_sustained of MEASURED_PEAKS.json ({peaks['src']}); bf16 sustained {peaks['bf16_sustained']}"),
                     "live_peaks": live,
                     "frac_of_live_peak": (round(gemm_tflops / live["f16_tflops" if f16_mode else "tf32_tflops"], 3) if live else None),
                     "launches_per_step": int(n_l.value // 2), "kernel_ms_per_step": round(t_ms.value / 2, 3),
                     "all_gemm_launches_per_step": int(a_n.value // 2), "all_gemm_ms_per_step": round(a_ms.value / 2, 3),
                     "all_gemm_tflops": round((a_fl.value / max(a_ms.value, 1e-9)) / 1e9, 1),
                     "how": "CUDA-event pair on the launching stream around every GEMM launch of 2 instrumented steps run right after the timed region; "
                            "achieved = sum of 2*M*N*K over the CTA-pair kernel's launches / sum of their durations",
                     "step_frac_of_peak": round(step_flops / (ms / a.steps / 1e3) / 1e12 / mma_peak, 3)},
    }
    out["cpu_baseline"] = run_cpu_sample(a.config, steps=2, warmup=1) if world == 1 and not a.no_cpu_baseline else None
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_live_peaks(device):
    """The tensor-pipe peaks of THIS chip in THIS run, measured as MEASURED_PEAKS.json's `how` says (torch.matmul 8192^3, 2*N^3 FLOPs, best of 10, CUDA events):
    fp16 and TF32 operands.  Library GEMMs are used here ONLY as the yardstick the hand-written kernels are held against."""
    import torch
    res = {}
    n = 8192
    for name, dt, tf32 in (("f16_tflops", torch.float16, False), ("tf32_tflops", torch.float32, True)):
        try:
            old = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            a_ = torch.randn(n, n, device=device, dtype=dt); b_ = torch.randn(n, n, device=device, dtype=dt)
            for _ in range(3):
                torch.matmul(a_, b_)
            best = 1e9
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); torch.matmul(a_, b_); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            res[name] = round(2.0 * n ** 3 / (best * 1e-3) / 1e12, 1)
            torch.backends.cuda.matmul.allow_tf32 = old
            del a_, b_
        except Exception as e:          # never let the yardstick break the bench line
            res[name] = None
            res[name + "_error"] = str(e)[:80]
    res["how"] = "torch.matmul 8192^3 (cuBLAS), best of 10, CUDA events, in this process right after the timed legs"
    return res


# ------------------------------------------------------------------------------------------------ the CPU (reference) arm
def build_cpu_step(config, B):
    """The same step as the oracle states it: plain PyTorch fp32 on the host cores (oracle/vit_oracle.py, oracle/admm_oracle.py).
    TEST/BASELINE infrastructure: this is the thing being compared against, never shipped."""
    import numpy as np
    import torch
    from oracle import admm_oracle as ao, fixtures as fx, vit_oracle as vo
    cfg = BENCH_CONFIGS[config]
    mt = cfg["model"]
    dims = fx.MODEL_DIMS[mt]
    C, H, L = dims["embed_dim"], dims["num_heads"], dims["depth"]
    Fh = int(C * dims.get("mlp_ratio", 4))
    t2t = bool(dims.get("t2t"))
    eps = 1e-5 if t2t else 1e-6
    sd, _ = fx.make_state_dict(mt, None, seed=730)
    params = {k: v.clone().requires_grad_(v.is_floating_point() and k != "pos_embed" or (k == "pos_embed" and not t2t)) for k, v in sd.items()}
    teacher = {k: v.clone() for k, v in sd.items()}
    if "gumbel.weight" not in params:       # the token-gate scorer (T2T: the reference's T2T_ViT has none; cfg5 applies the DeiT gate after tokens_to_token)
        params["gumbel.weight"] = (torch.randn(1, C, generator=torch.Generator().manual_seed(1)) * 0.02).requires_grad_(True)
        params["gumbel.bias"] = torch.zeros(1, requires_grad=True)
    names = [k for k in params if params[k].requires_grad and k not in ("gumbel.weight", "gumbel.bias") and "attn_skip" not in k and "mlp_skip" not in k]
    ms = {k: torch.zeros_like(params[k]) for k in names}; vs = {k: torch.zeros_like(params[k]) for k in names}
    x0, y0 = fx.make_batch(B, seed=730)
    state = {"step": 0}

    def fwd(p, x, **kw):
        if t2t:
            tok, _ = vo.t2t_tokens(p, x)
            return vo.forward(p, x, L, H, eps=eps, tokens=tok, **kw), tok
        return vo.forward(p, x, L, H, **kw), None

    if cfg["stage"] == "eval":      # validation loop body: eval forward + CE + top-1 (joint_train.py:199-246 / post_train.py:209-265)
        skip_e = [cfg.get("layout") and l in (8, 10) for l in range(L)]
        if cfg.get("layout"):
            W1 = [params[f"blocks.{i}.attn.proj.weight"].detach() for i in range(L)]
            W3 = [params[f"blocks.{i}.mlp.fc2.weight"].detach() for i in range(L)]
            m1, m3 = ao.masks(W1, W3, torch.tensor([[3.0, 1417.0]] * L), torch.full((L, H), 16.0), 64, Fh)
            with torch.no_grad():
                for i in range(L):
                    params[f"blocks.{i}.attn.proj.weight"].mul_(m1[i].float().unsqueeze(0))
                    params[f"blocks.{i}.mlp.fc2.weight"].mul_(m3[i].float().unsqueeze(0))
                    params[f"blocks.{i}.mlp.fc1.weight"].mul_(m3[i].float().unsqueeze(1))

        def step_eval():
            with torch.no_grad():
                logits = vo.forward(params, x0, L, H, skip=skip_e)
                loss = torch.nn.functional.cross_entropy(logits, y0)
                _ = (logits.argmax(1) == y0).sum()
            return float(loss)
        return step_eval

    if cfg["stage"] == 1:
        patch_gating = (3 * torch.ones(196)).requires_grad_(True)
        st = dict(s=torch.stack([torch.full((L,), 1.3), torch.full((L,), 0.26 * Fh + 0.5)], 1), r=torch.full((L, H), 9.2), y=torch.full((L, 2), 1e-3),
                  p=torch.full((L, H), 1e-3), z=torch.tensor(1e-3), gate=params["block_skip_gating"], gate_buf=[])
        macs = torch.tensor([vo.block_macs(1, 197, C, H, Fh)] * L, dtype=torch.float32)
        embed = 196 * C * 768
        hp = dict(lr=1e-4, slr=0.02, rlr=0.02, ylr=1e-4, plr=1e-4, zlr=1.0, budget=cfg["budget"], z_grad_clip=0.5, sl2wd=0.0, gating_weight=5e-4, d=64, Fh=Fh,
                  macs=macs, embed_macs=embed, full=float((embed + macs.sum()) * 2), use_gumbel=True, eps=0.1, gating_interval=50, global_step=0)

        def step():
            x = x0.clone()
            lam = float(np.random.beta(0.8, 0.8))
            x = x * lam + x.flip(0) * (1 - lam)
            tgt = vo.mixup_target(y0, 1000, lam, 0.1)
            blend = torch.stack([torch.nn.functional.gumbel_softmax(params["block_skip_gating"][i], tau=0.5, hard=False, eps=1e-10, dim=-1) for i in range(L)])
            if t2t:       # token slimming (mode 2): top 90 % by Gumbel-perturbed score
                tok, _ = vo.t2t_tokens(params, x)
                noise = -torch.empty(B, 196).exponential_().log()
                tmask, _ = vo.token_gate(params, tok, noise, 1.0, int(0.9 * 196))
                logits = vo.forward(params, x, L, H, eps=eps, blend=blend, token_mask=tmask, tokens=tok)
            else:
                logits = vo.forward(params, x, L, H, blend=blend, patch_scale=torch.sigmoid(patch_gating))
            with torch.no_grad():
                t_logits, _ = fwd(teacher, x, skip=[False] * L)
            loss, _, _ = vo.distillation_loss(logits, t_logits, tgt, 0.1, 1.0)
            for p in params.values():
                p.grad = None
            loss.backward()
            state["step"] += 1
            with torch.no_grad():
                live = [k for k in names if params[k].grad is not None]
                vo.clip_adamw_step([params[k] for k in live], [params[k].grad for k in live], [ms[k] for k in live], [vs[k] for k in live], state["step"], 1e-4)
                W1 = [params[f"blocks.{i}.attn.proj.weight"] for i in range(L)]
                W3 = [params[f"blocks.{i}.mlp.fc2.weight"] for i in range(L)]
                n1, n2 = [-torch.empty(L, 2).exponential_().log() for _ in range(2)]
                hp["global_step"] = state["step"]
                ao.step(st, W1, W3, hp, n1, n2, gate_grad=params["block_skip_gating"].grad, gate_sgd=lambda g: None)
            return float(loss)
        return step

    # Stage 2: fixed layout, masked weights, hard skip, clip + AdamW with timm's decay grouping, re-mask (post_train.py:351-383)
    skip = [l in (8, 10) for l in range(L)]
    W1 = [params[f"blocks.{i}.attn.proj.weight"].detach() for i in range(L)]
    W3 = [params[f"blocks.{i}.mlp.fc2.weight"].detach() for i in range(L)]
    m1, m3 = ao.masks(W1, W3, torch.tensor([[3.0, 1417.0]] * L), torch.full((L, H), 16.0), 64, Fh)
    masks = {}
    for i in range(L):
        masks[f"blocks.{i}.attn.proj.weight"] = m1[i].float().unsqueeze(0).expand(C, C)
        masks[f"blocks.{i}.mlp.fc2.weight"] = m3[i].float().unsqueeze(0).expand(C, Fh)
        masks[f"blocks.{i}.mlp.fc1.weight"] = m3[i].float().unsqueeze(1).expand(Fh, C)
    with torch.no_grad():
        for k, mk in masks.items():
            params[k].mul_(mk)

    def step2():
        x = x0.clone()
        lam = float(np.random.beta(0.8, 0.8))
        x = x * lam + x.flip(0) * (1 - lam)
        tgt = vo.mixup_target(y0, 1000, lam, 0.1)
        logits = vo.forward(params, x, L, H, skip=skip)
        with torch.no_grad():
            t_logits = vo.forward(teacher, x, L, H, skip=[False] * L)
        loss, _, _ = vo.distillation_loss(logits, t_logits, tgt, 0.1, 1.0)
        for p in params.values():
            p.grad = None
        loss.backward()
        state["step"] += 1
        with torch.no_grad():
            live = [k for k in names if params[k].grad is not None and k != "block_skip_gating"]
            vo.clip_adamw_step([params[k] for k in live], [params[k].grad for k in live], [ms[k] for k in live], [vs[k] for k in live], state["step"], 5e-4)
            for k, mk in masks.items():
                params[k].mul_(mk)
        return float(loss)
    return step2


CPU_SAMPLE_BATCH = {"small_s1": 16, "tiny_s1": 32, "base_s2": 8, "t2t_s1": 8, "small_eval": 32, "base_s2_eval": 16}


def run_cpu_sample(config, steps, warmup):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = CPU_SAMPLE_BATCH[config]
    step = build_cpu_step(config, B)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    what = {1: "student fwd+bwd, dense teacher fwd, CE+KD loss, clip+AdamW, ADMM step",
            2: "masked student fwd+bwd with blocks 8/10 skipped, dense teacher fwd, CE+KD loss, clip+AdamW, re-mask",
            "eval": "eval forward (masked-dense where the config has a layout), CrossEntropyLoss, top-1"}[BENCH_CONFIGS[config]["stage"]]
    return {"value": round(B * steps / dt, 2), "unit": "images/sec", "cores": cores, "kind": "port", "batch": B,
            "sample": f"{steps} full steps ({what}) of {BENCH_CONFIGS[config]['model']} at batch {B} "
                      f"through the oracle port (plain PyTorch fp32, {cores} threads) after {warmup} warm-up"}


def run_reference(a):
    """`--impl reference`: the reference's CPU path = the oracle port (the reference itself is Python that cannot travel to the
    GPU box; oracle/ is pinned to it bit-for-bit by tests/test_oracle.py and oracle/gen_golden*.py).  Rank 0 only."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    cfg = BENCH_CONFIGS[a.config]
    steps, warm = max(1, min(a.steps, 3)), max(1, min(a.warmup, 1))
    cb = run_cpu_sample(a.config, steps, warm)
    world = int(os.environ.get("WORLD_SIZE", 1))
    out = {"impl": "reference", "metric": cfg["metric"], "value": cb["value"], "unit": "images/sec",
           "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": round(cb["batch"] / cb["value"] * 1e3, 1), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": cfg["workload"] + f" -- on the host CPU: bounded sample of {cb['batch']} images per step (same step, same model)", "name": a.config},
           "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="uvc_b200", choices=["uvc_b200", "reference"])
    ap.add_argument("--config", default="small_s1", choices=sorted(BENCH_CONFIGS), help="which BASELINE.json configuration (default: configs[2], the headline)")
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--no-live-peaks", dest="no_live_peaks", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl != "reference" else a.warmup
    # stdout carries the ONE JSON line and nothing else: libraries that write to file descriptor 1 themselves (NCCL prints its version banner there,
    # the reference's own prints, warnings of C extensions) are sent to stderr for the whole run; the JSON line goes to the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if a.impl == "reference":
        return run_reference(a)
    run_gpu(a)


if __name__ == "__main__":
    main()
