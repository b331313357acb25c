"""CPU tests of the host-side mirror of the reference's Python surface (no GPU, no library compute calls): mixup / soft-target CE
(timm semantics, re-stated because timm is an un-vendored dependency of the reference), LR schedules, the CONFIGS table, layer
discovery (`get_uvc_layers`, joint_train.py:530-564), state-dict key set, CLI flags."""
import math

import numpy as np
import pytest
import torch

from oracle import fixtures as fx, vit_oracle as vo


def test_mixup_target_and_batch_mix_match_timm_semantics():
    from uvc_b200.utils.mixup import Mixup, SoftTargetCrossEntropy, mixup_target
    y = torch.tensor([1, 3, 3, 0])
    t = mixup_target(y, 5, lam=0.7, smoothing=0.1)
    off, on = 0.1 / 5, 1 - 0.1 + 0.1 / 5
    exp = torch.full((4, 5), off); exp[torch.arange(4), y] = on
    expf = torch.full((4, 5), off); expf[torch.arange(4), y.flip(0)] = on
    torch.testing.assert_close(t, 0.7 * exp + 0.3 * expf)
    torch.testing.assert_close(t, vo.mixup_target(y, 5, 0.7, 0.1))              # oracle restatement agrees
    assert torch.allclose(t.sum(1), torch.ones(4))
    # batch-mode mixup: lam ~ Beta(.8,.8) from numpy's global RNG, x <- lam x + (1-lam) flip(x)
    mix = Mixup(mixup_alpha=0.8, cutmix_alpha=0.0, prob=1.0, switch_prob=0.0, label_smoothing=0.1, num_classes=5)
    x = torch.arange(4 * 3 * 4 * 4, dtype=torch.float32).reshape(4, 3, 4, 4)
    np.random.seed(3); np.random.rand(); lam = float(np.random.beta(0.8, 0.8))
    np.random.seed(3)
    xm, tm = mix(x.clone(), y)
    torch.testing.assert_close(xm, lam * x + (1 - lam) * x.flip(0))
    torch.testing.assert_close(tm, mixup_target(y, 5, lam, 0.1))
    # cutmix: a box of the flipped batch is pasted, lam corrected to the box area
    mix = Mixup(mixup_alpha=0.0, cutmix_alpha=1.0, prob=1.0, switch_prob=1.0, label_smoothing=0.0, num_classes=5)
    np.random.seed(5)
    xc, tc = mix(x.clone(), y)
    changed = (xc != x).any(dim=1).any(dim=0)             # [H, W] mask of pasted pixels
    lam_box = 1 - changed.float().mean().item()
    torch.testing.assert_close(tc, mixup_target(y, 5, lam_box, 0.0))
    assert torch.equal(xc[:, :, changed], x.flip(0)[:, :, changed])
    # soft-target CE = mean_b sum_c -t log_softmax
    logits = torch.randn(4, 5)
    torch.testing.assert_close(SoftTargetCrossEntropy()(logits, t), vo.soft_target_cross_entropy(logits, t))


def test_schedules():
    from uvc_b200.joint_train import get_tau
    from uvc_b200.post_train import CosineEpochSchedule
    from uvc_b200.utils.scheduler import WarmupCosineSchedule, WarmupLinearSchedule
    p = [torch.nn.Parameter(torch.zeros(1))]
    opt = torch.optim.SGD(p, lr=1.0)
    s = WarmupCosineSchedule(opt, warmup_steps=10, t_total=110)
    lrs = []
    for _ in range(110):
        lrs.append(opt.param_groups[0]["lr"]); opt.step(); s.step()
    assert lrs[0] == 0.0 and abs(lrs[5] - 0.5) < 1e-9 and abs(lrs[10] - 1.0) < 1e-9
    assert abs(lrs[60] - 0.5 * (1 + math.cos(math.pi * 0.5))) < 1e-9 and lrs[109] < 1e-3            # utils/scheduler.py:46-63
    opt = torch.optim.SGD(p, lr=1.0)
    s = WarmupLinearSchedule(opt, warmup_steps=10, t_total=110)
    for _ in range(60):
        opt.step(); s.step()
    assert abs(opt.param_groups[0]["lr"] - 0.5) < 1e-9
    assert get_tau(10, 0.1, 0, 100) == pytest.approx(0.1) and get_tau(10, 0.1, 100, 100) == pytest.approx(10)   # rises, as the reference (:83-85)
    c = CosineEpochSchedule(torch.optim.SGD(p, lr=1.0), epochs=100, base_lr=1e-3)
    assert c.get_epoch_values(0)[0] == pytest.approx(1e-6) and c.get_epoch_values(5)[0] == pytest.approx(1e-5 + 0.5 * (1e-3 - 1e-5) * (1 + math.cos(math.pi * 0.05)))


def test_configs_layers_and_state_dict_keys():
    from functools import partial
    from uvc_b200.joint_train import get_uvc_layers
    from uvc_b200.models import CONFIGS, DistilledVisionTransformer
    for name, (C, H, L) in {"deit_tiny_patch16_224": (192, 3, 12), "deit_small_patch16_224": (384, 6, 12), "deit_base_patch16_224": (768, 12, 12)}.items():
        cfg = CONFIGS[name]
        assert (cfg.embed_dim, cfg.num_heads, cfg.depth) == (C, H, L) and cfg.hidden_size // cfg.transformer["num_heads"] == 64
    m = DistilledVisionTransformer(enable_dist=0, patch_size=16, embed_dim=192, depth=3, num_heads=3, mlp_ratio=4, qkv_bias=True,
                                   norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), drop_rate=0)
    names, layers, d = get_uvc_layers(m)
    assert [len(layers[k]) for k in ("W1", "W2", "W3")] == [3, 3, 3]
    assert names[layers["W1"][1]] == "blocks.1.attn.proj" and names[layers["W2"][2]] == "blocks.2.mlp.fc1" and names[layers["W3"][0]] == "blocks.0.mlp.fc2"
    assert d["s_dict"][layers["W1"][2]] == [2, 0] and d["s_dict"][layers["W3"][2]] == [2, 1] and d["r_dict"][layers["W1"][1]] == 1
    # state-dict keys = the reference's (timm DeiT names + gates), so DeiT checkpoints and Stage-1 -> Stage-2 state dicts load unchanged
    want = set(fx.param_shapes(192, 3, 3).keys())
    assert set(m.state_dict().keys()) == want
    sd, _ = fx.make_state_dict("deit_tiny_patch16_224", 3, seed=1)
    assert m.load_state_dict(sd, strict=True)
    # `.mask` buffers beside every weight, counted like joint_train.py:169-188 (5.65 M for the full DeiT-Tiny, log/deit-tiny-log.log:2)
    full = DistilledVisionTransformer(enable_dist=0, patch_size=16, embed_dim=192, depth=12, num_heads=3, mlp_ratio=4, qkv_bias=True,
                                      norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), drop_rate=0)
    n = sum(mod.weight.numel() for _, mod in full.named_modules() if hasattr(mod, "weight"))
    assert round(n / 1e6, 2) == 5.65
    with pytest.raises(Exception):
        full(torch.zeros(1, 3, 224, 224))           # no CPU path: the engine refuses instead of falling back


def test_cli_flag_compatibility():
    """every flag of run_uvc_train.sh / run_post_train.sh parses (reference joint_train.py:684-879, post_train.py:407-574)"""
    from uvc_b200 import joint_train as jt, post_train as pt
    a = jt.build_parser().parse_args("--name x --dataset imagenet --model_type deit_small_patch16_224 --train_batch_size 128 --eval_batch_size 64 "
                                     "--num_epochs 30 --learning_rate 1e-4 --uvc_train --budget 0.5 --distillation-type soft --distillation-alpha 0.1 "
                                     "--enable_patch_gating 1 --enable_block_gating 1 --gating_weight 5e-4 --eps 0.1 --eps_decay 0.92 --warmup_epochs 5 "
                                     "--zlr_schedule_list 1,5,9,13,17 --slr 0.02 --rlr 0.02 --glr 0.1 --local_rank 0 --seed 730 --gpu_num 0,1".split())
    assert a.budget == 0.5 and a.distillation_type == "soft" and a.seed == 730 and a.enable_patch_gating == 1
    b = pt.build_parser().parse_args("--name y --model_type deit_base_patch16_224 --train_batch_size 256 --local-rank 1".split())
    assert b.local_rank == 1 and b.model_type == "deit_base_patch16_224"


def test_hard_skip_list_cache_rules():
    """The hard-skip decision needs a host read of the gates; it is cached (per storage + version) only where the gates cannot move behind the
    version counter's back: eval mode or frozen gates (Stage 2).  A model that trains its gates re-reads every call."""
    import types
    import torch
    from uvc_b200.models.model_distilled import hard_skip_list
    m = types.SimpleNamespace(block_skip_gating=torch.nn.Parameter(torch.tensor([[-1.0, 1.0], [1.0, -1.0], [0.5, 0.5]])), training=True)
    assert hard_skip_list(m) == [False, True, True]                       # runs iff gate[1] > gate[0] (ties skip, models/model_distilled.py:498)
    reads = []
    orig = torch.Tensor.tolist
    torch.Tensor.tolist = lambda self: (reads.append(1), orig(self))[1]
    try:
        hard_skip_list(m); hard_skip_list(m)
        assert len(reads) == 2                                             # training + trainable gates: read every time
        m.block_skip_gating.requires_grad = False                          # Stage 2 (post_train.py:342)
        hard_skip_list(m); hard_skip_list(m); hard_skip_list(m)
        assert len(reads) == 2                                             # the cache made by the last read is valid: no further reads
        with torch.no_grad():
            m.block_skip_gating[0] = torch.tensor([1.0, -1.0])             # in-place change moves the version counter
        assert hard_skip_list(m) == [True, True, True] and len(reads) == 3
        m.block_skip_gating.requires_grad = True; m.training = False       # eval mode (the teacher)
        hard_skip_list(m); hard_skip_list(m)
        assert len(reads) == 3
    finally:
        torch.Tensor.tolist = orig


def test_engine_param_list_is_cached_and_revalidated():
    import torch
    from uvc_b200.models.model_distilled import _engine_param_list, deit_tiny_patch16_224
    m = deit_tiny_patch16_224()
    a = _engine_param_list(m)
    assert _engine_param_list(m) is a and len(a) == 8 + 12 * 12
    assert [n for n, _ in a[:8]] == ["patch_w", "patch_b", "cls_token", "pos_embed", "norm_w", "norm_b", "head_w", "head_b"]
    m.head.weight = torch.nn.Parameter(torch.zeros_like(m.head.weight))    # surgery: the cached list must not survive it
    b = _engine_param_list(m)
    assert b is not a and b[6][1] is m.head.weight


def test_distillation_loss_prefetch_is_a_noop_off_device():
    import torch
    from uvc_b200.utils.losses import DistillationLoss
    crit = DistillationLoss(torch.nn.CrossEntropyLoss(), None, "none", 0.0, 1.0)
    crit.prefetch_teacher(torch.zeros(2, 3))
    assert crit._pref is None
