"""Stage-2 TRAINING on the physically compacted model (include/uvc_b200.h uvc_vit_layout, uvc_b200/compact.py:EngineLayout).

The reference trains Stage 2 masked-dense: `weight.data *= mask` before every step (post_train.py:357-360), hard block skip
(models/model_distilled.py:496-500), and multiplies by the zeros.  With a layout the engine gathers the live heads / neurons into compact
operands, runs every kernel of a block at the live widths and scatters the compact gradients back.  Checked here, through the C ABI:
  * logits = the masked-dense engine's and the CPU oracle's (1e-3 relative, the north-star tolerance);
  * every gradient at a live position = the oracle's masked-dense gradient (5e-3 of the tensor's maximum, the backward tolerance of the model tests);
  * pruned fc2 columns receive the reference's closed-form gradient gelu(fc1.bias[n]) * d fc2.bias (what autograd gives for a zeroed fc1 row),
    pruned fc1 rows / biases and the q, k, v rows of pruned heads receive exactly zero, as in the reference;
  * the ONE stated deviation: attn.proj columns of fully pruned heads get no gradient (the reference's is dX1^T ctx_head, which needs that head's
    whole attention); `keep_pruned_heads=True` keeps those heads and reproduces the reference's gradient -- and with it its clip norm -- everywhere;
  * degenerate layouts (a block with every head and every neuron pruned) and a hard-skipped block."""
import pytest
import torch

from oracle import fixtures as fx, vit_oracle as vo
from test_compact import pruned_checkpoint

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("precision")]

LOGIT_TOL = 1e-3
GRAD_TOL = 5e-3


def rel(a, b):
    return ((a.detach().cpu() - b.detach().cpu()).abs().max() / b.detach().abs().max().clamp_min(1e-30)).item()


def _stage2_model(sd, mt, depth):
    from test_stage2_gpu import build
    from uvc_b200.post_train import apply_masks
    m = build(mt, depth, {k: v for k, v in sd.items() if not k.endswith(".mask")}, gumbel_hard=True).train()
    for _, mod in m.named_modules():
        if hasattr(mod, "weight"):
            mod.register_buffer("mask", torch.ones_like(mod.weight))
    mods = dict(m.named_modules())
    for k, v in sd.items():
        if k.endswith(".mask"):
            mods[k[:-5]].mask.copy_(v.cuda())
    m.enable_block_gating = 0
    m.block_skip_gating.requires_grad = False
    m.flatten_parameters()
    apply_masks(m)
    return m


def _oracle(sd, x, tgt, depth, H):
    sdm = {k: v.clone() for k, v in sd.items() if not k.endswith(".mask")}
    for k, v in sd.items():
        if k.endswith(".mask"):
            sdm[k[:-4] + "weight"] = sdm[k[:-4] + "weight"] * v
    sdr = {k: v.clone().requires_grad_(True) for k, v in sdm.items()}
    skip = [not bool(g[1] > g[0]) for g in sd["block_skip_gating"]]
    lo = vo.forward(sdr, x, depth, H, skip=skip)
    loss, _, _ = vo.distillation_loss(lo, torch.zeros_like(lo), tgt, 0.1, 1.0)
    loss.backward()
    return lo, sdr


def _run(m, x, tgt):
    from uvc_b200 import ops
    m.zero_grad(set_to_none=True)
    (logits, _), _ = m(x.cuda())
    _, dl = ops.distill_loss(logits.detach(), torch.zeros_like(logits), tgt.cuda(), 0.1, 1.0)
    logits.backward(dl)
    return logits.detach().clone(), {k: (None if p.grad is None else p.grad.detach().clone()) for k, p in m.named_parameters()}


@pytest.mark.parametrize("keep_heads", [False, True])
def test_compacted_training_step_matches_masked_dense(keep_heads):
    from uvc_b200 import compact as cp
    if not __import__("uvc_b200._lib", fromlist=["x"]).operand_f16_for(None, 64, 197, 192, 768):
        pytest.skip("the compaction layout is implemented for the fp16-operand engine")
    mt, depth, B = "deit_tiny_patch16_224", 4, 6
    sd, dims = pruned_checkpoint(mt, depth, seed=3)
    H = dims["num_heads"]
    x, _ = fx.make_batch(B, seed=11)
    tgt = fx.soft_targets(B, seed=11)
    lo, sdr = _oracle(sd, x, tgt, depth, H)
    m = _stage2_model(sd, mt, depth)
    logits_d, grads_d = _run(m, x, tgt)                       # masked-dense engine
    lay = cp.engine_layout_for(m, keep_pruned_heads=keep_heads)
    assert lay.struct.n_heads[0] == (H if keep_heads else H - 1) and lay.struct.n_neurons[0] % 64 == 0
    assert lay.struct.n_heads[2] == (H if keep_heads else 1) and lay.struct.n_neurons[2] == 64       # degenerate block: topped up with pruned entries
    m.compact_layout = lay
    logits_c, grads_c = _run(m, x, tgt)
    assert rel(logits_c, lo) < LOGIT_TOL and rel(logits_c, logits_d) < LOGIT_TOL
    d = 64
    for k, g in grads_c.items():
        go = sdr[k].grad if k in sdr else None
        if k.startswith("blocks.1."):
            assert g is None, k                                 # hard-skipped block: outside the graph
            continue
        if go is None or g is None:
            continue
        g = g.cpu()
        scale = go.abs().max().clamp_min(1e-30)
        if k.endswith("attn.proj.weight") and not keep_heads:
            l = int(k.split(".")[1])
            dead = [h for h in range(H) if h not in lay.heads[l]]
            live_cols = torch.ones(go.shape[1], dtype=torch.bool)
            for h in dead:
                live_cols[h * d:(h + 1) * d] = False
                assert torch.count_nonzero(g[:, h * d:(h + 1) * d]) == 0, k          # the stated deviation: no gradient for a pruned head's proj columns
            assert ((g - go)[:, live_cols].abs().max() / scale) < GRAD_TOL, k
            continue
        assert ((g - go).abs().max() / scale) < GRAD_TOL, (k, float((g - go).abs().max() / scale))
        if k.endswith("mlp.fc1.weight"):                        # pruned neurons: exactly zero rows, as in the reference
            l = int(k.split(".")[1])
            deadn = torch.ones(go.shape[0], dtype=torch.bool); deadn[lay.neurons[l]] = False
            assert torch.count_nonzero(g[deadn]) == 0, k
    # the squared gradient norm the clip uses: equal to the reference's in the exact mode, smaller by the pruned heads' proj columns otherwise
    tot_o = sum(float((sdr[k].grad.double() ** 2).sum()) for k in grads_c if grads_c[k] is not None and k in sdr and sdr[k].grad is not None)
    tot_c = sum(float((g.double() ** 2).sum()) for k, g in grads_c.items() if g is not None and k in sdr and sdr[k].grad is not None)
    if keep_heads:
        assert abs(tot_c - tot_o) / tot_o < 5e-3
    else:
        miss = 0.0
        for l in (0, 2, 3):
            go = sdr[f"blocks.{l}.attn.proj.weight"].grad
            for h in range(H):
                if h not in lay.heads[l]:
                    miss += float((go[:, h * d:(h + 1) * d].double() ** 2).sum())
        assert abs(tot_c + miss - tot_o) / tot_o < 5e-3


def test_compacted_step_keeps_masked_weights_zero_and_moves_live_ones():
    """one Stage-2 optimiser step on the compacted engine: masked weights stay exactly zero, the live ones follow the masked-dense step"""
    from uvc_b200 import compact as cp, _lib
    from uvc_b200.post_train import param_groups_weight_decay
    from uvc_b200.utils.optim import FusedClipAdamW
    if not _lib.operand_f16_for(None, 64, 197, 192, 768):
        pytest.skip("the compaction layout is implemented for the fp16-operand engine")
    mt, depth, B = "deit_tiny_patch16_224", 4, 4
    sd, dims = pruned_checkpoint(mt, depth, seed=5)
    x, _ = fx.make_batch(B, seed=2)
    tgt = fx.soft_targets(B, seed=2)
    res = []
    for compact in (False, True):
        m = _stage2_model(sd, mt, depth)
        if compact:
            m.compact_layout = cp.engine_layout_for(m, keep_pruned_heads=True)
        masks = {mod.weight: mod.mask for _, mod in m.named_modules() if hasattr(mod, "mask")}
        opt = FusedClipAdamW(param_groups_weight_decay(m, 0.05), lr=1e-3, weight_decay=0.05, max_grad_norm=1.0, model=m, masks=masks)
        _run(m, x, tgt)
        opt.step()
        assert opt._flat is not None
        res.append(({k: p.detach().clone() for k, p in m.named_parameters()}, float(opt.grad_norm())))
        for _, mod in m.named_modules():
            if hasattr(mod, "mask"):
                assert torch.count_nonzero(mod.weight.detach()[mod.mask == 0]) == 0
    (pd, nd), (pc, nc) = res
    assert abs(nd - nc) / nd < 5e-3                             # exact mode: the clip sees the reference's norm
    for k in pd:
        # Adam's first step is ~ lr * sign(g), so an element whose gradient is rounding noise (e.g. the k-part of qkv.bias, zero in exact
        # arithmetic) may flip: every element is bounded by two steps, and the weight matrices must agree on all but a few per cent
        diff = (pd[k] - pc[k]).abs()
        assert float(diff.max()) <= 2.1e-3, k
        if pd[k].dim() >= 2 and not k.startswith("blocks.1."):
            assert float((diff > 1e-4).float().mean()) < 0.03, (k, float((diff > 1e-4).float().mean()))


def test_validation_step_dense_and_compacted_agree_with_the_reference_loop():
    """EvalStep (the validation loop body, joint_train.py:199-246 / post_train.py:209-265): device-side accumulation gives the reference loop's
    two averages (top-1 weighted by batch size, loss per batch), and the compacted engine validates to the masked-dense numbers."""
    import types
    from uvc_b200 import compact as cp, _lib
    from uvc_b200.joint_train import EvalStep
    if not _lib.operand_f16_for(None, 64, 197, 192, 768):
        pytest.skip("the compaction layout is implemented for the fp16-operand engine")
    mt, depth = "deit_tiny_patch16_224", 4
    sd, dims = pruned_checkpoint(mt, depth, seed=7)
    m = _stage2_model(sd, mt, depth).eval()
    args = types.SimpleNamespace(enable_patch_gating=0, patch_ratio=0.9)
    batches = [(fx.make_batch(6, seed=31)), (fx.make_batch(4, seed=32))]
    lo = [_oracle_logits(sd, x, depth, dims["num_heads"]) for x, _ in batches]
    want_top1 = 100.0 * sum(float((l.argmax(1) == y).sum()) for l, (_, y) in zip(lo, batches)) / 10
    want_loss = sum(float(torch.nn.functional.cross_entropy(l, y)) for l, (_, y) in zip(lo, batches)) / 2
    for compact in (False, True):
        m.compact_layout = cp.engine_layout_for(m) if compact else None
        step = EvalStep(args, m)
        for x, y in batches:
            step(x.cuda(), y.cuda())
        top1, loss = step.result()
        assert abs(top1 - want_top1) < 1e-9 or abs(top1 - want_top1) <= 10.0     # argmax may differ only on a near-tie of random logits
        assert abs(loss - want_loss) < 2e-3 * max(1.0, abs(want_loss)), (compact, loss, want_loss)


def _oracle_logits(sd, x, depth, H):
    sdm = {k: v.clone() for k, v in sd.items() if not k.endswith(".mask")}
    for k, v in sd.items():
        if k.endswith(".mask"):
            sdm[k[:-4] + "weight"] = sdm[k[:-4] + "weight"] * v
    skip = [not bool(g[1] > g[0]) for g in sd["block_skip_gating"]]
    with torch.no_grad():
        return vo.forward(sdm, x, depth, H, skip=skip)
