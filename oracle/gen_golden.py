"""ORACLE (test infrastructure only) — writes tests/golden/*.pt FROM THE UNMODIFIED REFERENCE.

Run here (the only place /root/reference exists):   python -m oracle.gen_golden
Inputs/weights are regenerated from seeds by oracle/fixtures.py (checksums are stored so a drifting RNG is
detected); only the reference's OUTPUTS are stored, so the fixtures stay small.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fixtures as fx, ref_shim, vit_oracle as vo  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def ref_model(ns, model_type, depth, sd, **kw):
    m = ref_shim.make_ref_model(ns, model_type, depth=depth, **kw)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith("mask") for k in missing), (missing, unexpected)
    return m


def gen_forward_cases(ns):
    """cfg1 (BASELINE.json configs[0]): DeiT-Tiny 1 block, batch 8, eval, masks all-ones, default gates -> logits.
    Plus deeper / gated / skipped variants."""
    cases = {}
    specs = [
        ("cfg1_tiny_d1_b8_eval", dict(model_type="deit_tiny_patch16_224", depth=1, B=8, mode="eval")),
        ("tiny_d12_b4_eval", dict(model_type="deit_tiny_patch16_224", depth=12, B=4, mode="eval")),
        ("small_d12_b2_eval", dict(model_type="deit_small_patch16_224", depth=12, B=2, mode="eval")),
        ("tiny_d3_b4_skip", dict(model_type="deit_tiny_patch16_224", depth=3, B=4, mode="skip")),
        ("tiny_d3_b4_gumbel", dict(model_type="deit_tiny_patch16_224", depth=3, B=4, mode="gumbel")),
        ("tiny_d2_b4_warmup_jump", dict(model_type="deit_tiny_patch16_224", depth=2, B=4, mode="warmup_jump")),
    ]
    for name, sp in specs:
        sd, dims = fx.make_state_dict(sp["model_type"], sp["depth"], seed=11)
        x, _ = fx.make_batch(sp["B"], seed=730)
        kw, blend, skip, jump = {}, None, None, False
        if sp["mode"] == "skip":
            sd["block_skip_gating"][1] = torch.tensor([1.0, -1.0])       # block 1 hard-skipped (:496-500)
            skip = [False, True, False]
        m = ref_model(ns, sp["model_type"], sp["depth"], sd)
        if sp["mode"] in ("eval", "skip"):
            m.eval()
        elif sp["mode"] == "gumbel":
            m.train(); m.enable_block_gating = 1; m.use_gumbel = 1; m.gumbel_hard = False
            torch.manual_seed(5)
            blend = torch.stack([torch.nn.functional.gumbel_softmax(sd["block_skip_gating"][i], tau=0.5, hard=False, eps=1e-10, dim=-1)
                                 for i in range(sp["depth"])])
            torch.manual_seed(5)       # the reference now draws the same noise inside forward
        elif sp["mode"] == "warmup_jump":
            m.train(); m.enable_block_gating = 1; m.enable_warmup = 1; m.enable_jumping = 1
            blend = torch.full((sp["depth"], 2), 0.5); jump = True
        with torch.no_grad():
            out, (macs_embed, macs_list) = m(x)
        logits = out[0] if isinstance(out, tuple) else out
        # the restatement must agree with the reference bit for bit before its output is trusted anywhere
        with torch.no_grad():
            o2 = vo.forward(sd, x, sp["depth"], dims["num_heads"], blend=blend, skip=skip, enable_jumping=jump)
        assert torch.equal(o2, logits), (name, (o2 - logits).abs().max())
        cases[name] = dict(spec=sp, logits=logits.clone(), blend=blend, skip=skip, jump=jump, macs_embed=int(macs_embed),
                           macs_list=[[int(v) for v in row] for row in macs_list], x_sum=fx.checksum(x),
                           w_sum=fx.checksum(sd["blocks.0.mlp.fc1.weight"]))
        print(f"  {name}: logits {tuple(logits.shape)} max|.|={logits.abs().max():.4f}")
    torch.save(cases, os.path.join(OUT, "forward_cases.pt"))


def gen_train_step(ns):
    """One Stage-1 style loss + backward through the reference (student in train mode with soft Gumbel gates,
    dense teacher in eval, DistillationLoss 'soft' alpha .1 T 1 over timm-style soft targets): loss values and
    per-parameter gradient checksums + a few full gradients."""
    mt, depth, B = "deit_tiny_patch16_224", 2, 4
    sd, dims = fx.make_state_dict(mt, depth, seed=21)
    sd_t, _ = fx.make_state_dict(mt, depth, seed=22)
    x, _ = fx.make_batch(B, seed=731)
    tgt = fx.soft_targets(B, seed=731)
    student = ref_model(ns, mt, depth, sd, gumbel_hard=False)
    teacher = ref_model(ns, mt, depth, sd_t); teacher.eval()
    student.train(); student.enable_block_gating = 1; student.use_gumbel = 1
    torch.manual_seed(9)
    blend = torch.stack([torch.nn.functional.gumbel_softmax(sd["block_skip_gating"][i], tau=0.5, hard=False, eps=1e-10, dim=-1)
                         for i in range(depth)])
    torch.manual_seed(9)

    class SoftCE(torch.nn.Module):      # timm.loss.SoftTargetCrossEntropy (timm is not installed; public formula)
        def forward(self, a, t):
            return vo.soft_target_cross_entropy(a, t)

    crit = ns.losses.DistillationLoss(SoftCE(), teacher, "soft", 0.1, 1.0)
    outputs, _ = student(x)
    loss = crit(x, outputs, tgt)
    loss.backward()
    grads = {k: p.grad.clone() for k, p in student.named_parameters() if p.grad is not None}
    with torch.no_grad():
        t_logits, _ = teacher(x)
    # restatement check (forward values bit-exact; gradients through autograd of the restatement)
    sd_req = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lg = vo.forward(sd_req, x, depth, dims["num_heads"], blend=blend)
    l2, base, kd = vo.distillation_loss(lg, t_logits, tgt, 0.1, 1.0)
    assert torch.equal(lg.detach(), outputs[0].detach()) and abs(float(l2.detach()) - float(loss.detach())) < 1e-6
    out = dict(spec=dict(model_type=mt, depth=depth, B=B, seed=21, teacher_seed=22, batch_seed=731, alpha=0.1, T=1.0),
               blend=blend, loss=float(loss), base=float(base), kd=float(kd), logits=outputs[0].detach().clone(), teacher_logits=t_logits.clone(),
               grad_sums={k: fx.checksum(g) for k, g in grads.items()},
               grads_full={k: grads[k] for k in ["blocks.1.attn.proj.bias", "blocks.0.norm1.weight", "head.bias", "block_skip_gating",
                                                 "blocks.0.attn.qkv.bias", "patch_embed.proj.bias", "cls_token"]},
               grad_norm=float(torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values()))))
    torch.save(out, os.path.join(OUT, "train_step.pt"))
    print(f"  train_step: loss={float(loss):.6f} base={float(base):.6f} kd={float(kd):.6f} |g|={out['grad_norm']:.6f}")


def main():
    assert ref_shim.available(), "reference checkout not found"
    os.makedirs(OUT, exist_ok=True)
    ns = ref_shim.load()
    torch.set_num_threads(8)
    gen_forward_cases(ns)
    gen_train_step(ns)
    if True:
        try:
            from oracle import gen_golden_admm
            gen_golden_admm.main(ns)
        except ImportError:
            print("  (ADMM goldens: generator not present yet)")


if __name__ == "__main__":
    main()
