"""CPU tests of the oracle: (1) it reproduces the committed golden vectors, which were written by the
UNMODIFIED reference (oracle/gen_golden.py); (2) where /root/reference is present it is compared with the
reference modules directly, bit for bit; (3) the reference's own known answers (log/*.log) are reproduced."""
import os

import pytest
import torch

from oracle import fixtures as fx, ref_shim, vit_oracle as vo


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def test_oracle_reproduces_reference_golden_logits(golden_dir):
    cases = _load(golden_dir, "forward_cases.pt")
    for name, c in cases.items():
        if name == "small_d12_b2_eval":
            continue   # covered on the GPU box; keeps the CPU suite short
        sp = c["spec"]
        sd, dims = fx.make_state_dict(sp["model_type"], sp["depth"], seed=11)
        if sp["mode"] == "skip":
            sd["block_skip_gating"][1] = torch.tensor([1.0, -1.0])
        x, _ = fx.make_batch(sp["B"], seed=730)
        assert fx.checksum(x) == c["x_sum"], "synthetic batch drifted from the one the reference saw"
        assert fx.checksum(sd["blocks.0.mlp.fc1.weight"]) == c["w_sum"]
        with torch.no_grad():
            out = vo.forward(sd, x, sp["depth"], dims["num_heads"], blend=c["blend"], skip=c["skip"], enable_jumping=c["jump"])
        assert torch.equal(out, c["logits"]), name
        C, H = dims["embed_dim"], dims["num_heads"]
        want = vo.block_macs(sp["B"], 197, C, H, 4 * C)
        for row, skipped in zip(c["macs_list"], c["skip"] or [False] * sp["depth"]):
            assert row == ([] if skipped else want)
        assert c["macs_embed"] == sp["B"] * 196 * C * 256 * 3


def test_oracle_reproduces_reference_golden_train_step(golden_dir):
    g = _load(golden_dir, "train_step.pt")
    sp = g["spec"]
    sd, dims = fx.make_state_dict(sp["model_type"], sp["depth"], seed=sp["seed"])
    x, _ = fx.make_batch(sp["B"], seed=sp["batch_seed"])
    tgt = fx.soft_targets(sp["B"], seed=sp["batch_seed"])
    sd = {k: v.requires_grad_(True) for k, v in sd.items()}
    logits = vo.forward(sd, x, sp["depth"], dims["num_heads"], blend=g["blend"])
    loss, base, kd = vo.distillation_loss(logits, g["teacher_logits"], tgt, sp["alpha"], sp["T"])
    assert torch.equal(logits.detach(), g["logits"])
    assert abs(float(loss) - g["loss"]) < 1e-6 and abs(float(kd) - g["kd"]) < 1e-8
    loss.backward()
    for k, want in g["grads_full"].items():
        if k == "block_skip_gating":
            continue    # its gradient flows through the Gumbel sample inside the reference's forward
        torch.testing.assert_close(sd[k].grad, want, rtol=1e-5, atol=1e-7)


def test_known_answer_initial_flops():
    """`** Initial FLOP size: 2506.98M` for DeiT-Tiny (log/deit-tiny-log.log:7) from the MAC formulas."""
    C, H, L = 192, 3, 12
    embed = 196 * C * 256 * 3
    blocks = L * sum(vo.block_macs(1, 197, C, H, 4 * C))
    assert f"{2 * (embed + blocks) / 1e6:.2f}" == "2506.98"
    share = 2 * sum(vo.block_macs(1, 197, C, H, 4 * C)) / (2 * (embed + blocks))
    assert f"{100 * (1 - 2 * share):.2f}" == "83.72"        # two skipped blocks (log/deit-tiny-log.log warm-up epochs)


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present on this machine")
def test_oracle_equals_reference_modules():
    ns = ref_shim.load()
    sd, dims = fx.make_state_dict("deit_tiny_patch16_224", 2, seed=3)
    m = ref_shim.make_ref_model(ns, "deit_tiny_patch16_224", depth=2)
    m.load_state_dict(sd, strict=False)
    m.eval()
    x, _ = fx.make_batch(2, seed=1)
    with torch.no_grad():
        ref, _ = m(x)
        got = vo.forward(sd, x, 2, dims["num_heads"], skip=[False, False])
    assert torch.equal(ref, got)
    # loss restatement vs utils/losses.py
    t = torch.randn(2, 1000)
    tgt = fx.soft_targets(2, seed=5)

    class SoftCE(torch.nn.Module):
        def forward(self, a, b):
            return vo.soft_target_cross_entropy(a, b)

    class T(torch.nn.Module):
        def forward(self, inp):
            return t, None

    crit = ns.losses.DistillationLoss(SoftCE(), T(), "soft", 0.1, 3.0)
    want = crit(x, (ref, ref), tgt)
    got, _, _ = vo.distillation_loss(ref, t, tgt, 0.1, 3.0)
    assert torch.equal(want, got)
