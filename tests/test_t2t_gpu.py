"""GPU parity of the T2T-ViT-14 backbone through uvc_vit_forward / uvc_vit_backward (`pe_in` / `d_pe`), SURVEY.md §8 row a-T:
14 Blocks of C=384, H=6, Fh=1152 with LayerNorm eps 1e-5 and no qkv bias, fed by the tokens_to_token front end.

Checked against the golden vectors the unmodified reference wrote (tests/golden/t2t_cases.pt) and the oracle on this box's CPU.
Tolerance: logits within 1e-3 relative of the fp32 reference (north_star); gradients within 5e-3 of the gradient's max."""
import os

import pytest
import torch

from oracle import fixtures as fx, vit_oracle as vo

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("precision")]

LOGIT_TOL = 1e-3
GRAD_TOL = 5e-3


def rel(a, b):
    return ((a.cpu() - b.cpu()).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def sample(t, n=257):
    f = t.detach().flatten()
    return f[:: max(1, f.numel() // n)]


@pytest.fixture(scope="module")
def cases(golden_dir):
    return torch.load(os.path.join(golden_dir, "t2t_cases.pt"), weights_only=False)


def _inputs(c):
    sp = c["spec"]
    sd, dims = fx.make_state_dict("t2t_vit_14", sp["depth"], seed=23)
    if sp["mode"] == "skip_jump":
        sd["block_skip_gating"][1] = torch.tensor([1.0, -1.0])
    if sp["mode"] == "softgate":
        sd["block_skip_gating"] = torch.tensor([[-1.0, 1.0], [0.3, -0.2], [0.0, 2.0]])
    x, _ = fx.make_batch(sp["B"], seed=730)
    assert fx.checksum(x) == c["x_sum"]
    return sd, x


def build(depth, sd, **kw):
    from uvc_b200.T2TViT.models import T2T_ViT
    m = T2T_ViT(tokens_type='performer', embed_dim=384, depth=depth, num_heads=6, mlp_ratio=3., **kw)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and set(missing) == {"gumbel.weight", "gumbel.bias"}
    return m.cuda()


@pytest.mark.parametrize("name", ["t2t14_d14_b2_eval", "t2t14_d3_b4_skip_jump", "t2t14_d3_b4_softgate"])
def test_t2t_logits_match_reference_golden(cases, name):
    c = cases[name]
    sd, x = _inputs(c)
    kw = {}
    if c["jump"]:
        kw["enable_jumping"] = True
    if c["spec"]["mode"] == "softgate":
        kw["enable_block_gating"] = True
    m = build(c["spec"]["depth"], sd, **kw).eval()
    launches0 = _launches()
    with torch.no_grad():
        out, (macs_embed, macs_list) = m(x.cuda())
    assert _launches() > launches0                       # the backbone ran in libuvc_sm100.so
    e = rel(out, c["logits"])
    print(f"{name}: logits rel err {e:.3e}")
    assert e < LOGIT_TOL
    assert int(macs_embed) == c["macs_embed"] and [[int(v) for v in r] for r in macs_list] == c["macs_list"]   # skip decisions exact


def _launches():
    from uvc_b200 import _lib
    return _lib.load().uvc_launch_count()


def test_t2t_gradients_match_reference_golden(cases):
    """Backward through the engine (d_pe handed to torch autograd for the front end) against the reference's autograd."""
    c = cases["t2t14_d2_b4_grads"]
    sd, x = _inputs(c)
    m = build(2, sd).eval()                              # eval: the reference's performer dropouts are off in the golden run
    r = (torch.randn(4, 1000, generator=fx._gen(730, "dlogits")) * 0.1).cuda()
    out, _ = m(x.cuda())
    assert rel(out, c["logits"]) < LOGIT_TOL
    (out * r).sum().backward()
    named = dict(m.named_parameters())
    worst = 0.0
    for k, g in c["grads"].items():
        assert named[k].grad is not None, k
        e = rel(sample(named[k].grad), g["sample"])
        worst = max(worst, e)
        assert e < GRAD_TOL, (k, e)
    assert m.pos_embed.grad is None                      # fixed sinusoid table (t2t_vit.py:119)
    print(f"t2t grads: worst rel err {worst:.3e}")


def test_t2t_gumbel_gate_training_step_matches_oracle():
    """Training-mode forward + backward with a sampled Gumbel blend and token gate inputs, against the oracle driven with the SAME
    blend (drawn once here) — the T2T counterpart of the DeiT gated test; front end in eval to keep its dropouts out."""
    from uvc_b200.models.model_distilled import _VitFunction, _engine_param_list
    sd, _ = fx.make_state_dict("t2t_vit_14", 3, seed=29)
    x, _ = fx.make_batch(3, seed=731)
    m = build(3, sd, enable_block_gating=True, use_gumbel=True, gumbel_hard=False).train()
    torch.manual_seed(3)
    blend = torch.softmax(torch.randn(3, 2), dim=-1)
    tmask = (torch.rand(3, 196) > 0.1).float(); tmask[:, 0] = 1
    with torch.no_grad():
        tok, _ = vo.t2t_tokens(sd, x)
    tok_o = tok.clone().requires_grad_(True)
    sdo = {k: v.clone().requires_grad_(v.is_floating_point() and k != "pos_embed") for k, v in sd.items()}
    blend_o = blend.clone().requires_grad_(True)
    out_o = vo.forward(sdo, x, 3, 6, eps=1e-5, blend=blend_o, token_mask=tmask, tokens=tok_o)
    r = torch.randn(3, 1000) * 0.1
    (out_o * r).sum().backward()

    tok_g = tok.cuda().requires_grad_(True)
    blend_g = blend.cuda().requires_grad_(True)
    params = [p for _, p in _engine_param_list(m)]
    out = _VitFunction.apply(m, tok_g, blend_g, None, tmask.cuda(), None, *params)
    assert rel(out, out_o) < LOGIT_TOL
    (out * r.cuda()).sum().backward()
    assert rel(tok_g.grad, tok_o.grad) < GRAD_TOL
    assert rel(blend_g.grad, blend_o.grad) < GRAD_TOL
    for k in ("blocks.0.attn.qkv.weight", "blocks.2.mlp.fc2.weight", "blocks.1.norm2.bias", "cls_token", "head.weight"):
        assert rel(dict(m.named_parameters())[k].grad, sdo[k].grad) < GRAD_TOL, k


@pytest.mark.parametrize("B", [1, 3])
def test_front_end_tokens_and_gradients_match_oracle(B):
    """tokens_to_token alone (uvc_t2t_forward / uvc_t2t_backward, csrc/t2t_frontend.cu) against the oracle's restatement of
    t2t_vit.py:46-105 + token_performer.py:31-69 -- the restatement that is pinned bit for bit to the unmodified reference in test_t2t_oracle.py.
    Tokens within 1e-3 of the maximum, every parameter gradient within 5e-3 of its maximum."""
    sd, _ = fx.make_state_dict("t2t_vit_14", 1, seed=41)
    x, _ = fx.make_batch(B, seed=733)
    m = build(1, sd).eval()
    sdo = {k: v.clone().requires_grad_(v.is_floating_point() and k.startswith("tokens_to_token.") and not k.endswith(".w")) for k, v in sd.items()}
    tok_o, macs_o = vo.t2t_tokens(sdo, x)
    r = torch.randn(B, 196, 384, generator=fx._gen(733, "dtok")) * 0.05
    (tok_o * r).sum().backward()
    n0 = _launches()
    tok, macs = m.tokens_to_token(x.cuda())
    assert _launches() > n0 and int(macs) == int(macs_o)
    assert rel(tok, tok_o) < LOGIT_TOL
    (tok * r.cuda()).sum().backward()
    named = dict(m.named_parameters())
    worst = 0.0
    for k, v in sdo.items():
        if not v.requires_grad:
            continue
        assert named[k].grad is not None, k
        e = rel(named[k].grad, v.grad)
        worst = max(worst, e)
        assert e < GRAD_TOL, (k, e)
    assert named["tokens_to_token.attention1.w"].grad is None
    print(f"t2t front end B={B}: tokens rel err {rel(tok, tok_o):.2e}, worst gradient rel err {worst:.2e}")


def test_front_end_dropout_is_train_only_unbiased_and_replayed_in_backward():
    """nn.Dropout(0.1) of the two Token_performers (token_performer.py:20,28,51,66): off in eval; in train mode the tokens change, their mean over
    many draws approaches the eval tokens (keep / (1 - p) is unbiased to first order), and the backward replays the forward's mask
    (finite-difference check of one bias gradient along the same seed)."""
    sd, _ = fx.make_state_dict("t2t_vit_14", 1, seed=43)
    x, _ = fx.make_batch(2, seed=734)
    m = build(1, sd)
    t2t = m.tokens_to_token
    xe = x.cuda()
    with torch.no_grad():
        t_eval = t2t.eval()(xe)[0]
        t2t.train()
        torch.manual_seed(5)
        t_a = t2t(xe)[0]
        torch.manual_seed(5)
        t_b = t2t(xe)[0]
        t_c = t2t(xe)[0]
    # the seed comes from torch's (CPU) generator: same seed -> same mask (the per-image kptv sums are fp32 atomics, so equal up to summation order)
    assert rel(t_a, t_b) < 1e-3 and rel(t_a, t_c) > 1e-2
    assert rel(t_a, t_eval) > 1e-2                                      # dropout is really on
    # backward consistency: d/d(project.bias) of sum(tokens * r) is colsum(r) whatever the mask; d/d(attention2.mlp.2.bias) depends on the mask
    r = torch.randn(2, 196, 384, generator=fx._gen(734, "r")).cuda() * 0.05
    torch.manual_seed(9)
    tok = t2t(xe)[0]
    (tok * r).sum().backward()
    g = t2t.attention2.mlp[2].bias.grad.clone()
    base = float((tok.detach() * r).sum())
    j = int(g.abs().argmax())
    h = 1e-2
    with torch.no_grad():
        t2t.attention2.mlp[2].bias[j] += h
        torch.manual_seed(9)
        up = float((t2t(xe)[0] * r).sum())
    fd = (up - base) / h
    assert abs(fd - float(g[j])) < 0.05 * abs(float(g[j])) + 1e-3, (fd, float(g[j]))
