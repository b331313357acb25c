"""CPU: the oracle's T2T-ViT restatement (front end + backbone) against the golden vectors the UNMODIFIED reference wrote
(tests/golden/t2t_cases.pt, generator: oracle/gen_golden_t2t.py), and the host mirror's state-dict / front-end contract.

SURVEY.md §8 row a-T.  The reference cannot run T2T with block gating at HEAD (`F` is not imported in t2t_vit.py); the generator
supplied that name, so the soft-gate case pins the blend of t2t_vit.py:181-189 as written."""
import os

import pytest
import torch

from oracle import fixtures as fx, vit_oracle as vo

GRAD_KEYS_FRONT = ("tokens_to_token.project.weight", "tokens_to_token.attention2.kqv.weight", "tokens_to_token.attention1.kqv.weight")


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
    assert fx.checksum(x) == c["x_sum"] and fx.checksum(sd["blocks.0.mlp.fc1.weight"]) == c["w_sum"]
    return sd, x


def sample(t, n=257):
    f = t.detach().flatten()
    return f[:: max(1, f.numel() // n)]


@pytest.mark.parametrize("name", ["t2t14_d14_b2_eval", "t2t14_d3_b4_skip_jump", "t2t14_d3_b4_softgate"])
def test_oracle_reproduces_reference_t2t_logits(cases, name):
    c = cases[name]
    sd, x = _inputs(c)
    with torch.no_grad():
        tok, macs = vo.t2t_tokens(sd, x)
        out = vo.forward(sd, x, c["spec"]["depth"], 6, eps=1e-5, blend=c["blend"], skip=c["skip"], enable_jumping=c["jump"], tokens=tok)
    assert torch.allclose(sample(tok, 4001), c["tokens_sample"], rtol=0, atol=2e-5)      # same ops, same machine class: ~ulp
    assert (out - c["logits"]).abs().max() <= 2e-5 * c["logits"].abs().max()
    assert int(macs) == c["macs_embed"]
    N, C, H, Fh, B = 197, 384, 6, 1152, c["spec"]["B"]
    want = [vo.block_macs(B, N, C, H, Fh) if not (c["skip"] and c["skip"][i]) else [] for i in range(c["spec"]["depth"])]
    assert want == c["macs_list"]


def test_oracle_reproduces_reference_t2t_gradients(cases):
    c = cases["t2t14_d2_b4_grads"]
    sd, x = _inputs(c)
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    r = torch.randn(4, 1000, generator=fx._gen(730, "dlogits")) * 0.1
    assert fx.checksum(r) == c["dlogits_sum"]
    tok, _ = vo.t2t_tokens(sd, x)
    out = vo.forward(sd, x, 2, 6, eps=1e-5, tokens=tok)
    (out * r).sum().backward()
    for k, g in c["grads"].items():
        got = sample(sd[k].grad)
        assert (got - g["sample"]).abs().max() <= 1e-4 * g["sample"].abs().max() + 1e-7, k


def test_host_mirror_state_dict_matches_reference_keys():
    """Same keys and shapes as the reference T2T_ViT's state dict (listed from the reference by fixtures.param_shapes)."""
    from uvc_b200.T2TViT.models import T2T_ViT
    m = T2T_ViT(tokens_type='performer', embed_dim=384, depth=2, num_heads=6, mlp_ratio=3.)
    want = fx.param_shapes(embed_dim=384, depth=2, num_heads=6, mlp_ratio=3, t2t=True)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got.pop("gumbel.weight") == (1, 384) and got.pop("gumbel.bias") == (1,)     # the token scorer UVC adds (absent from the reference T2T)
    assert got == {k: tuple(v) for k, v in want.items()}
    assert not m.pos_embed.requires_grad and torch.equal(m.pos_embed, fx.sinusoid_table(197, 384))
    assert all(b.attn.qkv.bias is None for b in m.blocks) and m.norm.eps == 1e-5


def test_front_end_has_no_host_path_and_counts_macs_like_the_oracle(cases):
    """tokens_to_token runs behind uvc_t2t_forward (csrc/t2t_frontend.cu): on a CPU tensor it must fail loudly, never compute.  Its MAC
    bookkeeping is host arithmetic and must equal the reference's (through the oracle)."""
    from uvc_b200.T2TViT.models import T2T_ViT
    c = cases["t2t14_d2_b4_grads"]
    sd, x = _inputs(c)
    m = T2T_ViT(tokens_type='performer', embed_dim=384, depth=2, num_heads=6, mlp_ratio=3.).eval()
    m.load_state_dict(sd, strict=False)
    with torch.no_grad():
        _, macs_o = vo.t2t_tokens(sd, x)
    t2t = m.tokens_to_token
    assert t2t.attention1.macs(x.shape[0], 56 * 56) + t2t.attention2.macs(x.shape[0], 28 * 28) == int(macs_o)
    with pytest.raises(Exception, match="CUDA"):
        t2t(x)
    with pytest.raises(Exception, match="CUDA"):
        m(x)                                       # no CPU path for the backbone either
    with pytest.raises(RuntimeError, match="parameter container"):
        t2t.attention1(x)


@pytest.mark.skipif(not __import__("oracle.ref_shim", fromlist=["x"]).available(), reason="/root/reference not present on this machine")
def test_oracle_equals_reference_t2t_modules_live():
    """Where the reference checkout exists (the build container): the restatement against the UNMODIFIED T2T modules imported in place, on fresh
    seeds (not only the committed fixture) — tokens_to_token output, logits with hard skipping, and the MAC bookkeeping, bit for bit."""
    import importlib
    from oracle import ref_shim
    ref_shim.load()
    mod = importlib.import_module("T2TViT.models.t2t_vit")
    for seed, depth, B in ((101, 2, 2), (202, 3, 1)):
        sd, _ = fx.make_state_dict("t2t_vit_14", depth, seed=seed)
        if depth == 3:
            sd["block_skip_gating"][0] = torch.tensor([0.5, 0.5])            # not strictly greater: skipped (t2t_vit.py:192)
        m = mod.T2T_ViT(tokens_type='performer', embed_dim=384, depth=depth, num_heads=6, mlp_ratio=3.)
        m.block_skip_gating.data = sd["block_skip_gating"].clone()
        m.load_state_dict({k: v for k, v in sd.items() if k != "block_skip_gating"}, strict=False)
        m.eval()
        x, _ = fx.make_batch(B, seed=seed)
        with torch.no_grad():
            ref_logits, (ref_macs_embed, ref_macs_list) = m(x)
            ref_tok, _ = m.tokens_to_token(x)
            tok, macs = vo.t2t_tokens(sd, x)
            skip = [not bool(sd["block_skip_gating"][i, 1] > sd["block_skip_gating"][i, 0]) for i in range(depth)]
            out = vo.forward(sd, x, depth, 6, eps=1e-5, skip=skip, tokens=tok)
        assert torch.equal(tok, ref_tok) and torch.equal(out, ref_logits)
        assert int(macs) == int(ref_macs_embed)
        assert [[int(v) for v in r] for r in ref_macs_list] == [vo.block_macs(B, 197, 384, 6, 1152) if not s else [] for s in skip]
