"""ORACLE (test infrastructure only) — writes tests/golden/t2t_cases.pt FROM THE UNMODIFIED REFERENCE T2T-ViT.

Run here (the only place /root/reference exists):   python -m oracle.gen_golden_t2t

SURVEY.md §8 row a-T: the reference's `T2T_ViT` (T2TViT/models/t2t_vit.py:107-208) imports and runs on CPU for the
plain path (eval, hard block skipping, jumping connections).  Its block-gating branch (:181-189) cannot execute at
HEAD because `F` is never imported in that file; the generator supplies the missing name
(`module.F = torch.nn.functional`) — the source is not edited — so the softmax `distrib` blend of :186-188 is pinned too.
Two more load-time accommodations, neither touching the arithmetic: `block_skip_gating` is an expanded (overlapping)
tensor in the reference (:157) and cannot be `load_state_dict`-ed into, so it is assigned; dropouts are inactive (eval).

Inputs / weights come from seeds (oracle/fixtures.py); only reference OUTPUTS (and sub-sampled gradients) are stored.
"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fixtures as fx, ref_shim, vit_oracle as vo  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

GRAD_KEYS = ["cls_token", "head.bias", "norm.weight", "blocks.0.attn.qkv.weight", "blocks.0.mlp.fc1.bias", "blocks.1.mlp.fc2.weight",
             "blocks.1.attn.proj.bias", "blocks.0.norm1.weight", "tokens_to_token.project.weight", "tokens_to_token.project.bias",
             "tokens_to_token.attention2.kqv.weight", "tokens_to_token.attention1.kqv.weight", "tokens_to_token.attention1.mlp.0.weight"]


def sample(t, n=257):
    """A fixed sub-sample (every k-th element) — enough to pin a gradient without storing megabytes."""
    f = t.detach().flatten()
    return f[:: max(1, f.numel() // n)].clone()


def ref_t2t(mod, depth, sd, **kw):
    m = mod.T2T_ViT(tokens_type='performer', embed_dim=384, depth=depth, num_heads=6, mlp_ratio=3., **kw)   # t2t_vit_14, :244-250
    m.block_skip_gating.data = sd["block_skip_gating"].clone()
    missing, unexpected = m.load_state_dict({k: v for k, v in sd.items() if k != "block_skip_gating"}, strict=False)
    assert not unexpected and missing == ["block_skip_gating"], (missing, unexpected)
    return m.eval()


def main():
    if not ref_shim.available():
        raise SystemExit("reference checkout not present; golden vectors can only be generated where /root/reference exists")
    ref_shim.load()
    mod = importlib.import_module("T2TViT.models.t2t_vit")
    mod.F = torch.nn.functional      # the name t2t_vit.py:184,186 uses without importing
    os.makedirs(OUT, exist_ok=True)
    cases = {}

    specs = [("t2t14_d14_b2_eval", dict(depth=14, B=2, mode="eval")),
             ("t2t14_d3_b4_skip_jump", dict(depth=3, B=4, mode="skip_jump")),
             ("t2t14_d3_b4_softgate", dict(depth=3, B=4, mode="softgate")),
             ("t2t14_d2_b4_grads", dict(depth=2, B=4, mode="grads"))]
    for name, sp in specs:
        sd, dims = fx.make_state_dict("t2t_vit_14", sp["depth"], seed=23)
        x, _ = fx.make_batch(sp["B"], seed=730)
        blend, skip, jump, kw = None, None, False, {}
        if sp["mode"] == "skip_jump":
            sd["block_skip_gating"][1] = torch.tensor([1.0, -1.0])
            skip, jump, kw = [False, True, False], True, dict(enable_jumping=True)
        if sp["mode"] == "softgate":
            sd["block_skip_gating"] = torch.tensor([[-1.0, 1.0], [0.3, -0.2], [0.0, 2.0]])
            blend = torch.stack([torch.softmax(sd["block_skip_gating"][i], dim=0) for i in range(3)])
            kw = dict(enable_block_gating=True)
        m = ref_t2t(mod, sp["depth"], sd, **kw)
        entry = dict(spec=sp, blend=blend, skip=skip, jump=jump, x_sum=fx.checksum(x), w_sum=fx.checksum(sd["blocks.0.mlp.fc1.weight"]))
        if sp["mode"] == "grads":
            r = torch.randn(sp["B"], 1000, generator=fx._gen(730, "dlogits")) * 0.1
            for p in m.parameters():
                p.requires_grad_(True)
            logits, _ = m(x)
            (logits * r).sum().backward()
            named = dict(m.named_parameters())
            entry["dlogits_sum"] = fx.checksum(r)
            entry["grads"] = {k: dict(sample=sample(named[k].grad), sums=fx.checksum(named[k].grad)) for k in GRAD_KEYS}
            logits = logits.detach()
            macs_embed, macs_list = None, None
        else:
            with torch.no_grad():
                logits, (macs_embed, macs_list) = m(x)
        with torch.no_grad():
            tok_ref, _ = m.tokens_to_token(x)
            tok, macs_o = vo.t2t_tokens(sd, x)
            o2 = vo.forward(sd, x, sp["depth"], 6, eps=1e-5, blend=blend, skip=skip, enable_jumping=jump, tokens=tok)
        # the restatement must agree with the reference bit for bit before its output is trusted anywhere
        assert torch.equal(tok, tok_ref), (name, (tok - tok_ref).abs().max())
        assert torch.equal(o2, logits), (name, (o2 - logits).abs().max())
        entry.update(logits=logits.clone(), tokens_sample=sample(tok_ref, 4001), tokens_sum=fx.checksum(tok_ref))
        if macs_embed is not None:
            assert int(macs_embed) == int(macs_o)
            entry.update(macs_embed=int(macs_embed), macs_list=[[int(v) for v in row] for row in macs_list])
        cases[name] = entry
        print(f"  {name}: logits {tuple(logits.shape)} max|.|={logits.abs().max():.4f}")
    torch.save(cases, os.path.join(OUT, "t2t_cases.pt"))
    print("wrote", os.path.join(OUT, "t2t_cases.pt"))


if __name__ == "__main__":
    main()
