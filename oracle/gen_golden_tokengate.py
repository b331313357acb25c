"""ORACLE (test infrastructure only) — token-gate golden FROM THE UNMODIFIED REFERENCE.

Drives the reference's own `DistilledVisionTransformer.forward(x, tau, ratio)` (models/model_distilled.py:446-456, gumbel_softmax :36-63,
scatter :21-33; imported in place by oracle/ref_shim.py) on CPU with `tau > 0`, hard block skipping (no other random draw in the forward) and a
rewound generator, so the Gumbel noise the reference consumed is known and stored.  Records the noise, the kept-token indices, the mask values
and the logits, and asserts that `oracle/vit_oracle.token_gate` + `forward(token_mask=...)` reproduce them bit for bit.
Run here (the only place /root/reference exists):   python -m oracle.gen_golden_tokengate
"""
import os

import torch

from oracle import fixtures as fx, ref_shim, vit_oracle as vo

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = [dict(name="tiny_d2_b6_tau0.7", model_type="deit_tiny_patch16_224", depth=2, B=6, seed=9, batch_seed=4, tau=0.7, ratio=0.9, noise_seed=123),
         dict(name="tiny_d1_b16_tau2.5", model_type="deit_tiny_patch16_224", depth=1, B=16, seed=10, batch_seed=5, tau=2.5, ratio=0.9, noise_seed=7),
         dict(name="small_d1_b4_tau1_r0.5", model_type="deit_small_patch16_224", depth=1, B=4, seed=12, batch_seed=6, tau=1.0, ratio=0.5, noise_seed=31)]


def main(ns=None):
    ns = ns or ref_shim.load()
    out = {}
    for sp in CASES:
        sd, dims = fx.make_state_dict(sp["model_type"], sp["depth"], seed=sp["seed"])
        g = torch.Generator().manual_seed(sp["seed"] + 1000)
        sd["gumbel.weight"] = torch.randn(1, dims["embed_dim"], generator=g) * 0.5        # a trained-looking scorer (timm init would give near-constant scores)
        sd["gumbel.bias"] = torch.randn(1, generator=g) * 0.1
        x, _ = fx.make_batch(sp["B"], seed=sp["batch_seed"])
        m = ref_shim.make_ref_model(ns, sp["model_type"], depth=sp["depth"], gumbel_hard=True, enable_patch_gating=2)
        missing, unexpected = m.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected
        m.train()                                        # enable_block_gating = 0: blocks run by gate sign, no Gumbel draw besides the token gate's
        torch.manual_seed(sp["noise_seed"])
        noise = -torch.empty(sp["B"], 196).exponential_().log()
        # hook the multiply at :456 through the scorer's input/output: capture the mask the reference applied
        cap = {}
        orig = ns.model_distilled.gumbel_softmax

        def spy(logits, k=0.9, tau=1, hard=False, eps=1e-10, dim=-1):
            r = orig(logits, k=k, tau=tau, hard=hard, eps=eps, dim=dim)
            cap["ret"], cap["k"] = r.detach().clone(), k
            return r
        ns.model_distilled.gumbel_softmax = spy
        try:
            torch.manual_seed(sp["noise_seed"])
            with torch.no_grad():
                (logits, _), (macs_embed, macs_list) = m(x, sp["tau"], sp["ratio"])
        finally:
            ns.model_distilled.gumbel_softmax = orig
        k = int(sp["ratio"] * 196)
        assert cap["k"] == k
        mask_ref = cap["ret"].clone(); mask_ref[:, 0] = 1.0
        kept = mask_ref > 0.5
        # the restatement, fed the recorded noise, must agree bit for bit
        with torch.no_grad():
            pe = vo.patch_embed(sd, x)
            mask_o, _ = vo.token_gate(sd, pe, noise, sp["tau"], k)
            lo = vo.forward(sd, x, sp["depth"], dims["num_heads"], token_mask=mask_o)
        assert torch.equal(mask_o, mask_ref), sp["name"]
        assert torch.equal(lo, logits), float((lo - logits).abs().max())
        # how close the selection boundary is: gap between the k-th and (k+1)-th largest y, relative (a TF32-level score error of 1e-3 flips rows below it)
        out[sp["name"]] = dict(spec=sp, noise=noise, kept=kept, mask=mask_ref, logits=logits, x_sum=fx.checksum(x), gumbel_w=sd["gumbel.weight"], gumbel_b=sd["gumbel.bias"])
        print(f"  tokengate {sp['name']}: kept {int(kept.sum())} of {kept.numel()} tokens, logits checksum {fx.checksum(logits)}")
    torch.save(out, os.path.join(OUT, "tokengate.pt"))


if __name__ == "__main__":
    main()
