"""GPU parity of the token slimming gate (uvc_token_gate_* through uvc_b200.models.token_gate) against tests/golden/tokengate.pt, written by the
UNMODIFIED reference (oracle/gen_golden_tokengate.py: models/model_distilled.py:446-456 run with a rewound generator so the Gumbel draw is
known).  The kept-token INDICES must be equal, not merely close: north_star asks for bit-exact mask indices."""
import os

import pytest
import torch

from oracle import fixtures as fx, vit_oracle as vo

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("precision")]


def rel(a, b):
    return ((a.cpu() - b.cpu()).abs().max() / b.abs().max().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def golden(golden_dir):
    return torch.load(os.path.join(golden_dir, "tokengate.pt"), weights_only=False)


def build(sp, g):
    from test_model_gpu import build as b
    sd, dims = fx.make_state_dict(sp["model_type"], sp["depth"], seed=sp["seed"])
    sd["gumbel.weight"], sd["gumbel.bias"] = g["gumbel_w"], g["gumbel_b"]
    m = b(sp["model_type"], sp["depth"], sd, enable_patch_gating=2).train()
    return m, sd, dims


@pytest.mark.parametrize("name", ["tiny_d2_b6_tau0.7", "tiny_d1_b16_tau2.5", "small_d1_b4_tau1_r0.5"])
def test_kept_token_indices_equal_reference_golden(golden, name, monkeypatch):
    from uvc_b200.models import token_gate as tg
    g = golden[name]; sp = g["spec"]
    m, sd, dims = build(sp, g)
    x, _ = fx.make_batch(sp["B"], seed=sp["batch_seed"])
    assert all(abs(a - b) <= 1e-9 * max(1.0, abs(b)) for a, b in zip(fx.checksum(x), g["x_sum"]))     # fp64 sums: summation order differs between CPUs
    k = int(sp["ratio"] * 196)
    monkeypatch.setattr(tg, "gumbel_noise_like", lambda B, np_, device: g["noise"].to(device))
    pe, mask = tg.token_gate_mask(m, x.cuda(), None, sp["tau"], k)
    kept = mask.detach() > 0.5
    assert torch.equal(kept.cpu(), g["kept"]), f"{int((kept.cpu() != g['kept']).sum())} kept-token flags differ from the reference"
    assert (kept.sum(1) == g["kept"].sum(1).cuda()).all()
    assert rel(mask.detach(), g["mask"]) < 1e-6                       # the straight-through VALUE ((hard - y) + y), not just the indices
    (logits, _), _ = m(x.cuda(), sp["tau"], sp["ratio"])             # whole forward through the public API with the same draw
    assert rel(logits.detach(), g["logits"]) < 1e-3


def test_topk_ties_go_to_the_lower_index_like_torch_topk(golden):
    """identical scores and noise in groups -> exact ties in y; the kernel's rank rule must pick what torch.topk picks on this device"""
    from uvc_b200 import ops
    B, np_, C = 8, 196, 64
    gen = torch.Generator(device="cuda").manual_seed(3)
    feat = torch.randn(B, np_, C, device="cuda", generator=gen)
    feat[:, 50:120] = feat[:, 50:51]                                  # 70 identical rows per image
    feat[:, 150:] = 0.0                                               # and 46 zero rows
    noise = torch.zeros(B, np_, device="cuda")
    noise[:, ::2] = 0.25                                              # two noise levels -> tie groups of 35 / 23 members
    w = torch.randn(C, device="cuda", generator=gen)
    bias = torch.tensor([0.1], device="cuda")
    for k in (1, 30, 100, 176, 196):
        mask, y, ls, sc = ops.token_gate_fwd(feat.view(B * np_, C), w, None, bias, None, noise, 0.9, k, B, np_)
        idx = y.topk(k, dim=-1)[1]
        want = torch.zeros_like(y, dtype=torch.bool).scatter_(1, idx, True); want[:, 0] = True
        assert torch.equal(mask > 0.5, want), k
        ref_sc = feat @ w + bias
        assert rel(sc, ref_sc) < 1e-5


def test_gate_gradients_match_autograd_of_the_oracle_restatement(golden):
    """d(mask)/d(pe, gumbel.weight, gumbel.bias) through the straight-through estimator: uvc_token_gate_bwd + _apply vs autograd of vo.token_gate"""
    from uvc_b200.models.token_gate import _TokenGateFn
    g = golden["tiny_d1_b16_tau2.5"]; sp = g["spec"]
    B, C, k = sp["B"], 192, int(sp["ratio"] * 196)
    gen = torch.Generator().manual_seed(5)
    pe0 = torch.randn(B, 196, C, generator=gen)
    w0, b0 = torch.randn(1, C, generator=gen) * 0.3, torch.randn(1, generator=gen)
    up = torch.randn(B, 196, generator=gen)
    pe_r, w_r, b_r = (t.clone().requires_grad_(True) for t in (pe0, w0, b0))
    mask_r, _ = vo.token_gate({"gumbel.weight": w_r, "gumbel.bias": b_r}, pe_r, g["noise"], sp["tau"], k)
    (mask_r * up).sum().backward()
    pe, w, b = (t.cuda().requires_grad_(True) for t in (pe0, w0, b0))
    mask = _TokenGateFn.apply(pe.detach(), w.detach().reshape(-1).contiguous(), None, pe, None, w, b, g["noise"].cuda(), sp["tau"], k)
    assert torch.equal(mask.detach().cpu() > 0.5, mask_r.detach() > 0.5)
    (mask * up.cuda()).sum().backward()
    assert rel(pe.grad, pe_r.grad) < 1e-4 and rel(w.grad, w_r.grad) < 1e-4
    # log_softmax is shift invariant: the bias gradient is analytically zero, both sides hold rounding noise only
    assert float(b.grad.abs().max()) < 1e-5 * float(w.grad.abs().max()) and float(b_r.grad.abs().max()) < 1e-5 * float(w_r.grad.abs().max())
