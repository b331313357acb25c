"""Stage-2 physical compaction, host side (uvc_b200/compact.py): the layout compiler and the compact checkpoint must give
exactly the masked-dense forward the reference computes in Stage 2 (`weight.data *= mask` + hard block skip, post_train.py:228-231,
models/model_distilled.py:496-500).  On CPU the runner's operators are swapped for plain-torch ones with the same signatures (the index
bookkeeping is what is under test here); the GPU test runs the same runner on the sm_100a operators against the whole-model engine."""
import pytest
import torch
import torch.nn.functional as F

from oracle import fixtures as fx, vit_oracle as vo
from uvc_b200 import compact as cp


class TorchOps:
    """uvc_b200.ops look-alike (only what CompactViT calls)."""
    EPI_GELU, EPI_ROUND_TF32 = 2, 32

    @staticmethod
    def round_tf32(t):
        return t

    @staticmethod
    def im2col16(x, patch=16, round_tf32=False):
        return F.unfold(x, kernel_size=patch, stride=patch).transpose(1, 2).reshape(-1, x.shape[1] * patch * patch)

    @staticmethod
    def linear(x, w, bias=None, out=None, flags=0, R=None):
        y = F.linear(x, w, bias)
        if flags & TorchOps.EPI_GELU:
            y = F.gelu(y)
        return y + R if R is not None else y

    @staticmethod
    def layernorm_fwd(x, gamma, beta, eps, y=None, ldx=None, M=None, save_stats=True, round_tf32=False):
        C_ = gamma.numel()
        if ldx is not None and ldx != C_:
            x = x.reshape(-1)[: (M - 1) * ldx + C_].as_strided((M, C_), (ldx, 1))
        return F.layer_norm(x.reshape(-1, C_), (C_,), gamma, beta, eps), None, None

    @staticmethod
    def assemble_tokens(pe, cls, pos, pscale=None, tmask=None):
        if pscale is not None:
            pe = pe * pscale.view(1, -1, 1)
        B = pe.shape[0]
        return torch.cat([cls.view(1, 1, -1).expand(B, -1, -1), pe], 1) + pos.unsqueeze(0)

    @staticmethod
    def attention_fwd(qkv, B, H, N, d, scale=None, save_P=True):
        q, k, v = qkv.view(B, N, 3, H, d).permute(2, 0, 3, 1, 4)
        p = ((q @ k.transpose(-2, -1)) * d ** -0.5).softmax(-1)
        return (p @ v).transpose(1, 2).reshape(B * N, H * d), None


def pruned_checkpoint(model_type="deit_tiny_patch16_224", depth=4, seed=3):
    sd, dims = fx.make_state_dict(model_type, depth, seed=seed)
    H, C_ = dims["num_heads"], dims["embed_dim"]
    Fh = 4 * C_
    g = fx._gen(seed, "prune")
    for l in range(depth):
        pre = f"blocks.{l}."
        m1, m3, m2 = torch.ones(C_, C_), torch.ones(C_, Fh), torch.ones(Fh, C_)
        if l == 0:       # one whole head, 16 dims inside another head, 301 neurons
            m1[:, 64:128] = 0
            m1[:, torch.randperm(64, generator=g)[:16]] = 0
            dead = torch.randperm(Fh, generator=g)[:301]
            m3[:, dead] = 0; m2[dead, :] = 0
        if l == 2:       # degenerate block: every head and every neuron pruned (only the biases survive)
            m1[:] = 0; m3[:] = 0; m2[:] = 0
        sd[pre + "attn.proj.mask"], sd[pre + "mlp.fc2.mask"], sd[pre + "mlp.fc1.mask"] = m1, m3, m2
    sd["block_skip_gating"][1] = torch.tensor([1.0, -1.0])          # block 1 hard-skipped
    return sd, dims


def masked_dense(sd):
    out = dict(sd)
    for k in list(sd):
        if k.endswith(".mask"):
            out[k[:-4] + "weight"] = sd[k[:-4] + "weight"] * sd[k]
    return out


def test_layout_and_compact_forward_equal_masked_dense():
    sd, dims = pruned_checkpoint()
    H = dims["num_heads"]
    lay = cp.compile_layout(sd, H)
    assert lay["blocks"][1] is None and lay["blocks"][0]["heads"] == [0, 2] and lay["blocks"][0]["dims"] == [48, 64]
    assert lay["blocks"][0]["neurons"].numel() == 768 - 301 and lay["blocks"][2]["heads"] == [] and lay["blocks"][2]["neurons"].numel() == 0
    assert lay["blocks"][3]["heads"] == [0, 1, 2] and lay["blocks"][3]["neurons"].numel() == 768
    comp = cp.compact_state_dict(sd, lay)
    csd = comp["state_dict"]
    assert csd["blocks.0.attn.qkv.weight"].shape == (3 * 2 * 64, 192) and csd["blocks.0.attn.proj.weight"].shape == (192, 128)
    assert csd["blocks.0.mlp.fc1.weight"].shape[0] == csd["blocks.0.mlp.fc2.weight"].shape[1] == 472      # 467 live, padded to a multiple of 8
    assert not any(k.startswith("blocks.1.") for k in csd)
    m = cp.macs(lay)
    assert 0.40 < m["ratio"] < 0.50                                          # 2 of 4 blocks gone or empty, one thinned
    x, _ = fx.make_batch(3, seed=11)
    with torch.no_grad():
        want = vo.forward(masked_dense(sd), x, 4, H, skip=[False, True, False, False])
        got = cp.CompactViT(comp, backend=TorchOps)(x)
    assert (got - want).abs().max() <= 2e-5 * want.abs().max()


def test_layout_reproduces_the_reference_budget_of_the_fixed_50_percent_layout():
    """Known answer (BASELINE.json configs[3]): DeiT-Base with blocks 8 and 10 skipped and, in the live blocks, 3 heads, 16
    dimensions of every surviving head and 1417 neurons pruned evaluates to rho = 0.5001 under the reference's calc_flops with the deterministic
    gate.  Only masks, gates and shapes are needed for the layout."""
    C_, H, Fh, L = 768, 12, 3072, 12
    sd = {"block_skip_gating": torch.tensor([[-1.0, 1.0]] * L), "blocks.0.attn.proj.weight": torch.empty(C_, C_, device="meta"),
          "blocks.0.mlp.fc1.weight": torch.empty(Fh, C_, device="meta")}
    sd["block_skip_gating"][8] = torch.tensor([1.0, -1.0]); sd["block_skip_gating"][10] = torch.tensor([1.0, -1.0])
    for l in range(L):
        m1, m3 = torch.ones(1, C_, dtype=torch.uint8), torch.ones(1, Fh, dtype=torch.uint8)       # a mask column is live iff any entry is non-zero
        m1[:, : 3 * 64] = 0
        for h in range(3, H):
            m1[:, h * 64: h * 64 + 16] = 0
        m3[:, :1417] = 0
        sd[f"blocks.{l}.attn.proj.mask"], sd[f"blocks.{l}.mlp.fc2.mask"] = m1, m3
    lay = cp.compile_layout(sd, H)
    assert sum(b is None for b in lay["blocks"]) == 2 and lay["blocks"][0]["heads"] == list(range(3, 12)) and lay["blocks"][0]["dims"] == [48] * 9
    m = cp.macs(lay)
    assert abs(m["budget_ratio"] - 0.5001) < 1e-4
    assert m["budget_ratio"] < m["ratio"] < 0.53                                   # the runner keeps the 16 zeroed dims per head


def test_unpruned_checkpoint_compacts_to_itself():
    sd, dims = fx.make_state_dict("deit_tiny_patch16_224", 2, seed=5)
    lay = cp.compile_layout(sd, dims["num_heads"])                            # no masks, default gates: nothing to remove
    comp = cp.compact_state_dict(sd, lay)
    assert cp.macs(lay)["ratio"] == 1.0
    for k, v in comp["state_dict"].items():
        assert torch.equal(v, sd[k]), k
    x, _ = fx.make_batch(2, seed=12)
    with torch.no_grad():
        assert torch.allclose(cp.CompactViT(comp, backend=TorchOps)(x), vo.forward(sd, x, 2, dims["num_heads"]), atol=2e-5)


def test_export_cli_round_trip(tmp_path, capsys):
    sd, dims = pruned_checkpoint()
    src, dst = str(tmp_path / "stage2.pth.tar"), str(tmp_path / "compact.pt")
    torch.save({"model": {"module." + k: v for k, v in sd.items()}}, src)             # DDP-prefixed, wrapped: what the loops can write
    cp.main(["--checkpoint", src, "--out", dst, "--model_type", "deit_tiny_patch16_224"])
    assert "compact" in capsys.readouterr().out
    comp = torch.load(dst, map_location="cpu", weights_only=False)
    x, _ = fx.make_batch(2, seed=14)
    with torch.no_grad():
        want = vo.forward(masked_dense(sd), x, 4, dims["num_heads"], skip=[False, True, False, False])
        got = cp.CompactViT(comp, backend=TorchOps)(x)
    assert (got - want).abs().max() <= 2e-5 * want.abs().max()


@pytest.mark.gpu
def test_compact_runner_on_gpu_matches_engine_masked_dense():
    """Same runner on the sm_100a operators vs the whole-model engine on the masked-dense checkpoint (DeiT-Small width, 4 blocks)."""
    from functools import partial
    from uvc_b200.models.model_distilled import DistilledVisionTransformer
    sd, dims = pruned_checkpoint("deit_small_patch16_224", 4, seed=4)
    H = dims["num_heads"]
    comp = cp.compact_state_dict(sd, cp.compile_layout(sd, H))
    x, _ = fx.make_batch(4, seed=13)
    got = cp.CompactViT(comp).cuda()(x.cuda())
    with torch.no_grad():
        want = vo.forward(masked_dense(sd), x, 4, H, skip=[False, True, False, False])
    m = DistilledVisionTransformer(enable_dist=0, patch_size=16, mlp_ratio=4, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6),
                                   drop_rate=0, embed_dim=dims["embed_dim"], depth=4, num_heads=H)
    m.load_state_dict({k: v for k, v in masked_dense(sd).items() if not k.endswith(".mask")}, strict=False)
    with torch.no_grad():
        eng, _ = m.cuda().eval()(x.cuda())
    rel = lambda a, b: float((a.cpu() - b.cpu()).abs().max() / b.abs().max())
    assert rel(got, want) < 1e-3 and rel(got, eng) < 1e-3


@pytest.mark.gpu
def test_post_train_compact_eval_adapter_matches_the_model():
    """post_train's `--compact_eval` path: the adapter built from a live model (its registered `.mask` buffers and gates) returns the model's
    own eval logits (masks applied, block 1 skipped)."""
    from functools import partial
    from uvc_b200.models.model_distilled import DistilledVisionTransformer
    from uvc_b200.post_train import CompactEval, apply_masks
    sd, dims = pruned_checkpoint("deit_tiny_patch16_224", 4, seed=6)
    m = DistilledVisionTransformer(enable_dist=0, patch_size=16, mlp_ratio=4, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6),
                                   drop_rate=0, embed_dim=dims["embed_dim"], depth=4, num_heads=dims["num_heads"])
    for _, mod in m.named_modules():
        if hasattr(mod, "weight"):
            mod.register_buffer("mask", torch.ones_like(mod.weight))
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected
    m = m.cuda().eval()
    apply_masks(m)
    x, _ = fx.make_batch(5, seed=15)
    with torch.no_grad():
        want, _ = m(x.cuda())
        got, _ = CompactEval(m)(x.cuda(), -1, 0.9)
    assert float((got - want).abs().max() / want.abs().max()) < 1e-3


def test_clip_norm_closed_form_for_pruned_neurons():
    """Groundwork for the training side of the compaction: the reference clips over the gradients of MASKED weights too, and
    those are not zero.  For a pruned neuron n (fc1 row and fc2 column masked) they have a closed form that needs nothing from the pruned part:
    d fc2.weight[:, n] = gelu(fc1.bias[n]) * colsum(dY) (= gelu(b1[n]) * d fc2.bias) and d fc1.weight[n, :] = 0, d fc1.bias[n] = 0."""
    torch.manual_seed(0)
    M, C_, Fh = 37, 16, 40
    x = torch.randn(M, C_)
    W2, b1 = torch.randn(Fh, C_, requires_grad=True), torch.randn(Fh, requires_grad=True)
    W3, b2 = torch.randn(C_, Fh, requires_grad=True), torch.randn(C_, requires_grad=True)
    dead = torch.tensor([1, 5, 6, 22, 39])
    m2, m3 = torch.ones(Fh, C_), torch.ones(C_, Fh)
    m2[dead] = 0; m3[:, dead] = 0
    with torch.no_grad():                                   # `weight.data *= mask` (post_train.py:357-360): the parameters themselves are zeroed
        W2.mul_(m2); W3.mul_(m3)
    y = F.linear(F.gelu(F.linear(x, W2, b1)), W3, b2)
    r = torch.randn(M, C_)
    (y * r).sum().backward()
    want_sq = float((W3.grad[:, dead] ** 2).sum() + (W2.grad[dead] ** 2).sum() + (b1.grad[dead] ** 2).sum())
    colsum_dy = b2.grad                                     # = colsum(dY): computed by the backward anyway
    closed_sq = float((colsum_dy ** 2).sum() * (F.gelu(b1.detach()[dead]) ** 2).sum())
    assert abs(want_sq - closed_sq) <= 1e-5 * want_sq
    assert float(W2.grad[dead].abs().max()) == 0.0 and float(b1.grad[dead].abs().max()) == 0.0
    assert torch.allclose(W3.grad[:, dead], colsum_dy[:, None] * F.gelu(b1.detach()[dead])[None, :], rtol=1e-5, atol=1e-6)


def _tiny_sd(seed, depth=2, C_=128, H=2):
    """A 2-head, 128-wide model in the reference's key layout (fixtures.param_shapes), small enough for property tests."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in fx.param_shapes(embed_dim=C_, depth=depth, num_heads=H, mlp_ratio=1, num_classes=10).items():
        if k.endswith("skip_gating"):
            sd[k] = torch.tensor([-1.0, 1.0]).expand(shp).contiguous()
        elif "norm" in k and k.endswith("weight"):
            sd[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            sd[k] = torch.randn(shp, generator=g) * (0.05 if k.endswith("weight") else 0.02)
    return sd


def test_random_layouts_compact_equals_masked_dense():
    """Property: for ANY pattern of pruned proj columns / fc2 columns / fc1 rows and skipped blocks, the compact forward equals the masked-dense one."""
    from hypothesis import given, settings, strategies as st
    depth, C_, H, Fh = 2, 128, 2, 128

    @settings(max_examples=12, deadline=None)
    @given(seed=st.integers(0, 10 ** 6), p_col=st.floats(0.0, 1.0), p_neu=st.floats(0.0, 1.0), skip=st.lists(st.booleans(), min_size=depth, max_size=depth),
           kill_head=st.lists(st.booleans(), min_size=depth * H, max_size=depth * H))
    def run(seed, p_col, p_neu, skip, kill_head):
        sd = _tiny_sd(seed, depth, C_, H)
        g = torch.Generator().manual_seed(seed + 1)
        for l in range(depth):
            pre = f"blocks.{l}."
            keep1 = (torch.rand(C_, generator=g) >= p_col * 0.7).float()
            for h in range(H):
                if kill_head[l * H + h]:
                    keep1[h * 64:(h + 1) * 64] = 0
            keep3 = (torch.rand(Fh, generator=g) >= p_neu * 0.9).float()
            sd[pre + "attn.proj.mask"] = keep1.expand(C_, C_).contiguous()
            sd[pre + "mlp.fc2.mask"] = keep3.expand(C_, Fh).contiguous()
            sd[pre + "mlp.fc1.mask"] = keep3.view(Fh, 1).expand(Fh, C_).contiguous()
            if skip[l]:
                sd["block_skip_gating"][l] = torch.tensor([0.3, 0.3])
        lay = cp.compile_layout(sd, H)
        x = torch.randn(2, 3, 224, 224, generator=g)
        with torch.no_grad():
            want = vo.forward(masked_dense(sd), x, depth, H, skip=list(skip))
            got = cp.CompactViT(cp.compact_state_dict(sd, lay), backend=TorchOps)(x)
        assert (got - want).abs().max() <= 3e-5 * want.abs().max().clamp_min(1e-6)
        m = cp.macs(lay)
        assert 0.0 < m["budget_ratio"] <= m["ratio"] + 1e-12 <= 1.0 + 1e-12
    run()


def test_engine_layout_rules_on_cpu():
    """`EngineLayout` (the uvc_vit_layout the engine trains on): every executed block keeps >= 1 head and a multiple of 64 (>= 64) neurons, topped up
    with pruned entries; each index row is a permutation with the live entries first, ascending; hard-skipped blocks are left dense (never read)."""
    sd, dims = pruned_checkpoint()
    H = dims["num_heads"]
    lay = cp.compile_layout(sd, H)
    for keep in (False, True):
        el = cp.EngineLayout(lay, "cpu", keep_pruned_heads=keep)
        Fh = lay["Fh"]
        for l, b in enumerate(lay["blocks"]):
            nh, nn_ = el.struct.n_heads[l], el.struct.n_neurons[l]
            assert 1 <= nh <= H and 64 <= nn_ <= Fh and nn_ % 64 == 0
            hrow, nrow = el.head_idx[l].tolist(), el.neuron_idx[l].tolist()
            assert sorted(hrow) == list(range(H)) and sorted(nrow) == list(range(Fh))           # permutations
            assert hrow[:nh] == sorted(hrow[:nh]) and nrow[:nn_] == sorted(nrow[:nn_])            # live part ascending
            if b is None:
                assert nh == H and nn_ == Fh
                continue
            assert set(b["heads"]) <= set(hrow[:nh]) and set(b["neurons"].tolist()) <= set(nrow[:nn_])      # nothing live is dropped
            if keep:
                assert nh == H
            else:
                assert nh == max(1, len(b["heads"]))
            assert nn_ - b["neurons"].numel() < 64 or b["neurons"].numel() == 0
        assert 0 < el.executed_macs_ratio() <= 1.0
    assert cp.EngineLayout(lay, "cpu", keep_pruned_heads=True).executed_macs_ratio() >= cp.EngineLayout(lay, "cpu").executed_macs_ratio()
