"""Stage-2 physical compaction, host side (reference: post_train.py:228-231,357-360 re-mask dense weights every batch and
never realises the FLOPs it pruned; the hard block skip is models/model_distilled.py:496-500).

    layout   = compile_layout(state_dict, num_heads)          masks + gates -> live blocks / heads / neurons (index lists)
    compact  = compact_state_dict(state_dict, layout)         physically smaller tensors (ragged per block) + the layout
    model    = CompactViT(compact).to("cuda");  logits = model(x)     inference through the same sm_100a kernels (uvc_b200.ops)

What is removed is exactly what cannot influence the output (so the compact forward equals the masked-dense forward up to summation order):
  * a block whose gate says skip (`gate[1] <= gate[0]`, the reference's Stage-2 rule);
  * a head all of whose 64 input columns of `attn.proj.weight` are masked: its q, k, v rows, its attention and its proj columns;
  * a neuron whose column of `mlp.fc2.weight` is masked: the fc1 row (and bias entry) and the fc2 column.
Dimensions pruned INSIDE a surviving head (the `r` variables) keep their zeroed proj columns: the fused attention kernel works on 64-wide
heads.  `macs()` reports what the layout saves.

Training on the compacted model: `EngineLayout` / `engine_layout_for(model)` build the `uvc_vit_layout` the whole-model engine takes
(include/uvc_b200.h; csrc/compact.cu gathers live rows / columns while converting the weights and scatters the compact gradients back);
`post_train.py --compact_train {1,2}` switches it on.  What the gradient of a masked weight is under compaction -- the reference's
`clip_grad_norm_` counts it -- is spelled out in DESIGN.md section 5.7 and tested in tests/test_compact_train_gpu.py.
"""
import torch

from . import ops as _ops

HEAD_DIM = 64


def _live_cols(mask):
    return (mask != 0).any(dim=0)


def compile_layout(sd, num_heads, head_dim=HEAD_DIM):
    """-> {"L", "C", "H", "Fh", "blocks": [None (skipped) | {"heads": [h...], "dims": [live dims per kept head], "neurons": LongTensor}]}"""
    L = sd["block_skip_gating"].shape[0]
    C_ = sd["blocks.0.attn.proj.weight"].shape[0]
    Fh = sd["blocks.0.mlp.fc1.weight"].shape[0]
    assert C_ == num_heads * head_dim
    gate = sd["block_skip_gating"]
    blocks = []
    for l in range(L):
        if not bool(gate[l, 1] > gate[l, 0]):                 # models/model_distilled.py:498
            blocks.append(None)
            continue
        pre = f"blocks.{l}."
        m1 = sd.get(pre + "attn.proj.mask")
        live1 = _live_cols(m1) if m1 is not None else torch.ones(C_, dtype=torch.bool)
        per_head = live1.view(num_heads, head_dim)
        heads = [h for h in range(num_heads) if bool(per_head[h].any())]
        m3 = sd.get(pre + "mlp.fc2.mask")
        live3 = _live_cols(m3) if m3 is not None else torch.ones(Fh, dtype=torch.bool)
        blocks.append({"heads": heads, "dims": [int(per_head[h].sum()) for h in heads], "neurons": torch.nonzero(live3).flatten()})
    return {"L": L, "C": C_, "H": num_heads, "Fh": Fh, "head_dim": head_dim, "blocks": blocks}


def _pad_index(idx, total, multiple=8):
    """The GEMM operands need row strides that are multiples of 4 floats: top the live set up with a few pruned entries.  Their masked weights
    are zero (fc2 column), so they change nothing."""
    extra = (-idx.numel()) % multiple
    if extra == 0 or idx.numel() == 0:
        return idx
    dead = torch.ones(total, dtype=torch.bool)
    dead[idx] = False
    pad = torch.nonzero(dead).flatten()[:extra]
    return torch.sort(torch.cat([idx, pad]))[0]


class EngineLayout:
    """`uvc_vit_layout` for the whole-model engine (include/uvc_b200.h): Stage-2 TRAINING on the physically compacted model.  Built from a
    `compile_layout` result; owns the device index arrays the C struct points to.

    Rules the engine asks for: every executed block keeps >= 1 head and a multiple of 64 (>= 64) neurons -- the lists are topped up with
    pruned entries, whose masked weights are zero and therefore change nothing (post_train.py:357-360 keeps them zero).  Each row of
    `neuron_idx` / `head_idx` is a permutation: the live entries first (ascending), then the pruned ones.
    `keep_pruned_heads=True` keeps all heads (only blocks and neurons are compacted): the exact-clip-norm mode, see DESIGN.md."""

    def __init__(self, layout, device, keep_pruned_heads=False):
        from ._lib import VitLayout
        L, H, Fh = layout["L"], layout["H"], layout["Fh"]
        head_idx = torch.zeros(L, H, dtype=torch.int32)
        neuron_idx = torch.zeros(L, Fh, dtype=torch.int32)
        st = VitLayout()
        self.heads, self.neurons = [], []
        for l, b in enumerate(layout["blocks"]):
            if b is None:                                   # hard-skipped: never looked at
                heads, neurons = list(range(H)), torch.arange(Fh)
            else:
                heads = list(range(H)) if keep_pruned_heads else (list(b["heads"]) or [0])
                neurons = b["neurons"] if b["neurons"].numel() else torch.arange(0)
                neurons = _pad_index(neurons, Fh, 64) if neurons.numel() else torch.arange(min(64, Fh))
            dead_h = [h for h in range(H) if h not in set(heads)]
            head_idx[l] = torch.tensor(heads + dead_h, dtype=torch.int32)
            live = torch.zeros(Fh, dtype=torch.bool); live[neurons] = True
            neuron_idx[l] = torch.cat([torch.nonzero(live).flatten(), torch.nonzero(~live).flatten()]).to(torch.int32)
            st.n_heads[l], st.n_neurons[l] = len(heads), int(neurons.numel())
            self.heads.append(heads); self.neurons.append(neurons)
        self.head_idx, self.neuron_idx = head_idx.to(device).contiguous(), neuron_idx.to(device).contiguous()
        st.head_idx, st.neuron_idx = self.head_idx.data_ptr(), self.neuron_idx.data_ptr()
        self.struct, self.layout, self.keep_pruned_heads = st, layout, keep_pruned_heads

    def executed_macs_ratio(self, n_tokens=197):
        """MACs the compacted engine executes per image / dense MACs (heads at full width, neurons as padded)."""
        lay = self.layout
        C_, H, Fh, d, N = lay["C"], lay["H"], lay["Fh"], lay["head_dim"], n_tokens
        embed = (N - 1) * 768 * C_
        dense = embed + lay["L"] * (N * C_ * 3 * C_ + 2 * H * N * N * d + N * C_ * C_ + 2 * N * C_ * Fh)
        comp = embed
        for b, hs, ns in zip(lay["blocks"], self.heads, self.neurons):
            if b is not None:
                h, f = len(hs), int(ns.numel())
                comp += N * C_ * 3 * h * d + 2 * h * N * N * d + N * h * d * C_ + 2 * N * C_ * f
        return comp / dense


def engine_layout_for(model, keep_pruned_heads=False):
    """Masks + gates of a live Stage-2 model -> EngineLayout on the model's device."""
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()
          if k.endswith(".mask") or k == "block_skip_gating" or k in ("blocks.0.attn.proj.weight", "blocks.0.mlp.fc1.weight")}
    return EngineLayout(compile_layout(sd, model.blocks[0].attn.num_heads), model.cls_token.device, keep_pruned_heads)


def _masked(sd, key):
    w = sd[key]
    m = sd.get(key[:-len("weight")] + "mask") if key.endswith("weight") else None
    return w * m if m is not None else w                       # what the reference computes with after `weight.data *= mask`


def compact_state_dict(sd, layout):
    """Physically smaller checkpoint: global tensors unchanged, live blocks keep their original index in the key."""
    d, C_ = layout["head_dim"], layout["C"]
    out = {k: _masked(sd, k).clone() for k in ("cls_token", "pos_embed", "patch_embed.proj.weight", "patch_embed.proj.bias", "norm.weight",
                                                "norm.bias", "head.weight", "head.bias")}
    for l, b in enumerate(layout["blocks"]):
        if b is None:
            continue
        pre = f"blocks.{l}."
        for k in ("norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias", "attn.proj.bias", "mlp.fc2.bias"):
            out[pre + k] = sd[pre + k].clone()
        cols = torch.cat([torch.arange(h * d, (h + 1) * d) for h in b["heads"]]) if b["heads"] else torch.zeros(0, dtype=torch.long)
        rows = torch.cat([cols, C_ + cols, 2 * C_ + cols])
        out[pre + "attn.qkv.weight"] = _masked(sd, pre + "attn.qkv.weight")[rows].clone()
        qb = sd.get(pre + "attn.qkv.bias")
        if qb is not None:
            out[pre + "attn.qkv.bias"] = qb[rows].clone()
        out[pre + "attn.proj.weight"] = _masked(sd, pre + "attn.proj.weight")[:, cols].clone()
        n = _pad_index(b["neurons"], layout["Fh"])
        out[pre + "mlp.fc1.weight"] = _masked(sd, pre + "mlp.fc1.weight")[n].clone()
        out[pre + "mlp.fc1.bias"] = sd[pre + "mlp.fc1.bias"][n].clone()
        out[pre + "mlp.fc2.weight"] = _masked(sd, pre + "mlp.fc2.weight")[:, n].clone()
    return {"layout": layout, "state_dict": out}


def macs(layout, n_tokens=197, patch_macs=None):
    """Per-image MACs of the dense model and of the compact one (the reference's accounting, models/model_distilled.py:115-189)."""
    C_, H, Fh, d, N = layout["C"], layout["H"], layout["Fh"], layout["head_dim"], n_tokens
    embed = (N - 1) * 768 * C_ if patch_macs is None else patch_macs
    dense_block = N * C_ * 3 * C_ + 2 * H * N * N * d + N * C_ * C_ + 2 * N * C_ * Fh
    dense = embed + layout["L"] * dense_block
    comp = embed
    for b in layout["blocks"]:
        if b is None:
            continue
        h, f = len(b["heads"]), int(b["neurons"].numel())
        comp += N * C_ * 3 * h * d + 2 * h * N * N * d + N * h * d * C_ + 2 * N * C_ * f
    # the reference's own resource model (uvc_utils.py:409-471) with the deterministic Stage-2 gate: it also credits
    # the dimensions pruned INSIDE surviving heads (rho_r), which the compact runner does not cut yet
    budget = embed
    m01, m23, m45 = N * C_ * 3 * C_ + H * N * N * d, H * N * N * d + N * C_ * C_, 2 * N * C_ * Fh
    for b in layout["blocks"]:
        if b is None:
            continue
        rho0 = len(b["heads"]) / H
        rho_r = sum(b["dims"]) / C_
        rho1 = int(b["neurons"].numel()) / Fh
        budget += m01 * rho0 + m23 * rho_r + m45 * rho1
    return {"dense": dense, "compact": comp, "ratio": comp / dense, "budget_ratio": budget / dense}


class CompactViT(torch.nn.Module):
    """Inference forward of a compacted checkpoint, one uvc_b200.ops call per operator (the whole-model engine entry point takes one H and one
    Fh for all blocks; ragged blocks go through the per-operator entry points of the same library).  `backend` exists for the CPU test, which
    substitutes plain-torch operators with the same signatures to check the index bookkeeping where there is no GPU."""

    def __init__(self, compact, eps=1e-6, patch=16, backend=None):
        super().__init__()
        self.layout, self.eps, self.patch = compact["layout"], float(eps), int(patch)
        self.ops = backend if backend is not None else _ops
        self.names = sorted(compact["state_dict"])
        for i, k in enumerate(self.names):
            self.register_buffer(f"t{i}", compact["state_dict"][k].detach().float().contiguous())
        self._rounded = None

    def _tensors(self):
        """GEMM weights rounded to TF32 once (the tensor core truncates its operands)."""
        dev = self.t0.device
        if self._rounded is None or self._rounded[0] != dev:
            sd = {}
            for i, k in enumerate(self.names):
                t = getattr(self, f"t{i}")
                if k.endswith("weight") and t.dim() >= 2 and t.numel():
                    t = self.ops.round_tf32(t.reshape(t.shape[0], -1).contiguous())
                sd[k] = t
            self._rounded = (dev, sd)
        return self._rounded[1]

    @torch.no_grad()
    def forward(self, x, patch_scale=None):
        o, sd, lay = self.ops, self._tensors(), self.layout
        B, C_, d = x.shape[0], lay["C"], lay["head_dim"]
        cols = o.im2col16(x.contiguous().float(), self.patch, round_tf32=True)
        pe = o.linear(cols, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"])
        n_p = pe.shape[0] // B
        tok = o.assemble_tokens(pe.view(B, n_p, C_), sd["cls_token"].reshape(-1), sd["pos_embed"].reshape(-1, C_), patch_scale, None)
        N = n_p + 1
        xs = tok.view(B * N, C_)
        for l, b in enumerate(lay["blocks"]):
            if b is None:
                continue
            pre = f"blocks.{l}."
            h = len(b["heads"])
            if h:
                ln1, _, _ = o.layernorm_fwd(xs, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], self.eps, save_stats=False, round_tf32=True)
                qkv = o.linear(ln1, sd[pre + "attn.qkv.weight"], sd.get(pre + "attn.qkv.bias"), flags=o.EPI_ROUND_TF32)
                ctx, _ = o.attention_fwd(qkv, B, h, N, d, save_P=False)
                x1 = o.linear(ctx, sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"], R=xs)
            else:
                x1 = xs + sd[pre + "attn.proj.bias"]
            if b["neurons"].numel():
                ln2, _, _ = o.layernorm_fwd(x1, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], self.eps, save_stats=False, round_tf32=True)
                hid = o.linear(ln2, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"], flags=o.EPI_GELU | o.EPI_ROUND_TF32)
                xs = o.linear(hid, sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"], R=x1)
            else:
                xs = x1 + sd[pre + "mlp.fc2.bias"]
        cls_ln, _, _ = o.layernorm_fwd(xs, sd["norm.weight"], sd["norm.bias"], self.eps, ldx=N * C_, M=B, save_stats=False, round_tf32=True)
        return o.linear(cls_ln, sd["head.weight"], sd["head.bias"])


def export(checkpoint_path, out_path, num_heads):
    """Stage-1 / Stage-2 checkpoint (state dict with `.mask` buffers and `block_skip_gating`) -> compact checkpoint file.  Returns the MAC summary."""
    ck = torch.load(checkpoint_path, map_location="cpu")
    for key in ("model", "state_dict_ema", "state_dict"):
        if isinstance(ck, dict) and key in ck and isinstance(ck[key], dict):
            ck = ck[key]
            break
    ck = {(k[7:] if k.startswith("module.") else k): v for k, v in ck.items()}
    layout = compile_layout(ck, num_heads)
    torch.save(compact_state_dict(ck, layout), out_path)
    return macs(layout)


def main(argv=None):
    import argparse
    from .models import CONFIGS
    ap = argparse.ArgumentParser(description="Export a physically compacted checkpoint from a UVC checkpoint (masks + gates).")
    ap.add_argument("--checkpoint", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--model_type", default="deit_small_patch16_224", choices=list(CONFIGS.keys()))
    a = ap.parse_args(argv)
    m = export(a.checkpoint, a.out, CONFIGS[a.model_type].num_heads)
    print(f"dense {m['dense'] / 1e6:.1f} M MACs/image -> compact {m['compact'] / 1e6:.1f} M ({m['ratio'] * 100:.1f} %); wrote {a.out}")


if __name__ == "__main__":
    main()
