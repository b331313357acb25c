"""State and math of UVC's compression problem, with the per-step work on the device.

Mirror of the reference's `UVC/uvc_utils.py` (same public names and call signatures: `UVC_CP_MiniMax`,
`weight_list_to_scores`, `prox_w`, `prune_w`, `prune_w_mask`, `proj_dual`, `calc_flops`, `ste_ceil`,
`ste_floor`, `PresetLRScheduler`).  The reference walks Python lists of layers and heads and synchronises
with the device for every group norm (~65k syncs per step on DeiT-Base); here one call of
`uvc_admm_scores` produces the norms and the in-group RANKS of all layers, and selections become
`rank < k` inside the prox / mask / primal / dual kernels (include/uvc_b200.h, csrc/admm.cu).
"""
import ctypes as C

import torch
from torch import nn
from torch.nn import Parameter

from . import _lib
from ._lib import AdmmArgs


# ------------------------------------------------------------------------------------------ straight-through rounding
class SteFloor(torch.autograd.Function):
    """floor() with an identity gradient (reference uvc_utils.py:26-38)."""

    @staticmethod
    def forward(ctx, a):
        return torch.floor(a)

    @staticmethod
    def backward(ctx, g):
        return g


class SteCeil(torch.autograd.Function):
    """ceil() with an identity gradient (reference uvc_utils.py:40-52)."""

    @staticmethod
    def forward(ctx, a):
        return torch.ceil(a)

    @staticmethod
    def backward(ctx, g):
        return g


ste_floor = SteFloor.apply
ste_ceil = SteCeil.apply


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _AdmmDevice:
    """Device workspace + pointer tables for the ADMM kernels over a fixed set of prunable layers."""

    def __init__(self, uvc_layers, head_size):
        self.W1, self.W2, self.W3 = uvc_layers["W1"], uvc_layers["W2"], uvc_layers["W3"]
        self.L = len(self.W1)
        self.d = int(head_size)
        self.C = self.W1[0].in_features
        self.H = self.C // self.d
        self.Fh = self.W3[0].in_features
        self._sig = None

    def _refresh(self):
        ws = [m.weight for m in self.W1 + self.W3]
        sig = tuple(w.data_ptr() for w in ws)
        if sig == self._sig:
            return
        dev = ws[0].device
        if dev.type != "cuda":
            raise _lib.UvcError("the ADMM kernels run on CUDA weights only (no CPU fallback)")
        for w in ws:
            if not (w.dtype == torch.float32 and w.is_contiguous()):
                raise _lib.UvcError("ADMM weights must be contiguous fp32")
        L, Cc, H, Fh = self.L, self.C, self.H, self.Fh
        self.c1 = torch.empty(L, Cc, device=dev); self.c2 = torch.empty(L, H, device=dev); self.c3 = torch.empty(L, Fh, device=dev)
        self.rank1 = torch.empty(L, Cc, device=dev, dtype=torch.int32)
        self.rank2 = torch.empty(L, H, device=dev, dtype=torch.int32)
        self.rank3 = torch.empty(L, Fh, device=dev, dtype=torch.int32)
        self.out = torch.zeros(1, device=dev)
        self.w1_tab = (C.c_void_p * L)(*[m.weight.data_ptr() for m in self.W1])
        self.w3_tab = (C.c_void_p * L)(*[m.weight.data_ptr() for m in self.W3])
        self._sig = sig

    def args(self):
        self._refresh()
        a = AdmmArgs()
        a.L, a.H, a.d, a.Fh = self.L, self.H, self.d, self.Fh
        a.w1 = C.cast(self.w1_tab, AdmmArgs._PP); a.w3 = C.cast(self.w3_tab, AdmmArgs._PP)
        a.c1, a.c2, a.c3 = self.c1.data_ptr(), self.c2.data_ptr(), self.c3.data_ptr()
        a.rank1, a.rank2, a.rank3 = self.rank1.data_ptr(), self.rank2.data_ptr(), self.rank3.data_ptr()
        a.out = self.out.data_ptr()
        return a

    def call(self, name, a):
        lib = _lib.load()
        _lib.check(getattr(lib, name)(C.byref(a), _stream()), name)

    def scores(self):
        a = self.args()
        self.call("uvc_admm_scores", a)
        return a


def weight_list_to_scores(layer, layer_group_name, head_size=None):
    """Group norms of one layer (reference uvc_utils.py:54-73): W1 -> ([H, d] column norms, [H] head norms), W3 -> [Fh].
    Returned on the CPU like the reference does."""
    if layer_group_name == "W1":
        dev = _AdmmDevice({"W1": [layer], "W2": [], "W3": [_ScoreOnly(layer)]}, head_size)
        dev.scores()
        return dev.c1.view(dev.H, dev.d).cpu(), dev.c2.view(dev.H).cpu()
    if layer_group_name == "W3":
        dev = _AdmmDevice({"W1": [_ScoreOnly(layer, square=True)], "W2": [], "W3": [layer]}, layer.out_features)
        dev.scores()
        return dev.c3.view(-1).cpu()
    raise ValueError(layer_group_name)


class _ScoreOnly:
    """Placeholder partner layer for single-layer score queries (the kernel always walks a W1 and a W3)."""

    def __init__(self, layer, square=False):
        n = layer.out_features
        if square:
            self.weight = torch.zeros(n, n, device=layer.weight.device)
            self.in_features = n
        else:
            self.weight = torch.zeros(n, 4, device=layer.weight.device)
            self.in_features = 4
        self.out_features = n


class UVC_CP_MiniMax(nn.Module):
    """ADMM variables of the compression problem (reference uvc_utils.py:129-308).

    s[L,2]  #heads / #MLP neurons to remove per block        y[L,2]  their duals
    r[L,H]  #dims to remove inside each head                  p[L,H]  their duals
    z       dual of the FLOPs-budget constraint
    """

    def __init__(self, model, resource_fn, uvc_layers, uvc_layers_dict, head_size, num_heads, flops_list, z_init=1e-3, y_init=1e-3,
                 p_init=1e-3, args=None):
        super().__init__()
        self.model = model
        self.uvc_layers = uvc_layers
        self.uvc_layers_dict = uvc_layers_dict
        self.head_size = head_size
        n_layers = len(self.uvc_layers['W1'])
        self.n_layers = n_layers
        self.eps_decay = args.eps_decay
        dev = model.block_skip_gating.device
        self.s = Parameter(torch.zeros(n_layers, 2, device=dev))
        self.r = Parameter(torch.zeros(n_layers, num_heads, device=dev))
        self.y = Parameter(torch.full((n_layers, 2), float(y_init), device=dev))
        self.p = Parameter(torch.full((n_layers, num_heads), float(p_init), device=dev))
        self.z = Parameter(torch.tensor(float(z_init), device=dev))
        self.resource_fn = resource_fn
        self.enable_patch_gating = args.enable_patch_gating
        n_patches = model.patch_embed.num_patches if hasattr(model, "patch_embed") else model.num_patches     # T2T-ViT: tokens_to_token, no patch conv
        self.patch_gating = Parameter(3 * torch.ones(1, n_patches, 1, device=dev)) if self.enable_patch_gating == 1 else None
        self.update_patch()
        self.enable_part_gating = args.enable_part_gating
        self.enable_block_gating = args.enable_block_gating
        self.update_gating()
        self.s_ub = torch.zeros(n_layers, 2, device=dev)
        self.s_ub[:, 0] = num_heads
        self.s_ub[:, 1] = uvc_layers["W3"][0].in_features
        self.r_ub = torch.full((n_layers, num_heads), float(head_size), device=dev)
        self.num_heads = num_heads
        self.flops_list = flops_list
        self.args = args
        self._dev = _AdmmDevice(uvc_layers, head_size)
        # resource model constants (joint_train.py:1010-1012 probe at batch 1): fp32 like torch.Tensor(total_macs)
        embed_macs, total_macs = flops_list
        self._embed_macs = float(embed_macs)
        self._macs = torch.tensor([[float(v) for v in row] for row in total_macs], dtype=torch.float32)
        self.full_flops = float((torch.tensor(float(embed_macs)) + self._macs.sum()) * 2)       # calc_flops(full_model_flops=None)
        self._macs_dev = None
        self.noise_source = None

    # ---- reference API
    def ceiled_s(self):
        return ste_ceil(self.s)

    def ceiled_r(self):
        return ste_ceil(self.r)

    def update_gating(self):
        self.block_skip_gating = self.model.block_skip_gating if self.enable_block_gating else None
        self.attn_skip_gating = [] if self.enable_part_gating else None
        self.mlp_skip_gating = [] if self.enable_part_gating else None
        if self.enable_part_gating:
            for name, p in self.model.named_parameters():
                if "attn_skip_gating" in name:
                    self.attn_skip_gating.append(p)
                if "mlp_skip_gating" in name:
                    self.mlp_skip_gating.append(p)

    def update_patch(self):
        if self.enable_patch_gating == 1:
            self.model.patch_gating = self.patch_gating

    def update_eps(self):
        if not self.model.enable_warmup:
            print(f"[EPS update] {self.model.eps} =====> {self.model.eps * self.eps_decay} ")
            self.model.eps = self.model.eps * self.eps_decay

    # ---- device plumbing
    def admm_args(self, noise=None, gumbel_hard=False, warmup=False):
        a = self._dev.args()
        dev = self.s.device
        if self._macs_dev is None or self._macs_dev.device != dev:
            self._macs_dev = self._macs.to(dev)
        a.s, a.r, a.y, a.p, a.z = (t.data_ptr() for t in (self.s, self.r, self.y, self.p, self.z))
        gate = self.block_skip_gating
        a.gate = None if gate is None else gate.data_ptr()
        a.noise = None if noise is None else noise.data_ptr()
        a._keepalive = (noise, self._macs_dev)      # the kernel reads these after this function returns
        a.macs = self._macs_dev.data_ptr()
        a.embed_macs, a.full_flops = self._embed_macs, self.full_flops
        ar = self.args
        a.use_gumbel = 1 if getattr(ar, "use_gumbel", 0) else 0
        a.gumbel_hard = 1 if gumbel_hard else 0
        a.warmup = 1 if warmup else 0
        a.eps = float(self.model.eps)
        a.budget = float(getattr(ar, "budget", 0.0))
        a.sl2wd = float(getattr(ar, "sl2wd", 0.0))
        a.gating_weight = float(getattr(ar, "gating_weight", 0.0))
        return a

    def gumbel_noise(self):
        """One [L,2] Gumbel(0,1) draw exactly as F.gumbel_softmax makes it inside calc_flops (uvc_utils.py:446)."""
        gate = self.block_skip_gating
        if gate is None or not getattr(self.args, "use_gumbel", 0):
            return None
        if self.noise_source is not None:        # tests replay the numbers the reference drew
            return self.noise_source().to(gate.device).contiguous()
        return -torch.empty_like(gate, memory_format=torch.legacy_contiguous_format).exponential_().log()

    def run_resource_fn(self, gumbel_hard=False):
        """FLOPs fraction of the dense model for the current ceil(s), ceil(r), gates (reference :219-223)."""
        self._dev.scores()
        noise = self.gumbel_noise()
        a = self.admm_args(noise=noise, gumbel_hard=bool(gumbel_hard))
        self._dev.call("uvc_admm_resource", a)
        return self._dev.out[0].clone()

    def srloss2(self, budget):
        return self.run_resource_fn() - budget

    def zloss(self, budget):
        return self.z * (self.run_resource_fn() - budget)

    def _bottom_sums(self):
        d = self._dev
        d.scores()
        S, R = torch.ceil(self.s.detach()), torch.ceil(self.r.detach())
        b2 = (d.c2 * (d.rank2 < S[:, 0:1])).sum(1)
        b3 = (d.c3 * (d.rank3 < S[:, 1:2])).sum(1)
        b1 = (d.c1.view(d.L, d.H, d.d) * (d.rank1.view(d.L, d.H, d.d) < R.unsqueeze(-1))).sum(-1)
        return b1, b2, b3

    def get_least_s_norm(self):
        _, b2, b3 = self._bottom_sums()
        return torch.stack([b2, b3], dim=1)

    def get_least_r_norm(self):
        return self._bottom_sums()[0]

    def sloss1(self):
        n = self.get_least_s_norm()
        return self.y[:, 0].detach().dot(n[:, 0]) + self.y[:, 1].detach().dot(n[:, 1])

    def rloss1(self):
        return (self.p.detach() * self.get_least_r_norm()).sum()

    def yloss(self):
        n = self.get_least_s_norm()
        return self.y[:, 0].dot(n[:, 0]) + self.y[:, 1].dot(n[:, 1])

    def ploss(self):
        return (self.p * self.get_least_r_norm()).sum()


# ------------------------------------------------------------------------------------------ weight-side operators
def prox_w(minimax_model, optimizer):
    """Proximal shrink of the to-be-pruned columns (reference uvc_utils.py:315-345)."""
    d = minimax_model._dev
    d.scores()
    a = minimax_model.admm_args()
    a.lr = float(optimizer.param_groups[0]['lr'])
    d.call("uvc_admm_prox", a)


def _mask_tables(minimax_model):
    L = minimax_model._dev.L
    tabs = []
    for grp in ("W1", "W3", "W2"):
        ms = [m.mask for m in minimax_model.uvc_layers[grp]]
        for t in ms:
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise _lib.UvcError("mask buffers must be contiguous CUDA fp32 tensors")
        tabs.append((C.c_void_p * L)(*[t.data_ptr() for t in ms]))
    return tabs


def prune_w_mask(minimax_model, optimizer=None):
    """Rewrite the `.mask` buffers of W1 / W3 (columns) and W2 (rows) from the current s, r (reference :376-401)."""
    d = minimax_model._dev
    d.scores()
    a = minimax_model.admm_args()
    t1, t3, t2 = _mask_tables(minimax_model)
    a.m1, a.m3, a.m2 = C.cast(t1, AdmmArgs._PP), C.cast(t3, AdmmArgs._PP), C.cast(t2, AdmmArgs._PP)
    d.call("uvc_admm_masks", a)


def prune_w(minimax_model, optimizer=None):
    """Zero the selected columns / rows of the weights themselves (reference :348-372): masks into scratch, then multiply."""
    d = minimax_model._dev
    d.scores()
    a = minimax_model.admm_args()
    L = d.L
    scratch = {g: [torch.ones_like(m.weight) for m in minimax_model.uvc_layers[g]] for g in ("W1", "W3", "W2")}
    tabs = [(C.c_void_p * L)(*[t.data_ptr() for t in scratch[g]]) for g in ("W1", "W3", "W2")]
    a.m1, a.m3, a.m2 = (C.cast(t, AdmmArgs._PP) for t in tabs)
    d.call("uvc_admm_masks", a)
    with torch.no_grad():
        for g in ("W1", "W3", "W2"):
            for m, k in zip(minimax_model.uvc_layers[g], scratch[g]):
                m.weight.mul_(k)


def proj_dual(minimax_model):
    minimax_model.y.data.clamp_(min=0.0)
    minimax_model.p.data.clamp_(min=0.0)
    minimax_model.z.data.clamp_(min=0.0)


def calc_flops(s, r, uvc_layers_dict, uvc_layers, head_size, s_ub, r_ub, flops_list, gating, eps, full_model_flops=None, use_gumbel=False,
               gumbel_hard=False, args=None):
    """FLOPs model of the reference (uvc_utils.py:409-471), value only.  With full_model_flops=None it returns the dense
    model's FLOPs (2 * MACs); otherwise the fraction left after removing ceil(s) heads / neurons and ceil(r) head dims,
    scaled by the (Gumbel or soft-L0) block gates."""
    embed_macs, total_macs = flops_list
    macs = torch.tensor([[float(v) for v in row] for row in total_macs], dtype=torch.float32)
    if full_model_flops is None:
        return (torch.tensor(float(embed_macs)) + macs.sum()) * 2
    dev = _AdmmDevice(uvc_layers, head_size)
    a = dev.scores()
    device = dev.c1.device
    s_d, r_d = s.detach().to(device).contiguous().float(), r.detach().to(device).contiguous().float()
    macs_d = macs.to(device)
    block_gating = gating[0] if gating is not None else None
    noise = None
    a.s, a.r, a.macs = s_d.data_ptr(), r_d.data_ptr(), macs_d.data_ptr()
    if block_gating is not None:
        a.gate = block_gating.data_ptr()
        if use_gumbel:
            noise = -torch.empty_like(block_gating, memory_format=torch.legacy_contiguous_format).exponential_().log()
            a.noise = noise.data_ptr()
    a.embed_macs, a.full_flops = float(embed_macs), float(full_model_flops)
    a.use_gumbel, a.gumbel_hard, a.eps = (1 if use_gumbel else 0), (1 if gumbel_hard else 0), float(eps)
    dev.call("uvc_admm_resource", a)
    return dev.out[0].clone()


class PresetLRScheduler(object):
    """iteration -> lr table applied to param groups that carry the key `lr_name` (reference uvc_utils.py:475-499;
    NB the dual optimiser's groups only have 'lr', so with lr_name="zlr" this never fires - same as the reference)."""

    def __init__(self, decay_schedule):
        self.decay_schedule = decay_schedule
        print('\n\n=> Using a preset learning rate schedule:')
        print(decay_schedule)
        print("\n")
        self.for_once = True

    def __call__(self, optimizer, iteration, lr_name="zlr"):
        for group in optimizer.param_groups:
            if lr_name in group:
                lr = self.decay_schedule.get(iteration, group[lr_name])
                if group[lr_name] != lr:
                    print(f"====> The learning rate paramater \"{lr_name}\" changed from {group[lr_name]} to {lr}")
                    group[lr_name] = lr

    @staticmethod
    def get_lr(optimizer):
        for group in optimizer.param_groups:
            return group['lr']
