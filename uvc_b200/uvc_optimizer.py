"""One ADMM primal-dual step of UVC, on the device.

Mirror of the reference's `UVC/uvc_optimizer.py` (`uvc_optimizer`, `uvc_optimizer_gating`, `build_minimax_model`
keep their signatures and return values).  The reference's step is a Python loop nest with tens of thousands
of host<->device syncs; this one is six kernel launches and ONE device->host copy (the values the API returns):

    uvc_admm_scores -> uvc_admm_prox -> uvc_admm_scores -> uvc_admm_primal -> [gate SGD every `gating_interval`] -> uvc_admm_dual

Random numbers: the two [L,2] Gumbel draws of the reference's two `calc_flops` evaluations are made with the
same torch calls in the same order, so a run seeded like the reference consumes the generator identically.
"""
import torch

from .uvc_utils import UVC_CP_MiniMax, calc_flops, prox_w, proj_dual, weight_list_to_scores  # noqa: F401


class GateGradBuffer(list):
    """Stand-in for the reference's `gating_grad_list` (a Python list of per-step gate gradients, each weighted by
    `global_step % gating_interval`, averaged every `gating_interval` steps): the running sum lives on the device."""

    def __init__(self):
        super().__init__()
        self.acc = None
        self.count = 0


def _sgd_lr(opt, what):
    if opt is None:
        return 0.0
    if not isinstance(opt, torch.optim.SGD):
        raise NotImplementedError(f"the device ADMM step implements the shipped --{what} sgd (got {type(opt).__name__})")
    return float(opt.param_groups[0]['lr'])


class _Snapshot:
    """One asynchronous device->host copy into pinned memory (torch's caching host allocator recycles the buffer) + the event that fences it."""

    def __init__(self, flat):
        self.host = torch.empty(flat.shape, dtype=flat.dtype, pin_memory=True)
        self.host.copy_(flat, non_blocking=True)
        self.event = torch.cuda.Event()
        self.event.record()
        self._np = None

    def numpy(self):
        if self._np is None:
            self.event.synchronize()
            self._np = self.host.numpy()
        return self._np


class Deferred:
    """A host value that is fetched from its snapshot when first needed."""

    def __init__(self, snap, pick):
        self._snap, self._pick, self._val = snap, pick, None

    def get(self):
        if self._val is None:
            self._val = self._pick(self._snap.numpy())
        return self._val

    def __float__(self):
        return float(self.get())

    def __array__(self, dtype=None, copy=None):
        import numpy as np
        a = np.asarray(self.get())
        return a.astype(dtype) if dtype is not None else a

    def tolist(self):
        v = self.get()
        return v.tolist() if hasattr(v, "tolist") else v

    @property
    def size(self):
        return self.get().size

    @property
    def shape(self):
        return self.get().shape


def uvc_optimizer(optimizer, minimax_model, s_optimizer, r_optimizer, gating_optimizer, dual_optimizer, args, infos, save_budgets,
                  flops_list, z_grad_clip, global_step, gating_interval, gating_grad_list, lazy=False):
    """-> (cur_resource: float, s: np[L,2], r: np[L,H], gating: np[L,2] | None, gating_grad_list)   (reference :37-144)

    lazy=True (not in the reference; used by the hot loop): the four host values come back as `Deferred` handles over ONE asynchronous
    copy into pinned memory, resolved on first use (float(), np.asarray(), .tolist()).  The reference's signature forces a device->host
    synchronisation in every step although the values are only printed every `log_interval` steps; without it the host runs a step ahead
    and the GPU never waits for launches."""
    mm = minimax_model
    d = mm._dev
    warmup = bool(mm.model.enable_warmup)

    prox_w(mm, optimizer)                                   # :42   (scores of the pre-prox weights inside)
    d.scores()                                              # :46-48 read the post-prox weights
    noise1 = mm.gumbel_noise()                              # first Gumbel draw (srloss2)
    a = mm.admm_args(noise=noise1, gumbel_hard=False, warmup=warmup)
    a.z_grad_clip = float(z_grad_clip)
    a.slr, a.rlr = _sgd_lr(s_optimizer, "soptim"), _sgd_lr(r_optimizer, "roptim")
    gate = mm.block_skip_gating
    buf = gating_grad_list
    keep = []
    if not warmup and gating_optimizer is not None and gate is not None:
        if not isinstance(buf, GateGradBuffer):
            buf = GateGradBuffer()
        if buf.acc is None:
            buf.acc = torch.zeros_like(gate.data)
        if gate.grad is not None:
            gg = gate.grad.detach().contiguous()
            keep.append(gg)
            a.gate_grad = gg.data_ptr()
        a.gate_grad_acc = buf.acc.data_ptr()
        a.gate_mult = float(global_step % gating_interval)
    d.call("uvc_admm_primal", a)                            # :46-123
    cur = d.out.clone()

    if not warmup:
        if gating_optimizer is not None and gate is not None:
            buf.count += 1
            gating_optimizer.zero_grad()
            if (global_step + 1) % gating_interval == 0:    # :94-98
                gate.grad = buf.acc / float(buf.count)
                gating_optimizer.step()
                buf = GateGradBuffer()
                gating_optimizer.zero_grad()
        # dual ascent (:126-135) with the updated s, r (and gate), second Gumbel draw (zloss)
        noise2 = mm.gumbel_noise()
        b = mm.admm_args(noise=noise2, gumbel_hard=False)
        groups = dual_optimizer.param_groups
        b.zlr, b.ylr, b.plr = float(groups[0]['lr']), float(groups[1]['lr']), float(groups[2]['lr'])
        d.call("uvc_admm_dual", b)

    # one device->host copy for everything the API returns (:138-144)
    L, H = mm.s.shape[0], mm.r.shape[1]
    parts = [cur, mm.s.detach().reshape(-1), mm.r.detach().reshape(-1)]
    if gate is not None:
        parts.append(gate.detach().reshape(-1))
    if lazy:
        snap = _Snapshot(torch.cat(parts))
        return (Deferred(snap, lambda h: float(h[0])), Deferred(snap, lambda h: h[1:1 + 2 * L].reshape(L, 2).copy()),
                Deferred(snap, lambda h: h[1 + 2 * L:1 + 2 * L + L * H].reshape(L, H).copy()),
                Deferred(snap, lambda h: h[1 + 2 * L + L * H:].reshape(L, 2).copy()) if gate is not None else None, buf)
    host = torch.cat(parts).cpu().numpy()
    cur_resource = float(host[0])
    s_np = host[1:1 + 2 * L].reshape(L, 2).copy()
    r_np = host[1 + 2 * L:1 + 2 * L + L * H].reshape(L, H).copy()
    g_np = host[1 + 2 * L + L * H:].reshape(L, 2).copy() if gate is not None else None
    return cur_resource, s_np, r_np, g_np, buf


def uvc_optimizer_gating(optimizer, minimax_model, s_optimizer, r_optimizer, gating_optimizer, dual_optimizer, args, infos, save_budgets,
                         flops_list, *unused):
    """Gating-only variant (reference :148-161): resource evaluation + dual ascent on z only.  (The reference's caller
    passes 14 arguments to this 10-argument function and unpacks 5 values, i.e. `--enable_pruning 0` crashes there;
    extra positional arguments are accepted here and the 5-tuple shape of `uvc_optimizer` is returned.)"""
    mm = minimax_model
    cur = mm.run_resource_fn()
    res2 = mm.run_resource_fn()
    with torch.no_grad():
        mm.z.add_(float(dual_optimizer.param_groups[0]['lr']) * (res2 - float(args.budget)))
    proj_dual(mm)
    gate = mm.block_skip_gating
    return (float(cur), mm.s.detach().cpu().numpy(), mm.r.detach().cpu().numpy(),
            gate.detach().cpu().numpy() if gate is not None else None, unused[-1] if unused else [])


def build_minimax_model(model, layer_names, uvc_layers, uvc_layers_dict, args, flops_list, vanilla=False):
    """-> (minimax_model, dual_optimizer, s_optimizer, r_optimizer, gating_optimizer)   (reference :164-268)"""
    if not getattr(args, "flops_with_mhsa", 1):
        raise NotImplementedError("--flops_with_mhsa 0 selects the legacy flops2 cost model, which no shipped script uses")
    minimax_model = UVC_CP_MiniMax(model, resource_fn=None, uvc_layers=uvc_layers, uvc_layers_dict=uvc_layers_dict,
                                   head_size=args.head_size, num_heads=args.num_heads, flops_list=flops_list, args=args)
    resource_ub = minimax_model.full_flops

    def resource_fn(s_, r_, gating, eps, gumbel_hard=False):
        return calc_flops(s_, r_, uvc_layers_dict, uvc_layers, args.head_size, s_ub=minimax_model.s_ub, r_ub=minimax_model.r_ub,
                          flops_list=flops_list, gating=gating, full_model_flops=resource_ub, eps=eps, use_gumbel=args.use_gumbel,
                          gumbel_hard=gumbel_hard, args=args)

    minimax_model.resource_fn = resource_fn
    m = minimax_model.model
    m.enable_block_gating = args.enable_block_gating
    m.enable_part_gating = args.enable_part_gating
    m.enable_patch_gating = args.enable_patch_gating
    m.enable_jumping = args.enable_jumping
    m.use_gumbel = args.use_gumbel
    m.eps = args.eps
    m.enable_warmpup = args.enable_warmup       # (sic) the reference sets this misspelt attribute (:224)
    print(f"** Initial FLOP size: {resource_ub/1e6:.2f}M")
    if vanilla:
        return minimax_model

    def make(kind, param, lr):
        if kind == 'sgd':
            return torch.optim.SGD([param], lr, momentum=0.0, weight_decay=0.0)
        raise NotImplementedError(f"the device ADMM step implements the shipped optimiser 'sgd' for s / r (got '{kind}')")

    s_optimizer = make(args.soptim, minimax_model.s, args.slr)
    r_optimizer = make(args.roptim, minimax_model.r, args.rlr)
    gating_optimizer = torch.optim.SGD([minimax_model.block_skip_gating], args.glr, momentum=0.9, weight_decay=1e-4) \
        if args.enable_block_gating else None
    dual_optimizer = torch.optim.SGD([{'params': minimax_model.z, 'lr': args.zlr_schedule_list[0]},
                                      {'params': minimax_model.y, 'lr': args.ylr},
                                      {'params': minimax_model.p, 'lr': args.plr}], 1.0, momentum=0.0, weight_decay=0.0)
    return minimax_model, dual_optimizer, s_optimizer, r_optimizer, gating_optimizer
