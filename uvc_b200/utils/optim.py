"""Global-norm gradient clip + AdamW as device sweeps (`uvc_sqnorm_accum`, `uvc_clip_adamw`).

Replaces `torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)` + `torch.optim.AdamW.step()` of the UVC loops
(joint_train.py:271,428-429; post_train.py:377-379) with the same arithmetic (bias-corrected AdamW, decoupled weight decay,
clip coefficient min(1, max_norm / (norm + 1e-6)) applied to the gradients in place).  When the model's parameters live
in its flat arena (`model.flatten_parameters()`), the whole model is ONE launch for the norm and ONE for the update
(16 B/param read + 12 B/param written); anything outside the arena (gates, the token-gate Linear) is a launch per tensor.
Parameters whose .grad is None are skipped, as torch does (no weight decay either).
"""
import torch
from torch.optim import Optimizer

from .. import ops


class FusedClipAdamW(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=0.0, model=None, masks=None):
        """`model`: a DistilledVisionTransformer whose flat arenas are used when possible.
        `masks`: optional {Parameter: mask tensor}: the update is multiplied by the mask (Stage 2 keeps pruned weights at 0)."""
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.max_grad_norm = float(max_grad_norm or 0.0)
        self.model = model
        self.masks = masks or {}
        self._flat = None
        self._acc = None
        self.last_sqnorm = None       # device scalar: squared global gradient norm of the last step (before clipping)

    # ---- flat fast path ------------------------------------------------------------------------------------------
    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._flat = None            # the flat moment arenas are rebuilt from the freshly loaded per-parameter state

    def _flat_ready(self):
        """One launch for the norm and one for the update over the model's flat arenas -- also with Stage 2's masks and decay / no-decay
        parameter groups (post_train.py:299, timm create_optimizer) and with tenants that have no gradient (frozen T2T pos_embed, the weights
        of hard-skipped blocks): those differences are one option byte per element (uvc_clip_adamw_flags)."""
        m = self.model
        if m is None or m.flat_param is None:
            return None
        eng = m.engine_parameters()
        fg = m.flat_grad
        gid = {id(p): g for g in self.param_groups for p in g['params']}
        g0 = self.param_groups[0]
        if any((g['lr'], tuple(g['betas']), g['eps']) != (g0['lr'], tuple(g0['betas']), g0['eps']) for g in self.param_groups):
            return None
        wds = {float(g['weight_decay']) for g in self.param_groups if g['weight_decay']}
        if len(wds) > 1:
            return None
        lo, hi = fg.data_ptr(), fg.data_ptr() + fg.numel() * 4
        active = []
        for p in eng:
            a = id(p) in gid and p.grad is not None
            if a and not (lo <= p.grad.data_ptr() < hi):
                return None          # a gradient that does not live in the arena (set by hand): per-tensor path
            active.append(a)
        if not any(active):
            return None
        sig = (m.flat_param.data_ptr(), tuple(active), tuple(id(self.masks.get(p)) for p in eng), tuple(bool(gid[id(p)]['weight_decay']) if id(p) in gid else False for p in eng))
        if self._flat is None or self._flat["sig"] != sig:
            old = self._flat
            if old is not None and old["p"].data_ptr() == m.flat_param.data_ptr():
                fm, fv = old["m"], old["v"]
            else:
                fm, fv = torch.zeros_like(m.flat_param), torch.zeros_like(m.flat_param)
            flags = torch.zeros(m.flat_param.numel(), dtype=torch.uint8, device=m.flat_param.device)
            off = 0
            for p, a in zip(eng, active):     # expose per-parameter views so state_dict() has the usual exp_avg / exp_avg_sq entries
                n = p.numel()
                st = self.state[p]
                if "exp_avg" in st and st["exp_avg"].data_ptr() != fm[off:off + n].data_ptr():
                    fm[off:off + n].copy_(st["exp_avg"].reshape(-1)); fv[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                st["exp_avg"], st["exp_avg_sq"] = fm[off:off + n].view(p.shape), fv[off:off + n].view(p.shape)
                st.setdefault("step", 0)
                if a:
                    bits = 4 | (2 if gid[id(p)]['weight_decay'] else 0)
                    mask = self.masks.get(p)
                    if mask is None:
                        flags[off:off + n] = bits | 1
                    else:
                        flags[off:off + n] = (mask.reshape(-1) != 0).to(torch.uint8) | bits
                off += (n + 3) // 4 * 4
            self._flat = {"p": m.flat_param, "m": fm, "v": fv, "ids": {id(p) for p in eng}, "flags": flags, "sig": sig,
                          "wd": (wds.pop() if wds else 0.0), "active": [p for p, a in zip(eng, active) if a]}
        return self._flat

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        flat = self._flat_ready()
        dev = None
        todo = []     # (param, group) handled per tensor
        for group in self.param_groups:
            for p in group['params']:
                if p.grad is None:
                    continue
                dev = p.device
                if flat is not None and id(p) in flat["ids"]:
                    continue
                todo.append((p, group))
        if dev is None:
            return loss
        if self._acc is None or self._acc.device != dev:
            self._acc = torch.zeros(1, device=dev)
        acc = self._acc
        acc.zero_()
        if flat is not None:
            ops.sqnorm_accum_flags_(self.model.flat_grad, flat["flags"], acc)
        for p, _ in todo:
            if not p.grad.is_contiguous():
                p.grad = p.grad.contiguous()
            ops.sqnorm_accum_(p.grad, acc)
        self.last_sqnorm = acc
        if flat is not None:
            g = self.param_groups[0]
            step = int(self.state[flat["active"][0]]["step"]) + 1
            for p in flat["active"]:
                self.state[p]["step"] = step
            ops.clip_adamw_flags_(flat["p"], self.model.flat_grad, flat["m"], flat["v"], flat["flags"], acc, self.max_grad_norm, g['lr'], g['betas'][0],
                                  g['betas'][1], g['eps'], flat["wd"], step)
        for p, g in todo:
            st = self.state[p]
            if "exp_avg" not in st:
                st["exp_avg"], st["exp_avg_sq"], st["step"] = torch.zeros_like(p), torch.zeros_like(p), 0
            st["step"] = int(st["step"]) + 1
            if not p.is_contiguous():
                raise RuntimeError("FusedClipAdamW needs contiguous parameters")
            mask = self.masks.get(p)
            ops.clip_adamw_(p.data, p.grad, st["exp_avg"], st["exp_avg_sq"], acc, self.max_grad_norm, g['lr'], g['betas'][0], g['betas'][1],
                            g['eps'], g['weight_decay'], st["step"], mask=mask)
        return loss

    def grad_norm(self):
        """global gradient norm seen by the last step() (one device->host read)"""
        return float(self.last_sqnorm.sqrt()) if self.last_sqnorm is not None else 0.0
