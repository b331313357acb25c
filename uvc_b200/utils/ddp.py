"""Data-parallel gradient averaging for the UVC loops: one process per GPU, `torch.distributed` (NCCL over NVLink/NVSwitch on
a B200 box, gloo in the CPU tests) as plumbing.

Stands in for apex `DistributedDataParallel(model, message_size=..., gradient_predivide_factor=world, delay_allreduce=True)`
(joint_train.py:293, post_train.py:291): parameters are broadcast from rank 0 at wrap time, and at the END of every backward
all gradients are averaged over the ranks.  The UVC path shards over the image batch with exactly one exchange per step, so
the collective is a single fp32 all-reduce of the model's flat gradient arena (the engine's backward writes every weight
gradient into one contiguous buffer) plus one small flattened all-reduce for the few parameters outside it (gates).
Everything else of the step (ADMM state, Gumbel noise, mixup lambda) is replicated and relies on identical seeds on all
ranks, as in the reference; `broadcast_seed()` makes that contract explicit.
"""
import torch
import torch.distributed as dist
import torch.nn as nn
from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors


def broadcast_seed(seed, device="cpu"):
    """rank 0's seed for everyone (the reference relies on every rank being launched with the same --seed)."""
    if not (dist.is_available() and dist.is_initialized()):
        return int(seed)
    t = torch.tensor([int(seed)], dtype=torch.int64, device=device)
    dist.broadcast(t, 0)
    return int(t.item())


class DistributedDataParallel(nn.Module):
    def __init__(self, module, message_size=10000000, gradient_predivide_factor=1.0, delay_allreduce=True, process_group=None, **_unused):
        super().__init__()
        self.module = module
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.predivide = float(gradient_predivide_factor)
        self._queued = False
        self.bytes_reduced = 0
        if self.world > 1:
            flat = getattr(module, "flat_param", None)
            with torch.no_grad():
                if flat is not None:
                    dist.broadcast(flat, 0, group=self.group)
                    eng = {id(p) for p in module.engine_parameters()}
                else:
                    eng = set()
                rest = [p.data for p in module.parameters() if id(p) not in eng] + [b.data for b in module.buffers()]
                for t in rest:
                    dist.broadcast(t, 0, group=self.group)
            # one live wrapper per module: wrapping the same model again (Stage 2 inside joint_train wraps the Stage-1 model a second time)
            # detaches the earlier wrapper's hooks first, otherwise every backward would all-reduce the arena once per wrapper
            prev = getattr(module, "_uvc_ddp_wrapper", None)
            if prev is not None and prev is not self:
                prev.remove()
            self._handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in module.parameters() if p.requires_grad]
            object.__setattr__(module, "_uvc_ddp_wrapper", self)

    def remove(self):
        """Detach this wrapper's gradient hooks (the module stays usable, un-synchronised)."""
        for h in getattr(self, "_handles", []):
            h.remove()
        self._handles = []
        if getattr(self.module, "_uvc_ddp_wrapper", None) is self:
            object.__setattr__(self.module, "_uvc_ddp_wrapper", None)

    def forward(self, *a, **k):
        return self.module(*a, **k)

    # the first gradient of a backward pass queues one callback that runs when the whole backward has finished
    def _on_grad(self, _p):
        if not self._queued:
            self._queued = True
            torch.autograd.Variable._execution_engine.queue_callback(self._finalize)

    def _allreduce_mean(self, t):
        pre, post = 1.0 / self.predivide, self.predivide / self.world
        if t.is_cuda and self.predivide == float(self.world) and dist.get_backend(self.group) == "nccl":
            # apex's setting in the UVC loops (gradient_predivide_factor = world size): every rank's gradient is multiplied by 1 / world BEFORE
            # the sum.  That is exactly NCCL's AVG (a pre-multiplied sum inside the collective), so the separate elementwise pass over the
            # 88 MB arena disappears.
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
        else:
            if pre != 1.0:
                t.mul_(pre)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            if post != 1.0:
                t.mul_(post)
        self.bytes_reduced += t.numel() * t.element_size()

    def _finalize(self):
        self._queued = False
        m = self.module
        self.bytes_reduced = 0
        in_arena = set()
        fg = getattr(m, "flat_grad", None) if getattr(m, "engine_parameters", None) else None
        if fg is not None:
            lo, hi = fg.data_ptr(), fg.data_ptr() + fg.numel() * 4
            eng = [p for p in m.engine_parameters() if p.requires_grad]      # frozen arena tenants (T2T's sinusoid pos_embed) have no .grad
            if eng and all(p.grad is not None and lo <= p.grad.data_ptr() < hi for p in eng):
                self._allreduce_mean(fg)                      # ONE collective for the whole model
                in_arena = {id(p) for p in eng}
        rest = [p.grad for p in m.parameters() if p.grad is not None and id(p) not in in_arena]
        if rest:
            flat = _flatten_dense_tensors(rest)
            self._allreduce_mean(flat)
            for g, r in zip(rest, _unflatten_dense_tensors(flat, rest)):
                g.copy_(r)
