"""world_size-2 CPU (gloo) tests of the data-parallel host logic (uvc_b200/utils/ddp.py): rank-0 parameter broadcast at wrap time,
ONE flat all-reduce (mean) of the gradient arena at the end of backward + one for the parameters outside it, apex's
`gradient_predivide_factor` arithmetic, and the explicit seed agreement.  The N>1 GPU path runs the same code over NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


class ArenaModel(nn.Module):
    """mimics the engine model: parameters are views into one flat buffer, gradients accumulate into one flat arena"""

    def __init__(self, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.flat_param = torch.randn(40, generator=g)
        self.flat_grad = torch.zeros(40)
        self.w = nn.Parameter(self.flat_param[:32].view(4, 8)); self.b = nn.Parameter(self.flat_param[32:36])
        self.pos = nn.Parameter(self.flat_param[36:40], requires_grad=False)   # frozen arena tenant (like T2T-ViT's sinusoid pos_embed): no .grad
        self.gate = nn.Parameter(torch.randn(3, 2, generator=g))          # lives outside the arena (like block_skip_gating)
        self.w.grad = self.flat_grad[:32].view(4, 8); self.b.grad = self.flat_grad[32:36]
        self.register_buffer("mask", torch.full((4, 8), float(seed)))

    def engine_parameters(self):
        return [self.w, self.b, self.pos]

    def forward(self, x):
        return (x @ self.w.t() + self.b).sum() * self.gate.sum()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from uvc_b200.utils.ddp import DistributedDataParallel as DDP, broadcast_seed
        from uvc_b200.utils.dist_util import get_rank, get_world_size
        assert get_world_size() == world and get_rank() == rank
        assert broadcast_seed(730 + rank) == 730
        m = ArenaModel(seed=10 + rank)
        ddp = DDP(m, message_size=250000000, gradient_predivide_factor=world, delay_allreduce=True)
        # wrap time: everything (arena, outside parameters, buffers) equals rank 0's
        ref = ArenaModel(seed=10)
        assert torch.equal(m.flat_param, ref.flat_param) and torch.equal(m.gate.data, ref.gate.data) and torch.equal(m.mask, ref.mask)
        x = torch.full((5, 8), float(rank + 1))
        ddp(x).backward()
        # expected: mean over ranks of the single-rank gradients
        exp_w, exp_b, exp_g = torch.zeros(4, 8), torch.zeros(4), torch.zeros(3, 2)
        for r in range(world):
            mm = ArenaModel(seed=10)
            mm(torch.full((5, 8), float(r + 1))).backward()
            exp_w += mm.w.grad / world; exp_b += mm.b.grad / world; exp_g += mm.gate.grad / world
        torch.testing.assert_close(m.w.grad, exp_w); torch.testing.assert_close(m.b.grad, exp_b); torch.testing.assert_close(m.gate.grad, exp_g)
        assert m.w.grad.data_ptr() == m.flat_grad.data_ptr(), "gradients must stay in the flat arena (one collective per step)"
        assert ddp.bytes_reduced == m.flat_grad.numel() * 4 + m.gate.numel() * 4
        # a second step accumulates into zeroed grads the same way
        m.flat_grad.zero_(); m.gate.grad = None
        ddp(x).backward()
        torch.testing.assert_close(m.w.grad, exp_w)
        # wrapping the same module again (Stage 2 inside joint_train) replaces the first wrapper's hooks: still ONE exchange per backward
        ddp2 = DDP(m, gradient_predivide_factor=world)
        calls = []
        orig = ddp._finalize
        ddp._finalize = lambda: (calls.append(1), orig())
        m.flat_grad.zero_(); m.gate.grad = None
        ddp2(x).backward()
        torch.testing.assert_close(m.w.grad, exp_w)
        assert not calls and ddp2.bytes_reduced == m.flat_grad.numel() * 4 + m.gate.numel() * 4
        q.put((rank, "ok"))
    except Exception as e:       # surface the failure in the parent
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_ddp_two_ranks_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res
