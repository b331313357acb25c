"""Multi-GPU CORRECTNESS of the data-parallel path on real GPUs: two ranks over NCCL, each with half of a batch,
through uvc_b200.utils.ddp -- the averaged gradients in the flat arena must equal the single-process gradients of the whole batch, and one
optimiser step must leave both ranks with identical parameters.  Needs >= 2 GPUs (gpurun --gpus 2); skipped on a single-GPU box."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _build(depth=2):
    from functools import partial
    from oracle import fixtures as fx
    from uvc_b200.models.model_distilled import DistilledVisionTransformer
    sd, dims = fx.make_state_dict("deit_tiny_patch16_224", depth, seed=23)
    d = dict(dims); d["depth"] = depth
    m = DistilledVisionTransformer(enable_dist=0, patch_size=16, mlp_ratio=4, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), drop_rate=0,
                                   gumbel_hard=False, **d)
    m.load_state_dict(sd, strict=False)
    return m


def _step_grads(model, ddp, x, tgt, blend):
    from uvc_b200 import ops
    from uvc_b200.models.model_distilled import _VitFunction, _engine_param_list
    params = [p for _, p in _engine_param_list(model)]
    # the wrapper's hooks fire on the parameters' gradients; the forward goes through the same autograd node the public forward uses
    logits = _VitFunction.apply(model, x, blend, None, None, None, *params)
    out, dl = ops.distill_loss(logits.detach(), torch.zeros_like(logits), tgt, 0.0, 1.0)
    logits.backward(dl)
    return model.flat_grad.clone(), blend.grad.clone() if blend.grad is not None else None


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    # the half batch (788 rows) and the whole batch (1576 rows) must run the SAME LayerNorm-backward kernel for a summation-order-level comparison:
    # above 1024 rows the engine switches to the streamed kernel, whose fp32 results differ in the last bit and flip a few fp16 operand roundings
    os.environ["UVC_LN_BWD_REG"] = "1"
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from oracle import fixtures as fx
        from uvc_b200.utils.ddp import DistributedDataParallel as DDP
        from uvc_b200.utils.optim import FusedClipAdamW
        B = 8
        x, _ = fx.make_batch(B, seed=77)
        tgt = fx.soft_targets(B, seed=77)
        model = _build().cuda().train()
        if rank == 1:
            with torch.no_grad():
                model.head.weight.add_(1.0)            # rank 1 starts different: the wrap-time broadcast must erase it
        model.flatten_parameters()
        ddp = DDP(model, gradient_predivide_factor=world, delay_allreduce=True)
        lo, hi = rank * B // world, (rank + 1) * B // world
        blend = torch.tensor([[0.3, 0.7], [0.55, 0.45]], device="cuda").requires_grad_(True)
        g_ddp, _ = _step_grads(model, ddp, x[lo:hi].cuda(), tgt[lo:hi].cuda(), blend)
        torch.cuda.synchronize()
        assert ddp.bytes_reduced >= model.flat_grad.numel() * 4
        # single-process reference on the whole batch, same weights (rank 0's), no wrapper
        ref = _build().cuda().train()
        ref.flatten_parameters()
        blend_r = torch.tensor([[0.3, 0.7], [0.55, 0.45]], device="cuda").requires_grad_(True)
        g_ref, _ = _step_grads(ref, ref, x.cuda(), tgt.cuda(), blend_r)
        err = float((g_ddp - g_ref).abs().max() / g_ref.abs().max())
        # one optimiser step: parameters stay identical across ranks
        opt = FusedClipAdamW(model.parameters(), lr=1e-3, weight_decay=0.05, max_grad_norm=1.0, model=model)
        opt.step()
        mine = model.flat_param.clone()
        other = mine.clone()
        dist.broadcast(other, 0)
        same = bool(torch.equal(mine, other))
        q.put((rank, err, same, None))
    except Exception as e:      # surface the failure in the parent
        import traceback
        q.put((rank, None, None, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_rank_nccl_gradients_equal_single_rank_whole_batch():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, err, same, tb in res:
        assert tb is None, tb
        # fp16 operand storage: the two half-batch GEMMs and the whole-batch GEMM round the same fp16 products in a different order (fp32
        # accumulation); 1e-5 of the largest gradient entry is summation-order noise, far below the 1.5e-3 the engine is held to against fp32
        assert err < 2e-5, (rank, err)
        assert same, rank
