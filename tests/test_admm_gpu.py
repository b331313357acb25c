"""GPU parity of the device ADMM step (uvc_admm_* through uvc_b200.uvc_optimizer / uvc_utils) against the trajectory
the UNMODIFIED reference produced (tests/golden/admm_traj.pt, written by oracle/gen_golden_admm.py) and against the
oracle restatement on other shapes.  Selections (mask indices) must be exact; fp32 state within 1e-4 relative."""
import os
import types

import pytest
import torch

from oracle import admm_oracle as ao, fixtures as fx

pytestmark = pytest.mark.gpu


def get_uvc_layers(model):
    """the module scan of joint_train.py:530-564 (same as uvc_b200.joint_train.get_uvc_layers)"""
    from uvc_b200.joint_train import get_uvc_layers as g
    return g(model)


def build(model_type, depth, sd):
    from test_model_gpu import build as b
    m = b(model_type, depth, sd, gumbel_hard=False)
    for _, mod in m.named_modules():
        if hasattr(mod, "weight"):
            mod.register_buffer("mask", torch.ones_like(mod.weight))
    return m


class FakeOpt:
    def __init__(self, lr):
        self.param_groups = [{"lr": lr}]


def close(a, b, rtol=1e-4, atol=1e-6):
    torch.testing.assert_close(a.detach().cpu().float(), b.detach().cpu().float(), rtol=rtol, atol=atol)


def test_admm_trajectory_matches_reference_golden(golden_dir):
    from uvc_b200.uvc_optimizer import build_minimax_model, uvc_optimizer
    from uvc_b200.uvc_utils import prune_w_mask
    G = torch.load(os.path.join(golden_dir, "admm_traj.pt"), weights_only=False)
    sp = G["spec"]
    sd, dims = fx.make_state_dict(sp["model_type"], sp["depth"], seed=sp["seed"], wstd=0.05)
    model = build(sp["model_type"], sp["depth"], sd)
    args = types.SimpleNamespace(**G["args"])
    layer_names, uvc_layers, uvc_dict = get_uvc_layers(model)
    model.eval()
    with torch.no_grad():
        _, flops_list = model(torch.ones(1, 3, 224, 224, device="cuda"))
    assert int(flops_list[0]) == G["flops_list"][0] and [[int(v) for v in r] for r in flops_list[1]] == G["flops_list"][1]
    mm, dual_opt, s_opt, r_opt, g_opt = build_minimax_model(model, layer_names, uvc_layers, uvc_dict, args, flops_list)
    assert abs(mm.full_flops - G["full"]) <= 1e-6 * G["full"]
    model.train(); model.enable_warmup = 0
    with torch.no_grad():
        model.block_skip_gating.copy_(G["init"]["gate"])
        for k in ("s", "r", "y", "p", "z"):
            getattr(mm, k).copy_(G["init"][k])
    glist = []
    for step, t in enumerate(G["traj"]):
        model.block_skip_gating.grad = t["gate_grad"].cuda()
        noises = iter([t["noise1"], t["noise2"]])
        mm.noise_source = lambda: next(noises)
        mm.update_gating()
        cur, s_np, r_np, g_np, glist = uvc_optimizer(FakeOpt(sp["lr"]), mm, s_opt, r_opt, g_opt, dual_opt, args, {}, [], flops_list,
                                                     args.z_grad_clip, step, args.gating_interval, glist)
        assert abs(cur - t["cur"]) < 2e-6, (step, cur, t["cur"])
        for k in ("s", "r", "y", "p", "z"):
            close(getattr(mm, k), t[k])
        close(model.block_skip_gating, t["gate"])
        close(torch.from_numpy(s_np), t["s"]); close(torch.from_numpy(r_np), t["r"]); close(torch.from_numpy(g_np), t["gate"])
        for l in range(sp["depth"]):
            for grp, key in (("W1", "w1_sum"), ("W3", "w3_sum")):
                got = fx.checksum(uvc_layers[grp][l].weight)
                for a, b in zip(got, t[key][l]):
                    assert abs(a - b) <= 1e-5 * max(1.0, abs(b)), (step, grp, l, a, b)
    prune_w_mask(mm)
    for l in range(sp["depth"]):      # selections are index-exact
        assert torch.equal(uvc_layers["W1"][l].mask.cpu()[0], G["masks"]["w1"][l]) and (uvc_layers["W1"][l].mask.cpu() == G["masks"]["w1"][l]).all()
        assert (uvc_layers["W3"][l].mask.cpu() == G["masks"]["w3"][l]).all()
        assert (uvc_layers["W2"][l].mask.cpu() == G["masks"]["w2"][l].unsqueeze(1)).all()


def test_admm_warmup_returns_resource_only():
    from uvc_b200.uvc_optimizer import build_minimax_model, uvc_optimizer
    sd, dims = fx.make_state_dict("deit_tiny_patch16_224", 2, seed=8)
    model = build("deit_tiny_patch16_224", 2, sd)
    args = types.SimpleNamespace(head_size=64, num_heads=3, flops_with_mhsa=1, use_gumbel=1, enable_block_gating=1, enable_part_gating=0,
                                 enable_patch_gating=0, enable_jumping=0, eps=0.1, eps_decay=0.92, enable_warmup=1, soptim="sgd", roptim="sgd",
                                 slr=0.02, rlr=0.02, glr=0.1, ylr=1e-4, plr=1e-4, zlr_schedule_list=[1, 5], budget=0.5, sl2wd=0.0,
                                 gating_weight=5e-4, z_grad_clip=0.5, gating_interval=50)
    ln, ul, ud = get_uvc_layers(model)
    model.eval()
    with torch.no_grad():
        _, flops_list = model(torch.ones(1, 3, 224, 224, device="cuda"))
    mm, dual_opt, s_opt, r_opt, g_opt = build_minimax_model(model, ln, ul, ud, args, flops_list)
    model.train(); model.enable_warmup = 1
    noise = torch.zeros(2, 2)
    mm.noise_source = lambda: noise
    w_before = ul["W1"][0].weight.clone()
    cur, s_np, r_np, g_np, _ = uvc_optimizer(FakeOpt(1e-3), mm, s_opt, r_opt, g_opt, dual_opt, args, {}, [], flops_list, 0.5, 0, 50, [])
    # s = r = 0, gates [-1, 1], zero noise: g = softmax([-2, 2])[1]; resource = (embed + g * blocks) / (embed + blocks)
    g = torch.softmax(torch.tensor([-2.0, 2.0]), 0)[1].item()
    embed, blocks = float(flops_list[0]), float(sum(sum(r) for r in flops_list[1]))
    assert abs(cur - (embed + g * blocks) / (embed + blocks)) < 1e-5
    assert (s_np == 0).all() and (r_np == 0).all() and float(mm.z) == pytest.approx(1e-3)
    assert torch.equal(ul["W1"][0].weight, w_before)             # k = 0: the prox is a no-op
    # real / expected FLOPs prints (joint_train.py:509)
    hard = float(mm.run_resource_fn(gumbel_hard=True))
    assert abs(hard - 1.0) < 1e-6
    # lazy=True (the hot loop's variant): same values, fetched through one deferred pinned-memory copy instead of a synchronous one
    import numpy as np
    cur2, s2, r2, g2, _ = uvc_optimizer(FakeOpt(1e-3), mm, s_opt, r_opt, g_opt, dual_opt, args, {}, [], flops_list, 0.5, 1, 50, [], lazy=True)
    assert abs(float(cur2) - cur) < 1e-6 and np.array_equal(np.asarray(s2), mm.s.detach().cpu().numpy()) and np.asarray(r2).shape == r_np.shape
    assert g2.tolist() == mm.block_skip_gating.detach().cpu().numpy().tolist() and s2.size == s_np.size


@pytest.mark.parametrize("model_type,depth", [("deit_small_patch16_224", 12), ("deit_base_patch16_224", 2)])
def test_admm_step_matches_oracle_on_full_size_layers(model_type, depth):
    """DeiT-Small (BASELINE config) and DeiT-Base sized layers against the CPU restatement for 2 steps."""
    from uvc_b200.uvc_optimizer import build_minimax_model, uvc_optimizer
    sd, dims = fx.make_state_dict(model_type, depth, seed=77, wstd=0.05)
    model = build(model_type, depth, sd)
    H, C = dims["num_heads"], dims["embed_dim"]
    args = types.SimpleNamespace(head_size=64, num_heads=H, flops_with_mhsa=1, use_gumbel=1, enable_block_gating=1, enable_part_gating=0,
                                 enable_patch_gating=0, enable_jumping=0, eps=0.1, eps_decay=0.92, enable_warmup=0, soptim="sgd", roptim="sgd",
                                 slr=5.0, rlr=5.0, glr=0.1, ylr=1e-2, plr=1e-2, zlr_schedule_list=[1, 5], budget=0.5, sl2wd=0.0,
                                 gating_weight=5e-4, z_grad_clip=0.5, gating_interval=50)
    ln, ul, ud = get_uvc_layers(model)
    model.eval()
    with torch.no_grad():
        _, flops_list = model(torch.ones(1, 3, 224, 224, device="cuda"))
    mm, dual_opt, s_opt, r_opt, g_opt = build_minimax_model(model, ln, ul, ud, args, flops_list)
    model.train(); model.enable_warmup = 0
    L = depth
    gen = torch.Generator().manual_seed(3)
    s0 = torch.stack([torch.rand(L, generator=gen) * (H - 1), torch.rand(L, generator=gen) * 4 * C * 0.6], 1)
    r0 = torch.rand(L, H, generator=gen) * 40
    y0, p0, z0 = torch.rand(L, 2, generator=gen), torch.rand(L, H, generator=gen), torch.tensor(1.5)
    with torch.no_grad():
        for k, v in (("s", s0), ("r", r0), ("y", y0), ("p", p0), ("z", z0)):
            getattr(mm, k).copy_(v)
    W1 = [m.weight.detach().cpu().clone() for m in ul["W1"]]
    W3 = [m.weight.detach().cpu().clone() for m in ul["W3"]]
    st = dict(s=s0.clone(), r=r0.clone(), y=y0.clone(), p=p0.clone(), z=z0.clone(), gate=model.block_skip_gating.detach().cpu().clone(), gate_buf=[])
    macs = torch.Tensor(flops_list[1])
    hp = dict(lr=1e-3, slr=args.slr, rlr=args.rlr, ylr=args.ylr, plr=args.plr, zlr=1.0, budget=0.5, z_grad_clip=0.5, sl2wd=0.0,
              gating_weight=args.gating_weight, d=64, Fh=4 * C, macs=macs, embed_macs=flops_list[0], full=mm.full_flops, use_gumbel=True,
              eps=0.1, gating_interval=50)
    glist = []
    for step in range(2):
        gg = torch.randn(L, 2, generator=gen) * 0.01
        n1, n2 = [-torch.empty(L, 2).exponential_(generator=gen).log() for _ in range(2)]
        model.block_skip_gating.grad = gg.cuda()
        noises = iter([n1, n2]); mm.noise_source = lambda: next(noises)
        cur, s_np, r_np, g_np, glist = uvc_optimizer(FakeOpt(1e-3), mm, s_opt, r_opt, g_opt, dual_opt, args, {}, [], flops_list, 0.5, step, 50, glist)
        hp["global_step"] = step
        cur2 = ao.step(st, W1, W3, hp, n1, n2, gate_grad=gg, gate_sgd=lambda g: None)
        assert abs(cur - cur2) < 5e-6
        for k in ("s", "r", "y", "p", "z"):
            close(getattr(mm, k), st[k])
        for l in range(L):
            close(ul["W1"][l].weight, W1[l], rtol=1e-6, atol=0)
            close(ul["W3"][l].weight, W3[l], rtol=1e-6, atol=0)
    close(glist.acc, torch.cat(st["gate_buf"]).sum(0), rtol=1e-4, atol=1e-7)


def test_admm_selections_under_exact_ties_match_reference_golden(golden_dir):
    """Exactly-zero columns / a dead head / duplicated columns (the state Stage 2 lives in): tests/golden/admm_ties.pt was written by the
    UNMODIFIED reference (oracle/gen_golden_admm_ties.py).  Case A: every tie group lies entirely inside or outside each selection -> masks must
    be equal index for index and the trajectory must follow.  Case B: the boundary cuts a group of 100 equal (zero) scores with k = 60 -> torch.topk
    leaves the choice inside the group unspecified on the CPU; defined (and asserted) are the count taken from the group and everything outside it;
    this implementation takes the lower indices, as torch.topk does on CUDA."""
    from oracle.gen_golden_admm_ties import tie_state
    from uvc_b200.uvc_optimizer import build_minimax_model, uvc_optimizer
    from uvc_b200.uvc_utils import prune_w_mask
    G = torch.load(os.path.join(golden_dir, "admm_ties.pt"), weights_only=False)
    sp = G["spec"]
    sd, dims = fx.make_state_dict(sp["model_type"], sp["depth"], seed=sp["seed"], wstd=0.05)
    sd = tie_state(sd, sp["depth"])
    model = build(sp["model_type"], sp["depth"], sd)
    args = types.SimpleNamespace(**G["args"])
    layer_names, uvc_layers, uvc_dict = get_uvc_layers(model)
    model.eval()
    with torch.no_grad():
        _, flops_list = model(torch.ones(1, 3, 224, 224, device="cuda"))
    mm, dual_opt, s_opt, r_opt, g_opt = build_minimax_model(model, layer_names, uvc_layers, uvc_dict, args, flops_list)
    model.train(); model.enable_warmup = 0
    with torch.no_grad():
        for k in ("s", "r", "y", "p", "z"):
            getattr(mm, k).copy_(G["init"][k])
    prune_w_mask(mm)
    L = sp["depth"]
    for l in range(L):      # case A: index for index
        assert torch.equal(uvc_layers["W1"][l].mask.cpu()[0], G["masksA"]["w1"][l]) and (uvc_layers["W1"][l].mask.cpu() == G["masksA"]["w1"][l]).all(), l
        assert (uvc_layers["W3"][l].mask.cpu() == G["masksA"]["w3"][l]).all(), l
        assert (uvc_layers["W2"][l].mask.cpu() == G["masksA"]["w2"][l].unsqueeze(1)).all(), l
    glist = []
    for step, t in enumerate(G["traj"]):
        model.block_skip_gating.grad = torch.zeros(L, 2, device="cuda")
        noises = iter([t["noise1"], t["noise2"]])
        mm.noise_source = lambda: next(noises)
        mm.update_gating()
        cur, s_np, r_np, g_np, glist = uvc_optimizer(FakeOpt(sp["lr"]), mm, s_opt, r_opt, g_opt, dual_opt, args, {}, [], flops_list,
                                                     args.z_grad_clip, step, args.gating_interval, glist)
        assert abs(cur - t["cur"]) < 2e-6, (step, cur, t["cur"])
        for k in ("s", "r", "y", "p", "z"):
            close(getattr(mm, k), t[k])
        for l in range(L):
            for grp, key in (("W1", "w1_sum"), ("W3", "w3_sum")):
                for a, b in zip(fx.checksum(uvc_layers[grp][l].weight), t[key][l]):
                    assert abs(a - b) <= 1e-5 * max(1.0, abs(b)), (step, grp, l, a, b)
    # case B: the cut tie group
    cb = G["caseB"]
    with torch.no_grad():
        mm.s.copy_(cb["s"]); mm.r.copy_(G["init"]["r"])
    prune_w_mask(mm)
    for l in range(L):
        off = uvc_layers["W3"][l].mask.cpu()[0] == 0
        ref_off = cb["masks_w3"][l] == 0
        assert int(off.sum()) == cb["k"] == int(ref_off.sum())
        assert bool(off[cb["tie_group"]].sum() == cb["k"])                       # all taken from the tie group, like the reference
        assert torch.equal(off.nonzero().flatten(), cb["tie_group"][:cb["k"]])   # and inside it: the lower indices
