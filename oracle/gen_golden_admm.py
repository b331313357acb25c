"""ORACLE (test infrastructure only) — ADMM golden trajectory FROM THE UNMODIFIED REFERENCE.

Drives the reference's own `build_minimax_model` + `uvc_optimizer` (imported in place by oracle/ref_shim.py) on CPU
for a few steps on a DeiT-Tiny-shaped model with synthetic "trained-looking" weights, records the state after
every step, and asserts that `oracle/admm_oracle.py` (the restatement the GPU tests compare against) reproduces it.
The Gumbel noise the reference draws inside calc_flops is replayed from the same seed and stored, so the CUDA path
can be fed the same numbers.
"""
import os

import torch

from oracle import admm_oracle as ao, fixtures as fx, ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SPEC = dict(model_type="deit_tiny_patch16_224", depth=3, seed=31, steps=6, gating_interval=3, lr=5e-4,
            s0=[[0.4, 100.3], [1.2, 10.0], [0.0, 300.7]], r0_seed=5)


def init_state(L, H):
    s = torch.tensor(SPEC["s0"])
    g = torch.Generator().manual_seed(SPEC["r0_seed"])
    r = torch.rand(L, H, generator=g) * 20
    r[0, 0] = 0.0
    y = torch.full((L, 2), 1e-3) + torch.rand(L, 2, generator=g) * 0.5
    p = torch.full((L, H), 1e-3) + torch.rand(L, H, generator=g) * 0.5
    z = torch.tensor(2.0)
    return s, r, y, p, z


def gate_grad_for(step, L):
    g = torch.Generator().manual_seed(1000 + step)
    return torch.randn(L, 2, generator=g) * 0.01


def main(ns=None):
    ns = ns or ref_shim.load()
    mt, depth = SPEC["model_type"], SPEC["depth"]
    sd, dims = fx.make_state_dict(mt, depth, seed=SPEC["seed"], wstd=0.05)
    model = ref_shim.make_ref_model(ns, mt, depth=depth, gumbel_hard=False)
    model.load_state_dict(sd, strict=False)      # gates at their init [-1, 1] for the MAC probe (joint_train.py:1010-1012)
    ref_shim.register_masks(model)
    H, d = dims["num_heads"], 64
    args = ref_shim.default_args(num_heads=H, head_size=d, enable_warmup=0, zlr_schedule_list=[1, 5, 9, 13, 17], budget=0.5, z_grad_clip=0.5,
                                 gating_interval=SPEC["gating_interval"], gating_weight=5.0, sl2wd=0.01, slr=0.7, rlr=2.0, ylr=1e-2, plr=1e-2)
    layer_names, uvc_layers, uvc_dict = ns.get_uvc_layers(model)
    model.eval()
    with torch.no_grad():
        _, flops_list = model(torch.ones(1, 3, 224, 224))
    mm, dual_opt, s_opt, r_opt, g_opt = ns.uvc_optimizer.build_minimax_model(model, layer_names, uvc_layers, uvc_dict, args, flops_list)
    model.train()
    model.enable_warmup = 0
    sd["block_skip_gating"] = torch.tensor([[-1.0, 1.0], [0.3, -0.2], [0.1, 0.4]])
    with torch.no_grad():
        model.block_skip_gating.copy_(sd["block_skip_gating"])
    L = depth
    s, r, y, p, z = init_state(L, H)
    with torch.no_grad():
        mm.s.copy_(s); mm.r.copy_(r); mm.y.copy_(y); mm.p.copy_(p); mm.z.copy_(z)

    class FakeOpt:
        param_groups = [{"lr": SPEC["lr"]}]

    # restatement state
    W1 = [m.weight.detach().clone() for m in uvc_layers["W1"]]
    W3 = [m.weight.detach().clone() for m in uvc_layers["W3"]]
    gate = model.block_skip_gating.detach().clone()
    st = dict(s=s.clone(), r=r.clone(), y=y.clone(), p=p.clone(), z=z.clone(), gate=gate, gate_buf=[])
    macs = torch.Tensor(flops_list[1])
    full = float((flops_list[0] + macs.sum()) * 2)
    hp = dict(lr=SPEC["lr"], slr=args.slr, rlr=args.rlr, ylr=args.ylr, plr=args.plr, zlr=float(args.zlr_schedule_list[0]), budget=args.budget,
              z_grad_clip=args.z_grad_clip, sl2wd=args.sl2wd, gating_weight=args.gating_weight, d=d, Fh=W3[0].shape[1], macs=macs,
              embed_macs=flops_list[0], full=full, use_gumbel=True, eps=args.eps, gating_interval=args.gating_interval)
    mom = {"buf": None}

    def gate_sgd(grad):     # torch.optim.SGD(momentum .9, weight_decay 1e-4, lr glr) on the gate
        dp = grad + 1e-4 * st["gate"]
        mom["buf"] = dp.clone() if mom["buf"] is None else mom["buf"] * 0.9 + dp
        st["gate"] -= args.glr * mom["buf"]

    traj, glist = [], []
    for step in range(SPEC["steps"]):
        gg = gate_grad_for(step, L)
        model.block_skip_gating.grad = gg.clone()
        torch.manual_seed(500 + step)
        n1 = -torch.empty(L, 2).exponential_().log()
        n2 = -torch.empty(L, 2).exponential_().log()
        torch.manual_seed(500 + step)
        mm.update_gating()
        cur, s_np, r_np, g_np, glist = ns.uvc_optimizer.uvc_optimizer(FakeOpt(), mm, s_opt, r_opt, g_opt, dual_opt, args, {}, [], flops_list,
                                                                    args.z_grad_clip, step, args.gating_interval, glist)
        hp["global_step"] = step
        cur2 = ao.step(st, W1, W3, hp, n1, n2, gate_grad=gg, gate_sgd=gate_sgd)
        # the restatement must track the reference (fp32 closed forms vs autograd: tiny rounding differences allowed)
        for name, a, b in [("s", mm.s, st["s"]), ("r", mm.r, st["r"]), ("y", mm.y, st["y"]), ("p", mm.p, st["p"]), ("z", mm.z, st["z"]),
                           ("gate", model.block_skip_gating, st["gate"])]:
            torch.testing.assert_close(a.detach(), b, rtol=2e-5, atol=1e-6, msg=lambda m: f"step {step} {name}: {m}")
        assert abs(cur - cur2) < 1e-6, (step, cur, cur2)
        for l in range(L):
            torch.testing.assert_close(uvc_layers["W1"][l].weight.detach(), W1[l], rtol=1e-6, atol=0)
            torch.testing.assert_close(uvc_layers["W3"][l].weight.detach(), W3[l], rtol=1e-6, atol=0)
        traj.append(dict(cur=cur, s=mm.s.detach().clone(), r=mm.r.detach().clone(), y=mm.y.detach().clone(), p=mm.p.detach().clone(),
                         z=mm.z.detach().clone(), gate=model.block_skip_gating.detach().clone(), noise1=n1, noise2=n2, gate_grad=gg,
                         w1_sum=[fx.checksum(m.weight) for m in uvc_layers["W1"]], w3_sum=[fx.checksum(m.weight) for m in uvc_layers["W3"]]))
        print(f"  admm step {step}: resource={cur:.6f} z={float(mm.z):.5f} s={mm.s.detach().flatten().tolist()}")
    # masks after the last step (prune_w_mask) + the resource prints
    ns.uvc_utils.prune_w_mask(mm)
    m1, m3 = ao.masks(W1, W3, st["s"], st["r"], d, hp["Fh"])
    for l in range(L):
        assert torch.equal(uvc_layers["W1"][l].mask[0].bool(), m1[l]) and torch.equal(uvc_layers["W3"][l].mask[0].bool(), m3[l])
        assert torch.equal(uvc_layers["W2"][l].mask[:, 0].bool(), m3[l])
    out = dict(spec=SPEC, args={k: v for k, v in vars(args).items()}, flops_list=(int(flops_list[0]), [[int(v) for v in row] for row in flops_list[1]]),
               full=full, init=dict(s=s, r=r, y=y, p=p, z=z, gate=sd["block_skip_gating"]), traj=traj,
               masks=dict(w1=[m.mask[0].clone() for m in uvc_layers["W1"]], w3=[m.mask[0].clone() for m in uvc_layers["W3"]],
                          w2=[m.mask[:, 0].clone() for m in uvc_layers["W2"]]))
    torch.save(out, os.path.join(OUT, "admm_traj.pt"))
    print(f"  admm_traj: {len(traj)} steps, initial FLOPs {full/1e6:.2f}M")


if __name__ == "__main__":
    main()
