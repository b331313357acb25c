"""ORACLE (test infrastructure only) — ADMM selections under EXACT TIES, from the UNMODIFIED reference.

Stage 2 and late Stage 1 hold many exactly-zero columns (pruned heads / head dims / neurons), so the "k smallest" selections of
uvc_utils.py:54-73,315-401 meet exact ties.  The reference selects with torch.topk(largest=False, sorted=False) on CPU tensors, whose order
AMONG equal values is whatever libstdc++'s nth_element / partial_sort leaves (probed here: it is not "lower index first"); what IS defined is
every selection whose boundary does not cut through a group of equal scores.  This generator drives the reference's own prune_w_mask and
uvc_optimizer on such a state:
  case A  zero columns / a zero head / duplicated columns, all tie groups entirely inside or entirely outside each selection (the state Stage 2
          is in: k equals the number of pruned groups, or exceeds it) -> masks and the ADMM trajectory are tie-order independent: stored, and the
          CUDA path must reproduce the masks index for index;
  case B  a selection boundary that cuts a group of 100 zero columns with k = 60 -> only the COUNT taken from the tie group and everything
          outside it is defined; both are stored (the reference's own choice inside the group is stored too, for the record).
Run here (the only place /root/reference exists):   python -m oracle.gen_golden_admm_ties
"""
import os

import torch

from oracle import admm_oracle as ao, fixtures as fx, ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SPEC = dict(model_type="deit_tiny_patch16_224", depth=2, seed=41, steps=2, lr=5e-4, gating_interval=3)


def tie_state(sd, depth):
    """zeroed / duplicated columns written into the state dict (the same edit is replayed by the GPU test)"""
    d = 64
    for l in range(depth):
        w1 = sd[f"blocks.{l}.attn.proj.weight"]
        w3 = sd[f"blocks.{l}.mlp.fc2.weight"]
        w2 = sd[f"blocks.{l}.mlp.fc1.weight"]
        w1[:, [3, 7, 20]] = 0.0                    # three dead dims in head 0
        w1[:, 2 * d:3 * d] = 0.0                   # head 2 entirely dead
        w1[:, d + 9] = w1[:, d + 5]                # duplicated columns in head 1 (equal norms, both far from the boundary)
        zero3 = torch.arange(5, 705, 7)[:100]      # 100 dead neurons
        w3[:, zero3] = 0.0; w2[zero3, :] = 0.0
        w3[:, 1] = w3[:, 0]                        # duplicated neuron columns
    return sd


def main(ns=None):
    ns = ns or ref_shim.load()
    mt, depth = SPEC["model_type"], SPEC["depth"]
    sd, dims = fx.make_state_dict(mt, depth, seed=SPEC["seed"], wstd=0.05)
    sd = tie_state(sd, depth)
    model = ref_shim.make_ref_model(ns, mt, depth=depth, gumbel_hard=False)
    model.load_state_dict(sd, strict=False)
    ref_shim.register_masks(model)
    H, d = dims["num_heads"], 64
    args = ref_shim.default_args(num_heads=H, head_size=d, enable_warmup=0, zlr_schedule_list=[1, 5], budget=0.5, z_grad_clip=0.5,
                                 gating_interval=SPEC["gating_interval"], gating_weight=5.0, sl2wd=0.0, slr=0.05, rlr=0.05, ylr=1e-2, plr=1e-2)
    layer_names, uvc_layers, uvc_dict = ns.get_uvc_layers(model)
    model.eval()
    with torch.no_grad():
        _, flops_list = model(torch.ones(1, 3, 224, 224))
    mm, dual_opt, s_opt, r_opt, g_opt = ns.uvc_optimizer.build_minimax_model(model, layer_names, uvc_layers, uvc_dict, args, flops_list)
    model.train(); model.enable_warmup = 0
    L = depth
    # case A: every selection takes its whole tie group.  s0 = 1 head (the dead one), r: 5 dims in head 0 (3 dead + 2 live), 2 in head 1, any in the
    # dead head (its 64-way tie is hidden by the head-level mask), s1 = 130 neurons (100 dead + 30 live)
    sA = torch.tensor([[0.7, 129.2]] * L)
    rA = torch.tensor([[4.3, 1.5, 9.9]] * L)
    y0 = torch.full((L, 2), 0.3); p0 = torch.full((L, H), 0.2); z0 = torch.tensor(1.5)
    with torch.no_grad():
        mm.s.copy_(sA); mm.r.copy_(rA); mm.y.copy_(y0); mm.p.copy_(p0); mm.z.copy_(z0)
    ns.uvc_utils.prune_w_mask(mm)
    masksA = dict(w1=[m.mask[0].clone() for m in uvc_layers["W1"]], w3=[m.mask[0].clone() for m in uvc_layers["W3"]],
                  w2=[m.mask[:, 0].clone() for m in uvc_layers["W2"]])
    W1 = [m.weight.detach().clone() for m in uvc_layers["W1"]]; W3 = [m.weight.detach().clone() for m in uvc_layers["W3"]]
    m1, m3 = ao.masks(W1, W3, sA, rA, d, W3[0].shape[1])
    for l in range(L):      # the restatement (ties -> lower index) agrees wherever the result is defined
        assert torch.equal(masksA["w1"][l].bool(), m1[l]) and torch.equal(masksA["w3"][l].bool(), m3[l]), l
        assert int((masksA["w3"][l] == 0).sum()) == 130 and int((masksA["w1"][l] == 0).sum()) == 64 + 5 + 2

    class FakeOpt:
        param_groups = [{"lr": SPEC["lr"]}]
    gate = model.block_skip_gating.detach().clone()
    st = dict(s=sA.clone(), r=rA.clone(), y=y0.clone(), p=p0.clone(), z=z0.clone(), gate=gate, gate_buf=[])
    macs = torch.Tensor(flops_list[1]); full = float((flops_list[0] + macs.sum()) * 2)
    hp = dict(lr=SPEC["lr"], slr=args.slr, rlr=args.rlr, ylr=args.ylr, plr=args.plr, zlr=float(args.zlr_schedule_list[0]), budget=args.budget,
              z_grad_clip=args.z_grad_clip, sl2wd=args.sl2wd, gating_weight=args.gating_weight, d=d, Fh=W3[0].shape[1], macs=macs,
              embed_macs=flops_list[0], full=full, use_gumbel=True, eps=args.eps, gating_interval=args.gating_interval)
    traj, glist = [], []
    for step in range(SPEC["steps"]):
        gg = torch.zeros(L, 2)
        model.block_skip_gating.grad = gg.clone()
        torch.manual_seed(900 + step)
        n1 = -torch.empty(L, 2).exponential_().log(); n2 = -torch.empty(L, 2).exponential_().log()
        torch.manual_seed(900 + step)
        mm.update_gating()
        cur, s_np, r_np, g_np, glist = ns.uvc_optimizer.uvc_optimizer(FakeOpt(), mm, s_opt, r_opt, g_opt, dual_opt, args, {}, [], flops_list,
                                                                    args.z_grad_clip, step, args.gating_interval, glist)
        hp["global_step"] = step
        cur2 = ao.step(st, W1, W3, hp, n1, n2, gate_grad=gg, gate_sgd=lambda g: None)
        for name, a, b in [("s", mm.s, st["s"]), ("r", mm.r, st["r"]), ("y", mm.y, st["y"]), ("p", mm.p, st["p"]), ("z", mm.z, st["z"])]:
            torch.testing.assert_close(a.detach(), b, rtol=2e-5, atol=1e-6, msg=lambda m: f"ties step {step} {name}: {m}")
        assert abs(cur - cur2) < 1e-6
        traj.append(dict(cur=cur, s=mm.s.detach().clone(), r=mm.r.detach().clone(), y=mm.y.detach().clone(), p=mm.p.detach().clone(), z=mm.z.detach().clone(),
                         noise1=n1, noise2=n2, w1_sum=[fx.checksum(m.weight) for m in uvc_layers["W1"]], w3_sum=[fx.checksum(m.weight) for m in uvc_layers["W3"]]))
        print(f"  ties step {step}: resource={cur:.6f} s={mm.s.detach().flatten().tolist()}")
    # case B: the boundary cuts the 100 dead neurons (k = 60) and the dead head's dims only (already covered); stored for the record
    with torch.no_grad():
        mm.s.copy_(torch.tensor([[0.7, 59.5]] * L)); mm.r.copy_(rA)
    ns.uvc_utils.prune_w_mask(mm)
    masksB = dict(w3=[m.mask[0].clone() for m in uvc_layers["W3"]])
    zero3 = torch.arange(5, 705, 7)[:100]
    for l in range(L):
        off = masksB["w3"][l] == 0
        assert int(off.sum()) == 60 and bool(off[zero3].sum() == 60)          # all 60 come out of the tie group; which 60 is the unspecified part
    out = dict(spec=SPEC, args={k: v for k, v in vars(args).items()}, init=dict(s=sA, r=rA, y=y0, p=p0, z=z0), masksA=masksA, traj=traj,
               caseB=dict(s=torch.tensor([[0.7, 59.5]] * L), masks_w3=masksB["w3"], tie_group=zero3, k=60), full=full)
    torch.save(out, os.path.join(OUT, "admm_ties.pt"))
    print("  admm_ties written")


if __name__ == "__main__":
    main()
