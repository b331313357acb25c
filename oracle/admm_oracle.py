"""ORACLE (test infrastructure only) — CPU restatement of one ADMM step of the reference.

Restates `uvc_optimizer` (uvc_optimizer.py:37-144) with `prox_w` (uvc_utils.py:315-345), `weight_list_to_scores`
(:54-73), `LeastSsum` (:75-92), `sloss1/rloss1/srloss2/yloss/ploss/zloss` (:177-269), `calc_flops` (:409-471) and
`prune_w_mask` (:376-401) as closed forms over plain tensors, with the Gumbel noise passed in explicitly
(the reference draws it inside `calc_flops`).  Selections use `torch.topk(largest=False)` exactly like the
reference.  PINNED: `oracle/gen_golden_admm.py` drives the UNMODIFIED reference for several steps and asserts this
file reproduces its trajectory (s, r, y, p, z, gates, resource, weights); the trajectory is committed as
`tests/golden/admm_traj.pt`.
"""
import math

import torch


def scores_w1(W, d):
    """uvc_utils.py:57-69: per-column sq-norms [H,d] and per-head sq-norms [H] (separate reductions)."""
    H = W.shape[1] // d
    c1 = (W ** 2).sum(0).view(H, d)
    c2 = torch.stack([(W[:, h * d:(h + 1) * d].reshape(-1) ** 2).sum(0) for h in range(H)])
    return c1, c2


def scores_w3(W):
    return (W ** 2).sum(0)


def bottom_idx(v, k):
    return torch.topk(v, int(k), largest=False, sorted=False)[1]


def kth_smallest(v, k):
    """LeastSsum.backward factor (uvc_utils.py:79-90): the (k+1)-th smallest value, or the max if k+1 > numel."""
    idx = int(k) + 1
    if idx <= v.numel():
        return torch.topk(v, idx, largest=False, sorted=True)[0][-1]
    return v.max()


def prox(W1, W3, s, r, y, p, lr, d):
    """uvc_utils.py:315-345 (in place on the weight lists)."""
    S, R = torch.ceil(s), torch.ceil(r)
    for l, W in enumerate(W1):
        c1, c2 = scores_w1(W, d)
        for h in range(c1.shape[0]):
            idx = bottom_idx(c1[h], R[l, h])
            W[:, idx + h * d] /= (1.0 + 2.0 * lr * p[l, h].item())
        for h in bottom_idx(c2, S[l, 0]):
            W[:, h * d:(h + 1) * d] /= (1.0 + 2.0 * lr * y[l, 0].item())
    for l, W in enumerate(W3):
        idx = bottom_idx(scores_w3(W), S[l, 1])
        W[:, idx] /= (1.0 + 2.0 * lr * y[l, 1].item())


def flops_fraction(W1, S, R, gate, noise, macs, embed_macs, full, d, Fh, use_gumbel=True, gumbel_hard=False, eps=0.1):
    """calc_flops with full_model_flops given (uvc_utils.py:409-471).  Returns (flops, terms for the gradient)."""
    L, H = R.shape
    C = H * d
    rho0 = ((H - S[:, 0]) / H)
    rho1 = ((Fh - S[:, 1]) / Fh)
    a = torch.full((L,), float(C))
    pruned = torch.zeros(L, H, dtype=torch.bool)
    for l, W in enumerate(W1):
        _, c2 = scores_w1(W, d)
        idx = bottom_idx(c2, S[l, 0])
        pruned[l, idx] = True
        a[l] -= S[l, 0] * d
        for h in range(H):
            if not pruned[l, h]:
                a[l] -= R[l, h]
    rhor = a / C
    ok0, ok1, okr = [((v >= 0) & (v <= 1)).float() for v in (rho0, rho1, rhor)]
    rho0, rho1, rhor = rho0.clamp(0, 1), rho1.clamp(0, 1), rhor.clamp(0, 1)
    if gate is None:
        g = torch.ones(L); dg = torch.zeros(L, 2)
    elif use_gumbel:
        u = (gate + noise) / 0.5
        soft = u.softmax(1)
        g = soft[:, 1]
        dg1 = soft[:, 1] * (1 - soft[:, 1]) / 0.5
        dg = torch.stack([-dg1, dg1], 1)
        if gumbel_hard:
            g = (u[:, 1] > u[:, 0]).float()
    else:
        t = gate[:, 1] ** 2
        g = t / (t + eps)
        dg = torch.stack([torch.zeros(L), 2 * gate[:, 1] * eps / (t + eps) ** 2], 1)
    tm = macs * g.unsqueeze(1)
    total = embed_macs + (tm[:, 0] * rho0).sum() + (tm[:, 1] * rho0).sum() + (tm[:, 2] * rhor).sum() + (tm[:, 3] * rhor).sum() \
        + (tm[:, 4] * rho1).sum() + (tm[:, 5] * rho1).sum()
    flops = total * 2 / full
    terms = dict(g=g, dg=dg, rho0=rho0, rho1=rho1, rhor=rhor, ok0=ok0, ok1=ok1, okr=okr, pruned=pruned)
    return flops, terms


def step(state, W1, W3, hp, noise1, noise2, gate_grad=None, gate_sgd=None):
    """One uvc_optimizer call (non-warm-up unless hp['warmup']).  `state`: dict of s, r, y, p, z, gate (tensors, updated in
    place), gate_buf (list).  `hp`: lr, slr, rlr, ylr, plr, zlr, budget, z_grad_clip, sl2wd, gating_weight, d, Fh, macs,
    embed_macs, full, use_gumbel, eps, global_step, gating_interval, warmup.  `gate_sgd(grad)`: applies the gate optimiser
    step (momentum SGD) when the interval elapses.  Returns cur_resource."""
    s, r, y, p, z, gate = (state[k] for k in ("s", "r", "y", "p", "z", "gate"))
    d, Fh = hp["d"], hp["Fh"]
    L, H = r.shape
    C = H * d
    prox(W1, W3, s, r, y, p, hp["lr"], d)
    S, R = torch.ceil(s), torch.ceil(r)
    flops, t = flops_fraction(W1, S, R, gate, noise1, hp["macs"], hp["embed_macs"], hp["full"], d, Fh, hp["use_gumbel"], False, hp["eps"])
    cur = float(flops - hp["budget"]) + hp["budget"]
    if hp.get("warmup"):
        return cur
    sr = float(flops) - hp["budget"]
    passg = 1.0 if -hp["z_grad_clip"] <= sr <= hp["z_grad_clip"] else 0.0
    c = 2.0 / hp["full"]
    w01, w23, w45 = hp["macs"][:, 0] + hp["macs"][:, 1], hp["macs"][:, 2] + hp["macs"][:, 3], hp["macs"][:, 4] + hp["macs"][:, 5]
    gs1 = torch.zeros(L, 2); gr1 = torch.zeros(L, H)
    for l in range(L):
        c1, c2 = scores_w1(W1[l], d)
        gs1[l, 0] = y[l, 0] * kth_smallest(c2, S[l, 0])
        gs1[l, 1] = y[l, 1] * kth_smallest(scores_w3(W3[l]), S[l, 1])
        for h in range(H):
            gr1[l, h] = p[l, h] * kth_smallest(c1[h], R[l, h])
    s_ub = torch.tensor([float(H), float(Fh)]).expand(L, 2)
    gs1 = gs1 + hp["sl2wd"] * (s / s_ub)
    gr1 = gr1 + hp["sl2wd"] * (r / d)
    gs2 = torch.stack([c * t["g"] * (w01 * (-1.0 / H) * t["ok0"] + w23 * (-d / C) * t["okr"]),
                       c * t["g"] * (w45 * (-1.0 / Fh) * t["ok1"])], 1) * passg
    gr2 = (c * t["g"] * w23 * (-1.0 / C) * t["okr"]).unsqueeze(1) * (~t["pruned"]).float() * passg
    s_grad = gs1 + z * gs2
    r_grad = gr1 + z * gr2
    if gate is not None and gate_sgd is not None:
        inner = w01 * t["rho0"] + w23 * t["rhor"] + w45 * t["rho1"]
        gres = c * inner.unsqueeze(1) * t["dg"] * passg
        G = gate_grad + z * hp["gating_weight"] * gres
        state["gate_buf"].append(G.unsqueeze(0) * (hp["global_step"] % hp["gating_interval"]))
        if (hp["global_step"] + 1) % hp["gating_interval"] == 0:
            gate_sgd(torch.cat(state["gate_buf"]).mean(0))
            state["gate_buf"] = []

    def projected_sgd(v, g, vmax, lr):
        over, under = v >= vmax, v <= 0
        g = torch.where(over, g.clamp(min=0.0), g)
        g = torch.where(under, g.clamp(max=0.0), g)
        coef = min(1.0 / (float(g.abs().max()) + 1e-6), 1.0)
        v -= lr * (g * coef)
        v.clamp_(min=0.0)
        v[over] = vmax[over]

    projected_sgd(s, s_grad, (s_ub - 1 - 1e-8).clamp(min=0.0), hp["slr"])
    projected_sgd(r, r_grad, (torch.full((L, H), float(d)) - 1 - 1e-8).clamp(min=0.0), hp["rlr"])
    # dual ascent with the new s, r
    S, R = torch.ceil(s), torch.ceil(r)
    for l in range(L):
        c1, c2 = scores_w1(W1[l], d)
        y[l, 0] += hp["ylr"] * torch.topk(c2, int(S[l, 0]), largest=False)[0].sum()
        y[l, 1] += hp["ylr"] * torch.topk(scores_w3(W3[l]), int(S[l, 1]), largest=False)[0].sum()
        for h in range(H):
            p[l, h] += hp["plr"] * torch.topk(c1[h], int(R[l, h]), largest=False)[0].sum()
    flops2, _ = flops_fraction(W1, S, R, gate, noise2, hp["macs"], hp["embed_macs"], hp["full"], d, Fh, hp["use_gumbel"], False, hp["eps"])
    z += hp["zlr"] * (flops2 - hp["budget"])
    y.clamp_(min=0.0); p.clamp_(min=0.0); z.clamp_(min=0.0)
    return cur


def masks(W1, W3, s, r, d, Fh):
    """prune_w_mask selections (uvc_utils.py:376-401): boolean keep-masks over W1 columns, W3 columns (= W2 rows)."""
    S, R = torch.ceil(s), torch.ceil(r)
    m1, m3 = [], []
    for l, W in enumerate(W1):
        c1, c2 = scores_w1(W, d)
        keep = torch.ones(W.shape[1], dtype=torch.bool)
        for h in range(c1.shape[0]):
            keep[bottom_idx(c1[h], R[l, h]) + h * d] = False
        for h in bottom_idx(c2, S[l, 0]):
            keep[h * d:(h + 1) * d] = False
        m1.append(keep)
    for l, W in enumerate(W3):
        keep = torch.ones(W.shape[1], dtype=torch.bool)
        keep[bottom_idx(scores_w3(W), S[l, 1])] = False
        m3.append(keep)
    return m1, m3
