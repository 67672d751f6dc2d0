"""TEST INFRASTRUCTURE (oracle): CPU restatement of the WMF surrogate the AIA / Leg-UP attackers retrain once per
attack step.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this.

Follows recad/model/attacker/aia.py:
  * WeightedMF (aia.py:222-247): P [n_rows, dim], Q [n_items, dim] ~ N(0, 0.1), Q drawn FIRST; prediction P Q^T.
  * BaseTrainer.weighted_mse_loss (aia.py:283-289): weights = weight_pos where data > 0 else weight_neg;
    loss = sum_rows sum_items w (data - logits)^2.
  * WMFTrainer.fit_adv (aia.py:431-489): epochs 1 .. epoch_num - unroll_steps are plain torch.optim.Adam steps
    (lr, weight_decay added to the gradient) over np.random.shuffle'd row batches of `batch_size`, data DETACHED;
    the last `unroll_steps` epochs run inside `higher.innerloop_ctx(model, optimizer)`: the same loop with a
    differentiable optimizer, data NOT detached, so the returned predictions carry a graph back to `data`.

`higher` (facebookresearch/higher, requirements.txt pins no version; 0.2.1 is the last release) is absent from this
image and from /root/reference.  Its DifferentiableAdam.\\_update (higher/optim.py) is restated below from the published
source: it continues from the torch optimizer's state (exp_avg, exp_avg_sq, step) and uses the OLDER Adam form
    g += weight_decay * p;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g g;
    p -= lr * sqrt(1 - b2^t) / (1 - b1^t) * m / (sqrt(v) + eps)
(eps is added BEFORE the bias correction, unlike torch.optim.Adam).  PARITY of the unrolled epochs is UNPINNED (the
reference cannot run them here); the plain epochs are pinned against the live reference code by
tests/golden/make_golden_wmf.py (unroll_steps = 0, with an import stub for `higher` that is never exercised).
"""
import math

import numpy as np
import torch


def init_wmf(n_rows, n_items, dim):
    """aia.py:230-236: consumes the GLOBAL torch CPU generator, Q first."""
    Q = torch.zeros([n_items, dim]).normal_(mean=0, std=0.1)
    P = torch.zeros([n_rows, dim]).normal_(mean=0, std=0.1)
    return P, Q


def weighted_mse(data, logits, weight_pos, weight_neg):
    w = torch.ones_like(data) * weight_neg
    w[data > 0] = weight_pos
    return (w * (data - logits) ** 2).sum(1)


def epoch_orders(n_rows, n_epochs):
    """The row orders of fit_adv: ONE idx_list shuffled in place at the start of every epoch (aia.py:443-447, 468)."""
    idx = np.arange(n_rows)
    out = []
    for _ in range(n_epochs):
        np.random.shuffle(idx)
        out.append(idx.copy())
    return out


def fit_adv(data, epoch_num, unroll_steps, dim=16, lr=1e-2, weight_decay=1e-5, batch_size=16, weight_pos=1.0, weight_neg=0.0,
            P0=None, Q0=None, orders=None, betas=(0.9, 0.999), eps=1e-8):
    """-> (predictions [n_rows, n_items] (with a graph to `data` when it requires grad and unroll_steps > 0), P, Q).
    P0 / Q0: initial factors (default: drawn like the reference); orders: list of epoch_num row orders (default: drawn
    from np.random like the reference)."""
    n_rows, n_items = data.shape
    if P0 is None:
        P0, Q0 = init_wmf(n_rows, n_items, dim)
    if orders is None:
        orders = epoch_orders(n_rows, epoch_num)
    b1, b2 = betas
    P, Q = P0.clone().double().float(), Q0.clone().double().float()
    mP, vP, mQ, vQ = (torch.zeros_like(t) for t in (P, P, Q, Q))
    step = 0
    n_plain = epoch_num - unroll_steps
    with torch.no_grad():
        D = data.detach()
        for e in range(n_plain):                                    # torch.optim.Adam._single_tensor_adam
            idx = orders[e]
            for s in range(0, n_rows, batch_size):
                b = torch.as_tensor(idx[s:s + batch_size], dtype=torch.long)
                Pb = P[b]
                w = torch.where(D[b] > 0, torch.full_like(D[b], weight_pos), torch.full_like(D[b], weight_neg))
                R = w * (D[b] - Pb @ Q.t())
                gP = torch.zeros_like(P)
                gP.index_add_(0, b, -2.0 * (R @ Q))
                gQ = -2.0 * (R.t() @ Pb)
                step += 1
                bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
                for p, g, m, v in ((Q, gQ, mQ, vQ), (P, gP, mP, vP)):
                    g = g + weight_decay * p
                    m.lerp_(g, 1 - b1)
                    v.mul_(b2).addcmul_(g, g, value=1 - b2)
                    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
                    p.addcdiv_(m, denom, value=-(lr / bc1))
    # unrolled epochs: functional, differentiable w.r.t. `data` (higher.optim.DifferentiableAdam)
    for e in range(n_plain, epoch_num):
        idx = orders[e]
        for s in range(0, n_rows, batch_size):
            b = torch.as_tensor(idx[s:s + batch_size], dtype=torch.long)
            Db = data[b]
            w = torch.where(Db.detach() > 0, torch.full_like(Db, weight_pos), torch.full_like(Db, weight_neg)).detach()
            R = w * (Db - P[b] @ Q.t())
            gPb = -2.0 * (R @ Q)
            gP = torch.zeros_like(P).index_add(0, b, gPb)
            gQ = -2.0 * (R.t() @ P[b])
            step += 1
            bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
            step_size = lr * math.sqrt(bc2) / bc1
            new = []
            for p, g, m, v in ((Q, gQ, mQ, vQ), (P, gP, mP, vP)):
                g = g + weight_decay * p
                m = m * b1 + (1 - b1) * g
                v = v * b2 + (1 - b2) * g * g
                p = p - step_size * m / (v.sqrt() + eps)
                new.append((p, m, v))
            (Q, mQ, vQ), (P, mP, vP) = new
    return P @ Q.t(), P, Q
