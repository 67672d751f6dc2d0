"""TEST INFRASTRUCTURE (oracle): CPU restatement of the AUSH attacker's training step and fake-profile generation
(SURVEY.md 8f row 4).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this.

Follows recad/model/attacker/aush.py and the batch generator of recad/dataset/explicit.py, dense numpy float32 throughout
(the reference is dense torch float32):
  * batches (explicit.py:166-188 + utils.py:192-196 filler_filter_mat): users with at least filler_num rated items outside
    selected_ids + target_id_list, np.random.permutation, consecutive slices of batch_size;
  * sample_fillers (aush.py:59-76): per row np.random.choice(filler_num, with replacement) from
    list(set(rated columns) & filler_pool) -- the list order is CPython's set order, taken from the same expression;
  * ZR mask (aush.py:111-117): (real == 0) on the selected columns, np.random.shuffle of the argwhere'd pool, the first
    floor(len * (1 - ZR_ratio)) entries leave the mask;
  * generator (aush.py:254-266) Linear(I,128)-Sigmoid-Linear(128,I)-Sigmoid x 5, evaluated and DETACHED (aush.py:126-128):
    no loss reaches its parameters, G_optimizer.step() (aush.py:169) finds no gradient and the generator never changes
    (pinned: tests/golden/make_golden_aush.py prints "generator moved by 0.0");
  * discriminator (aush.py:269-283) 3 x (Linear-Sigmoid) of width 150 + Linear(150,1)-Sigmoid; one Adam step
    (torch.optim.Adam defaults, lr_d) per batch on d_loss = (BCE(D(real m), 1) + BCE(D(fake m), 0)) / 2 with
    m = fillers_mask + selects_mask (aush.py:134-146);
  * reported per batch (aush.py:150-166): BCE(D_updated(fake m), 1), MSE(fake sel, 5 sel), MSE(fake sel ZR, template sel ZR);
    train_step returns the means over the batches in the order (d_loss, g_loss_rec, g_loss_shilling, g_loss_gan);
  * generate_fake (aush.py:182-230).
Pinned against the live reference by tests/test_oracle_golden.py on tests/golden/aush_synth.npz.
"""
import math

import numpy as np

F32 = np.float32


def sigmoid(x):
    return (F32(1) / (F32(1) + np.exp(-x, dtype=F32))).astype(F32)


def eligible_rows(train_mat, selected_ids, target_id_list, filler_num):
    """utils.py:192-196."""
    rated = (train_mat > 0).astype(np.float64)
    rated[:, list(selected_ids) + list(target_id_list)] = 0
    return np.where(rated.sum(1) >= filler_num)[0]


def draw_fillers(real, n_items, selected_ids, target_id_list, filler_num):
    """aush.py:59-76 on the GLOBAL numpy generator."""
    pool = set(range(n_items)) - set(selected_ids) - set(target_id_list)
    mask = np.zeros_like(real)
    for b, row in enumerate(real):
        cand = list(set(np.argwhere(row > 0).flatten()) & pool)
        mask[b, np.random.choice(size=filler_num, replace=True, a=cand)] = 1
    return mask


def draw_zr(real, selects_mask, zr_ratio):
    """aush.py:111-117."""
    zr = (real == 0) * selects_mask
    pool = np.argwhere(zr)
    np.random.shuffle(pool)
    pool = pool[: math.floor(len(pool) * (1 - zr_ratio))]
    zr[pool[:, 0], pool[:, 1]] = 0
    return zr


class Net:
    """A stack of Linear + Sigmoid layers in float32 with a hand-written backward."""

    def __init__(self, weights, biases):
        self.W = [np.array(w, dtype=F32) for w in weights]
        self.b = [np.array(b, dtype=F32) for b in biases]

    def forward(self, x, keep=False):
        acts = [x.astype(F32)]
        for W, b in zip(self.W, self.b):
            acts.append(sigmoid(acts[-1] @ W.T + b))
        return (acts[-1], acts) if keep else acts[-1]

    def backward(self, acts, d_out):
        """d_out = d loss / d (last activation); returns gradients of every W, b."""
        gW, gb = [None] * len(self.W), [None] * len(self.W)
        d = d_out
        for l in range(len(self.W) - 1, -1, -1):
            dz = d * acts[l + 1] * (F32(1) - acts[l + 1])
            gW[l] = dz.T @ acts[l]
            gb[l] = dz.sum(0)
            d = dz @ self.W[l]
        return gW, gb


def bce(p, y):
    """nn.BCELoss (mean): logs clamped at -100."""
    lp = np.maximum(np.log(p, dtype=F32), F32(-100))
    lq = np.maximum(np.log(F32(1) - p, dtype=F32), F32(-100))
    return F32(-(y * lp + (F32(1) - y) * lq).mean(dtype=F32))


def bce_grad(p, y):
    """d mean-BCE / d p (torch: (p - y) / max(p (1 - p), 1e-12) / n)."""
    return ((p - y) / np.maximum(p * (F32(1) - p), F32(1e-12)) / F32(p.size)).astype(F32)


class Adam:
    """torch.optim.Adam defaults (betas 0.9 / 0.999, eps 1e-8, no weight decay), _single_tensor_adam arithmetic."""

    def __init__(self, params, lr):
        self.params, self.lr, self.t = params, lr, 0
        self.m = [np.zeros_like(p) for p in params]
        self.v = [np.zeros_like(p) for p in params]

    def step(self, grads):
        self.t += 1
        bc1, bc2 = 1 - 0.9 ** self.t, 1 - 0.999 ** self.t
        for p, g, m, v in zip(self.params, grads, self.m, self.v):
            m += F32(1 - 0.9) * (g - m)
            v *= F32(0.999)
            v += F32(1 - 0.999) * g * g
            p -= F32(self.lr / bc1) * (m / (np.sqrt(v) / F32(math.sqrt(bc2)) + F32(1e-8)))


class AushOracle:
    def __init__(self, train_mat, G_state, D_state, selected_ids=(62,), filler_num=36, attack_num=50, ZR_ratio=0.2, batch_size=256,
                 lr_d=0.001):
        self.train_mat = np.asarray(train_mat, dtype=F32)
        self.n_items = self.train_mat.shape[1]
        self.selected_ids, self.filler_num, self.attack_num = list(selected_ids), filler_num, attack_num
        self.ZR_ratio, self.batch_size = ZR_ratio, batch_size
        self.G = Net([G_state["main.0.weight"], G_state["main.2.weight"]], [G_state["main.0.bias"], G_state["main.2.bias"]])
        self.D = Net([D_state[f"main.{2 * l}.weight"] for l in range(4)], [D_state[f"main.{2 * l}.bias"] for l in range(4)])
        self.opt = Adam(self.D.W + self.D.b, lr_d)

    def batches(self, target_id_list):
        idx = eligible_rows(self.train_mat, self.selected_ids, target_id_list, self.filler_num)
        idx = np.random.permutation(idx)
        for s in range(0, len(idx), self.batch_size):
            yield idx[s:s + self.batch_size]

    def train_step(self, target_id_list):
        out = []
        for users in self.batches(target_id_list):
            real = self.train_mat[users].astype(F32)
            fillers = draw_fillers(real, self.n_items, self.selected_ids, target_id_list, self.filler_num)
            selects = np.zeros_like(fillers)
            selects[:, self.selected_ids] = 1
            patch = np.zeros_like(fillers)
            patch[:, self.selected_ids] = 5                       # aush.py:109: the SELECTED columns, not the targets
            zr = draw_zr(real, selects, self.ZR_ratio).astype(F32)
            template = real * fillers
            fake = template + self.G.forward(template) * F32(5) * selects + patch
            m = fillers + selects
            n = len(users)
            ones, zeros = np.ones((n, 1), dtype=F32), np.zeros((n, 1), dtype=F32)
            # discriminator step
            p_real, acts_r = self.D.forward(real * m, keep=True)
            p_fake, acts_f = self.D.forward(fake * m, keep=True)
            d_loss = F32(0.5) * (bce(p_real, ones) + bce(p_fake, zeros))
            gWr, gbr = self.D.backward(acts_r, F32(0.5) * bce_grad(p_real, ones))
            gWf, gbf = self.D.backward(acts_f, F32(0.5) * bce_grad(p_fake, zeros))
            self.opt.step([a + b for a, b in zip(gWr, gWf)] + [a + b for a, b in zip(gbr, gbf)])
            # what the generator phase reports
            gan = bce(self.D.forward(fake * m), ones)
            shilling = F32(((fake * selects - selects * F32(5)) ** 2).mean(dtype=F32))
            rec = F32(((fake * selects * zr - selects * template * zr) ** 2).mean(dtype=F32))
            out.append((float(d_loss), float(rec), float(shilling), float(gan)))
        return tuple(np.mean([o[k] for o in out]) for k in range(4))

    def generate_fake(self, target_id_list):
        idx = eligible_rows(self.train_mat, self.selected_ids, target_id_list, self.filler_num)
        idx = np.random.permutation(idx)
        idx = idx[np.random.randint(0, len(idx), self.attack_num)]
        real = self.train_mat[idx]
        fillers = draw_fillers(real, self.n_items, self.selected_ids, target_id_list, self.filler_num)
        selects = np.zeros_like(fillers)
        selects[:, self.selected_ids] = 1
        patch = np.zeros_like(fillers)
        patch[:, list(target_id_list)] = 5
        template = real * fillers
        fake = template + self.G.forward(template) * F32(5) * selects + patch
        sel = np.round(fake[:, self.selected_ids])
        fake[:, self.selected_ids] = np.clip(sel, 1, 5)
        return fake
