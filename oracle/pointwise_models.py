"""Oracle: the two pointwise (BCE) victims, MF and NCF/NeuMF, on CPU (torch fp32).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates:
  * ``MF.__init__/forward/train_step``   recad/model/victim/mf.py:10-69
  * ``NCF.__init__/forward/train_step``  recad/model/victim/ncf.py:9-153
    (only the default 'NeuMF-end' variant with dropout 0 is on the hot path)
  * loss: nn.BCEWithLogitsLoss (mean); optimiser torch.optim.Adam(lr)
"""
import numpy as np
import torch
import torch.nn.functional as F


def _t(x, dtype=torch.float32):
    return torch.as_tensor(np.asarray(x), dtype=dtype).clone()


class MFOracle:
    """y = <U_u, V_i> + b_u + b_i + mean, mean = factor_num constant (mf.py:26,40-47)."""

    def __init__(self, user_emb, user_bias, item_emb, item_bias, mean=3.0, lr=1e-3):
        self.P = [torch.nn.Parameter(_t(a)) for a in (user_emb, user_bias, item_emb, item_bias)]
        self.mean = float(mean)
        # parameter order of the reference module: user_emb, user_bias, item_emb, item_bias
        self.optimizer = torch.optim.Adam(self.P, lr=lr)

    def forward(self, users, items):
        U, bu, V, bi = self.P
        users, items = _t(users, torch.int64), _t(items, torch.int64)
        return (U[users] * V[items]).sum(1) + bu[users].squeeze(-1) + bi[items].squeeze(-1) + self.mean

    def step(self, users, items, labels):
        pred = self.forward(users, items)
        loss = F.binary_cross_entropy_with_logits(pred, _t(labels))
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        return loss.item()

    def train_epoch(self, batches):
        tot, k = 0.0, 0
        for u, i, y in batches:
            tot += self.step(u, i, y)
            k += 1
        return tot / k


class NCFOracle:
    """NeuMF-end: concat(GMF u*i [f], MLP(cat(u', i')) [f]) -> Linear(2f, 1).

    MLP layer l (0-based) is Linear(f*2^(L-l), f*2^(L-l-1)) + ReLU (ncf.py:41-47);
    MLP embeddings have width f*2^(L-1) (ncf.py:34-39).
    ``params`` is a dict with keys: ug, ig, um, im, W (list), b (list), Wp, bp."""

    def __init__(self, params, lr=1e-3):
        self.ug, self.ig, self.um, self.im = (torch.nn.Parameter(_t(params[k])) for k in ("ug", "ig", "um", "im"))
        self.W = [torch.nn.Parameter(_t(w)) for w in params["W"]]
        self.b = [torch.nn.Parameter(_t(b)) for b in params["b"]]
        self.Wp, self.bp = torch.nn.Parameter(_t(params["Wp"])), torch.nn.Parameter(_t(params["bp"]))
        # nn.Module.parameters() order of the reference: the four embeddings,
        # then MLP (W0, b0, W1, b1, ...), then predict (W, b)
        plist = [self.ug, self.ig, self.um, self.im]
        for w, b in zip(self.W, self.b):
            plist += [w, b]
        plist += [self.Wp, self.bp]
        self.optimizer = torch.optim.Adam(plist, lr=lr)

    def forward(self, users, items):
        users, items = _t(users, torch.int64), _t(items, torch.int64)
        gmf = self.ug[users] * self.ig[items]
        h = torch.cat((self.um[users], self.im[items]), -1)
        for w, b in zip(self.W, self.b):
            h = torch.relu(F.linear(h, w, b))
        return F.linear(torch.cat((gmf, h), -1), self.Wp, self.bp).view(-1)

    def step(self, users, items, labels):
        pred = self.forward(users, items)
        loss = F.binary_cross_entropy_with_logits(pred, _t(labels))
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        return loss.item()

    def train_epoch(self, batches):
        tot, k = 0.0, 0
        for u, i, y in batches:
            tot += self.step(u, i, y)
            k += 1
        return tot / k
