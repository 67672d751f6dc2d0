"""Oracle: LightGCN propagation, BPR step and dense Adam on CPU (torch fp32).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates:
  * ``LightGCN.computer``      recad/model/victim/lightgcn.py:82-113
  * ``LightGCN.__dropout_x``   recad/model/victim/lightgcn.py:62-72 (graph dropout, off by default)
  * ``LightGCN.getEmbedding``  recad/model/victim/lightgcn.py:122-130
  * ``LightGCN.train_step``    recad/model/victim/lightgcn.py:132-172
  * ``LightGCN.forward``       recad/model/victim/lightgcn.py:174-183
  * optimiser: torch.optim.Adam(lr), defaults betas (0.9, 0.999), eps 1e-8
    (recad/utils.py:181-189, lightgcn.py:17-19)

Two forms are given: ``*_autograd`` follows the reference op by op and lets
torch differentiate; ``*_manual`` is the closed form of SURVEY.md a-M that the
CUDA kernels implement (Horner backward with the symmetric A_hat, fused Adam).
tests/test_oracle_golden.py checks both against the live reference's output.
"""
import numpy as np
import torch


def csr_to_torch(indptr, indices, data, n):
    return torch.sparse_csr_tensor(
        torch.as_tensor(np.asarray(indptr), dtype=torch.int64),
        torch.as_tensor(np.asarray(indices), dtype=torch.int64),
        torch.as_tensor(np.asarray(data), dtype=torch.float32),
        size=(n, n),
    )


def csr_to_torch_coo(indptr, indices, data, n):
    """The layout the reference itself holds (coalesced COO, implicit.py:295-296)."""
    indptr = np.asarray(indptr)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    idx = torch.as_tensor(np.stack([rows, np.asarray(indices, dtype=np.int64)]))
    return torch.sparse_coo_tensor(idx, torch.as_tensor(np.asarray(data), dtype=torch.float32), (n, n)).coalesce()


def dropout_graph(graph_coo, keep_prob):
    """__dropout_x (lightgcn.py:62-72) on the coalesced COO graph: one torch.rand(nnz) draw of the global CPU generator;
    an entry stays when int(rand + keep_prob) is 1 and is rescaled by 1 / keep_prob."""
    index, values = graph_coo.indices().t(), graph_coo.values()
    keep = (torch.rand(len(values)) + keep_prob).int().bool()
    return torch.sparse_coo_tensor(index[keep].t(), values[keep] / keep_prob, graph_coo.shape)


def propagate(graph, user_emb, item_emb, n_layers):
    """computer(): E0 = cat(U, I); E(k+1) = A_hat E(k); out = mean_k E(k)."""
    all_emb = torch.cat([user_emb, item_emb])
    embs = [all_emb]
    for _ in range(n_layers):
        all_emb = torch.sparse.mm(graph, all_emb)
        embs.append(all_emb)
    light_out = torch.mean(torch.stack(embs, dim=1), dim=1)
    return torch.split(light_out, [user_emb.shape[0], item_emb.shape[0]])


def bpr_loss(graph, user_emb, item_emb, n_layers, users, pos, neg, lam):
    """train_step body up to final_loss (lightgcn.py:137-165)."""
    all_users, all_items = propagate(graph, user_emb, item_emb, n_layers)
    u, p, n = all_users[users], all_items[pos], all_items[neg]
    u0, p0, n0 = user_emb[users], item_emb[pos], item_emb[neg]
    reg = 0.5 * (u0.norm(2).pow(2) + p0.norm(2).pow(2) + n0.norm(2).pow(2)) / float(len(users))
    pos_scores = torch.sum(u * p, dim=1)
    neg_scores = torch.sum(u * n, dim=1)
    loss = torch.mean(torch.nn.functional.softplus(neg_scores - pos_scores))
    return loss + lam * reg


class LightGCNOracle:
    """Holds the two tables + Adam exactly as the reference module does."""

    def __init__(self, graph, user_emb, item_emb, n_layers=3, lam=1e-4, lr=1e-3, keep_prob=None):
        """keep_prob: graph dropout as the reference applies it with config['dropout'] on and the module in training mode
        (graph must then be the coalesced COO tensor): every propagation draws a fresh mask (lightgcn.py:90-96)."""
        self.graph = graph
        self.keep_prob, self.training = keep_prob, True
        self.user_emb = torch.nn.Parameter(torch.as_tensor(user_emb, dtype=torch.float32).clone())
        self.item_emb = torch.nn.Parameter(torch.as_tensor(item_emb, dtype=torch.float32).clone())
        self.n_layers, self.lam = n_layers, lam
        self.optimizer = torch.optim.Adam([self.user_emb, self.item_emb], lr=lr)

    def step(self, users, pos, neg):
        users, pos, neg = (torch.as_tensor(np.asarray(x), dtype=torch.int64) for x in (users, pos, neg))
        loss = bpr_loss(self._graph(), self.user_emb, self.item_emb, self.n_layers, users, pos, neg, self.lam)
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        return loss.item()

    def _graph(self):
        return dropout_graph(self.graph, self.keep_prob) if self.keep_prob is not None and self.training else self.graph

    def train_epoch(self, batches):
        """train_step (lightgcn.py:132-172): returns the mean of per-batch final_loss."""
        total, k = 0.0, 0
        for users, pos, neg in batches:
            total += self.step(users, pos, neg)
            k += 1
        return total / k

    @torch.no_grad()
    def forward(self, users, items):
        """forward (lightgcn.py:174-183): <O_u, O_i>, no sigmoid."""
        au, ai = propagate(self._graph(), self.user_emb, self.item_emb, self.n_layers)
        users = torch.as_tensor(np.asarray(users), dtype=torch.int64)
        items = torch.as_tensor(np.asarray(items), dtype=torch.int64)
        return torch.sum(au[users] * ai[items], dim=1)

    @torch.no_grad()
    def final_embeddings(self):
        au, ai = propagate(self.graph, self.user_emb, self.item_emb, self.n_layers)
        return au, ai


# --------------------------------------------------------------------------- #
# closed form (what the kernels implement), numpy/torch without autograd
# --------------------------------------------------------------------------- #
def softplus(x):
    """torch.nn.functional.softplus, beta=1, threshold=20."""
    return torch.where(x > 20, x, torch.log1p(torch.exp(x)))


@torch.no_grad()
def manual_step(graph, E, m, v, step, users, pos, neg, n_users, n_layers, lam, lr,
                b1=0.9, b2=0.999, eps=1e-8):
    """One BPR step in closed form (SURVEY.md a-M).  E, m, v: [N, D] float32
    tensors updated in place; item ids are offset by n_users.  ``step`` is the
    1-based Adam step.  Returns final_loss (python float)."""
    users = torch.as_tensor(np.asarray(users), dtype=torch.int64)
    pos = torch.as_tensor(np.asarray(pos), dtype=torch.int64) + n_users
    neg = torch.as_tensor(np.asarray(neg), dtype=torch.int64) + n_users
    B = users.numel()
    L1 = n_layers + 1
    acc, X = E.clone(), E
    for _ in range(n_layers):
        X = torch.sparse.mm(graph, X)
        acc += X
    O = acc / L1
    x = (O[users] * O[neg]).sum(1) - (O[users] * O[pos]).sum(1)
    reg = 0.5 * ((E[users] ** 2).sum() + (E[pos] ** 2).sum() + (E[neg] ** 2).sum()) / B
    loss = softplus(x).mean() + lam * reg
    s = (torch.sigmoid(x) / B).unsqueeze(1)
    gO = torch.zeros_like(E)
    gO.index_add_(0, users, s * (O[neg] - O[pos]))
    gO.index_add_(0, pos, -s * O[users])
    gO.index_add_(0, neg, s * O[users])
    g = gO / L1
    t = g.clone()
    for _ in range(n_layers):
        t = g + torch.sparse.mm(graph, t)
    G = t
    for idx in (users, pos, neg):
        G.index_add_(0, idx, (lam / B) * E[idx])
    m.mul_(b1).add_(G, alpha=1 - b1)
    v.mul_(b2).addcmul_(G, G, value=1 - b2)
    # torch.optim.Adam (single-tensor form): denom = sqrt(v)/sqrt(bc2) + eps;
    # p -= (lr/bc1) * m/denom
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = v.sqrt() / (bc2 ** 0.5) + eps
    E.addcdiv_(m, denom, value=-(lr / bc1))
    return float(loss)
