"""WMF surrogate of the attack loop on the CUDA path: the plain epochs against a golden run of the
reference's own WMFTrainer (tests/golden/make_golden_wmf.py), the unrolled epoch and its reverse pass against the
oracle's restatement of higher's differentiable Adam (oracle/wmf.py; parity of that phase is UNPINNED in the reference:
`higher` is not installable here)."""
import os

import numpy as np
import pytest
import torch

from oracle import wmf as owmf
from tests import util

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _trainer(n_rows, n_items, dim, batch, wpos, wneg, wd):
    from recad_b200 import surrogate
    return surrogate.WMFTrainer(n_users=n_rows, n_items=n_items, device=torch.device(DEV), hidden_dim=dim, lr=1e-2, weight_decay=wd,
                                batch_size=batch, weight_pos=wpos, weight_neg=wneg)


@pytest.mark.parametrize("case", ["a", "b"])
@pytest.mark.parametrize("cluster", ["8", "1"])
def test_plain_epochs_match_the_reference_trainer(case, cluster, monkeypatch):
    monkeypatch.setenv("RECAD_WMF_CLUSTER", cluster)
    z = util.load("wmf_plain.npz")
    dim, batch, epochs, wpos, wneg, wd = z[f"{case}_hp"]
    data = torch.tensor(z[f"{case}_data"], requires_grad=True)
    tr = _trainer(*data.shape, int(dim), int(batch), float(wpos), float(wneg), float(wd))
    # the epoch shuffles come from np.random exactly where the reference takes them
    np.random.set_state(("MT19937", z[f"{case}_np_key"], int(z[f"{case}_np_pos"]), 0, 0.0))
    pred = tr.fit_adv(data, int(epochs), 0, init=(torch.tensor(z[f"{case}_P0"]), torch.tensor(z[f"{case}_Q0"])))
    for got, want in ((pred, z[f"{case}_pred"]), (tr.P, z[f"{case}_P"]), (tr.Q, z[f"{case}_Q"])):
        got = got.detach().cpu().numpy()
        assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max() + 1e-6, np.abs(got - want).max()


def test_initial_factors_follow_the_reference_generator_order():
    from recad_b200 import surrogate
    torch.manual_seed(7)
    P, Q = owmf.init_wmf(70, 40, 16)
    torch.manual_seed(7)
    tr = surrogate.WMFTrainer(n_users=70, n_items=40, device=torch.device(DEV), hidden_dim=16, lr=1e-2, weight_decay=1e-5,
                              batch_size=16, weight_pos=1.0, weight_neg=0.0)
    P2, Q2 = tr._initialize()
    z = util.load("wmf_plain.npz")
    assert torch.equal(P, P2) and torch.equal(Q, Q2) and np.array_equal(P.numpy(), z["a_P0"]) and np.array_equal(Q.numpy(), z["a_Q0"])


@pytest.mark.parametrize("n_rows,n_items,dim,batch,epochs,unroll,wpos,wneg,wd", [
    (70, 40, 16, 16, 3, 1, 1.0, 0.0, 1e-5),          # the reference's defaults (default.py:177-184) at toy size
    (45, 130, 8, 7, 3, 2, 2.0, 0.5, 1e-3),           # ragged batches, two unrolled epochs, both weights
    (64, 300, 32, 32, 2, 1, 1.0, 0.1, 0.0),          # widest supported tile, no decay
])
def test_unrolled_epochs_and_reverse_pass_match_the_oracle(n_rows, n_items, dim, batch, epochs, unroll, wpos, wneg, wd):
    rng = np.random.default_rng(n_rows + n_items)
    data = ((rng.random((n_rows, n_items)) < 0.2) * rng.integers(1, 6, (n_rows, n_items))).astype(np.float32)
    n_fake = 6
    data[-n_fake:] = (rng.random((n_fake, n_items)) * 5 * (rng.random((n_fake, n_items)) < 0.7)).astype(np.float32)
    torch.manual_seed(3)
    P0, Q0 = owmf.init_wmf(n_rows, n_items, dim)
    np.random.seed(9)
    orders = owmf.epoch_orders(n_rows, epochs)
    up = torch.tensor(rng.standard_normal((n_rows, n_items)).astype(np.float32))      # d loss / d predictions
    # oracle: torch CPU autograd through the functional restatement
    d_ref = torch.tensor(data, requires_grad=True)
    pred_ref, _, _ = owmf.fit_adv(d_ref, epochs, unroll, dim=dim, lr=1e-2, weight_decay=wd, batch_size=batch, weight_pos=wpos,
                                  weight_neg=wneg, P0=P0, Q0=Q0, orders=orders)
    (pred_ref * up).sum().backward()
    # CUDA path
    d_gpu = torch.tensor(data, device=DEV, requires_grad=True)
    tr = _trainer(n_rows, n_items, dim, batch, wpos, wneg, wd)
    pred = tr.fit_adv(d_gpu, epochs, unroll, init=(P0, Q0), orders=np.stack(orders))
    (pred * up.to(DEV)).sum().backward()
    pr, pg = pred_ref.detach().numpy(), pred.detach().cpu().numpy()
    assert np.abs(pg - pr).max() <= 1e-4 * np.abs(pr).max() + 1e-6, np.abs(pg - pr).max()
    gr, gg = d_ref.grad.numpy(), d_gpu.grad.cpu().numpy()
    assert np.abs(gr).max() > 0
    assert np.abs(gg - gr).max() <= 2e-3 * np.abs(gr).max(), (np.abs(gg - gr).max(), np.abs(gr).max())
    # the attacker only uses the fake rows' gradient (aia.py:128-135: the genuine rows are constants)
    assert np.abs(gg[-n_fake:] - gr[-n_fake:]).max() <= 2e-3 * np.abs(gr[-n_fake:]).max()
    # rows trained with weight_neg = 0 get no gradient where nothing was observed
    if wneg == 0.0:
        assert np.all(gg[data <= 0] == 0)


def test_cluster_size_does_not_change_the_bits_of_a_run(monkeypatch):
    """warp -> CTA -> cluster sums run in a fixed order: one cluster size = one result, run after run."""
    rng = np.random.default_rng(5)
    data = ((rng.random((50, 90)) < 0.3) * rng.integers(1, 6, (50, 90))).astype(np.float32)
    torch.manual_seed(1)
    P0, Q0 = owmf.init_wmf(50, 90, 16)
    np.random.seed(2)
    orders = np.stack(owmf.epoch_orders(50, 4))
    outs = []
    for _ in range(2):
        monkeypatch.setenv("RECAD_WMF_CLUSTER", "8")
        tr = _trainer(50, 90, 16, 16, 1.0, 0.0, 1e-5)
        outs.append(tr.fit_adv(torch.tensor(data, device=DEV, requires_grad=True), 4, 1, init=(P0, Q0), orders=orders).detach().clone())
    assert torch.equal(outs[0], outs[1])
