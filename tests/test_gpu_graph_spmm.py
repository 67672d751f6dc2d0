"""GPU parity: graph construction (bit-exact), in-place injection, SpMM + fused layer mean.
Everything goes through the C ABI (recad_b200.ops -> librecad_b200.so); the oracle is the checker."""
import numpy as np
import pytest
import torch

from oracle import graph as og
from oracle import lightgcn as olg
from recad_b200 import synthetic
from tests import util

pytestmark = pytest.mark.gpu
META = util.meta()


def _dev():
    return torch.device("cuda:0")


def _build(u, i, U, I, **kw):
    from recad_b200 import ops
    return ops.Graph.from_edges(torch.as_tensor(u, device=_dev()), torch.as_tensor(i, device=_dev()), U, I, **kw)


def _assert_same_csr(g, ptr, col, val):
    gp, gc, gv = g.to_numpy()
    assert np.array_equal(gp, ptr)
    assert np.array_equal(gc, col)
    assert gv.tobytes() == np.asarray(val, dtype=np.float32).tobytes()     # value BITS


def test_dev_graph_bit_exact_both_edge_sets():
    tr, va, te = util.dicts("dev")
    z = util.load("dev_graph.npz")
    U, I = META["dev"]["n_users"], META["dev"]["n_items"]
    g = _build(*og.graph_edges_reference(tr, va, te), U, I)
    _assert_same_csr(g, z["crow"], z["col"], z["val"])
    assert util.sha(*g.to_numpy()) == META["dev_graph"]["sha"]
    u, i, _, _ = og.flatten_dict(tr)
    _assert_same_csr(_build(u, i, U, I), z["crow_train"], z["col_train"], z["val_train"])


def test_game_and_ml1m_shaped_graph_hashes():
    tr, va, te = util.dicts("game")
    m = META["game"]
    g = _build(*og.graph_edges_reference(tr, va, te), m["n_users"], m["n_items"])
    assert util.sha(*g.to_numpy()) == m["graph_sha"]
    u, i, _, _ = og.flatten_dict(tr)
    g = _build(u, i, m["n_users"], m["n_items"])
    assert g.nnz == m["graph_train_nnz"] and util.sha(*g.to_numpy()) == m["graph_train_sha"]
    tr, _, _ = synthetic.make_splits(synthetic.ML1M, seed=0)
    u, i, _, _ = og.flatten_dict(tr)
    mm = META["ml1m_shaped"]
    rng = np.random.default_rng(1)
    perm = rng.permutation(len(u))                       # the builder must not depend on the input order
    g = _build(u[perm], i[perm], mm["n_users"], mm["n_items"])
    assert g.nnz == mm["graph_train_nnz"] and util.sha(*g.to_numpy()) == mm["graph_train_sha"]
    assert int(g.degree.max()) == mm["max_degree"]


@pytest.mark.parametrize("case", ["empty", "single", "duplicates", "isolated", "skewed"])
def test_graph_edge_cases_match_oracle(case):
    rng = np.random.default_rng(7)
    if case == "empty":
        U, I, u, i = 5, 4, np.zeros(0, np.int64), np.zeros(0, np.int64)
    elif case == "single":
        U, I, u, i = 1, 1, np.array([0]), np.array([0])
    elif case == "duplicates":
        U, I = 6, 5
        u, i = rng.integers(0, U, 200), rng.integers(0, I, 200)   # ~7 copies of every pair
    elif case == "isolated":
        U, I = 50, 40
        u, i = rng.integers(10, 20, 30), rng.integers(0, 5, 30)    # most rows have degree 0
    else:
        U, I = 300, 2000
        u = np.concatenate([np.zeros(1500, np.int64), rng.integers(0, U, 3000)])   # user 0 has > seg_len items
        i = np.concatenate([np.arange(1500), rng.integers(0, I, 3000)])
    g = _build(u, i, U, I, seg_len=64)
    ptr, col, val, _, deg = og.norm_adj_csr(np.asarray(u, np.int64), np.asarray(i, np.int64), U, I)
    _assert_same_csr(g, ptr, col, val)
    assert np.array_equal(g.degree.cpu().numpy(), deg)
    if case == "skewed":
        assert g.n_mrow > 0 and g.n_seg > g.n_rows


def test_bad_ids_fail_loudly():
    from recad_b200 import ops
    with pytest.raises(ops.RecadError):
        _build(np.array([0, 9]), np.array([0, 1]), 3, 3)


@pytest.mark.parametrize("with_edges", [False, True])
def test_append_users_equals_rebuild(with_edges):
    """In-place injection == building the injected graph from scratch (bit-exact)."""
    tr, va, te = util.dicts("dev")
    U, I = META["dev"]["n_users"], META["dev"]["n_items"]
    u, i, _, _ = og.flatten_dict(tr)
    g = _build(u, i, U, I)
    rng = np.random.default_rng(3)
    F = 50
    rows = [sorted(rng.choice(I, size=rng.integers(0, 40), replace=False).tolist()) if with_edges else [] for _ in range(F)]
    fake_rowptr = torch.tensor(np.concatenate([[0], np.cumsum([len(r) for r in rows])]), dtype=torch.int64)
    fake_items = torch.tensor([x for r in rows for x in r], dtype=torch.int32)
    g2 = g.append_users(U, I, fake_rowptr, fake_items)
    fu = np.concatenate([np.full(len(r), U + k, np.int64) for k, r in enumerate(rows)] + [np.zeros(0, np.int64)])
    fi = np.array([x for r in rows for x in r], dtype=np.int64)
    ptr, col, val, _, deg = og.norm_adj_csr(np.concatenate([u, fu]), np.concatenate([i, fi]), U + F, I)
    _assert_same_csr(g2, ptr, col, val)
    assert np.array_equal(g2.degree.cpu().numpy(), deg)


def test_array_dataset_injection_matches_oracle_graph():
    """ArrayImplicitData.inject_data (device-resident, the 1M-user path): the fake_array -> rows rule of
    implicit.py:107-114 (rating > filter_num, trailing empty rows are not users) and the extended graph."""
    from recad_b200 import dataset
    U, I = 900, 400
    u, i = synthetic.make_edges(U, I, 9000, seed=11)
    dev = torch.device("cuda:0")
    data = dataset.ArrayImplicitData("t", U, I, (torch.as_tensor(u, device=dev), torch.as_tensor(i, device=dev)), dev, prefetch=False)
    rng = np.random.default_rng(5)
    fake = rng.integers(0, 6, size=(12, I)).astype(np.float32)          # ratings 0..5, keep the 5s
    fake[3] = 0                                                         # an interior user with nothing left
    fake[10:] = 4                                                       # trailing rows filtered away entirely
    d2 = data.inject_data("explicit", fake, filter_num=4)
    assert d2.n_users == U + 10 and d2.n_items == I
    fr, fc = np.nonzero(fake > 4)
    ptr, col, val, _, deg = og.norm_adj_csr(np.concatenate([u, fr + U]), np.concatenate([i, fc]), U + 10, I)
    _assert_same_csr(d2.Graph, ptr, col, val)
    rp, rc = d2.train_csr()
    assert rp[-1] == len(u) + len(fr) and rp[U + 4] == rp[U + 3]        # user U + 3 has no positives
    assert data.n_users == U and data.Graph.n_rows == U + I              # the clean dataset is untouched


# ------------------------------------------------------------------ SpMM
def _rand_graph(U, I, E, seed, seg_len=256):
    u, i = synthetic.make_edges(U, I, E, seed=seed)
    return _build(u, i, U, I, seg_len=seg_len), og.norm_adj_csr(u, i, U, I)


@pytest.mark.parametrize("D", [8, 16, 32, 64, 128, 20, 256])
def test_spmm_matches_oracle(D):
    from recad_b200 import ops
    g, (ptr, col, val, _, _) = _rand_graph(700, 300, 20000, seed=D, seg_len=64)   # item rows ~67 long: multi-segment
    N = g.n_rows
    A = olg.csr_to_torch(ptr, col, val, N)
    torch.manual_seed(D)
    X = torch.randn(N, D)
    Cc = torch.randn(N, D)
    ref = torch.sparse.mm(A, X)
    Xd, Cd = X.to(_dev()), Cc.to(_dev())
    Y = torch.empty_like(Xd)
    Z = torch.empty_like(Xd)
    ops.spmm(g, Xd, Y, Cd, Z, 0.25)
    assert torch.allclose(Y.cpu(), ref, rtol=1e-5, atol=1e-6)
    assert torch.allclose(Z.cpu(), 0.25 * (Cc + ref), rtol=1e-5, atol=1e-6)
    Z2 = Cd.clone()                                    # Z aliasing C, no Y
    ops.spmm(g, Xd, None, Z2, Z2, 1.0)
    assert torch.allclose(Z2.cpu(), Cc + ref, rtol=1e-5, atol=1e-6)
    Y3 = torch.empty_like(Xd)                          # Y only
    ops.spmm(g, Xd, Y3, None, None, 1.0)
    assert torch.equal(Y3, Y)                          # deterministic: bitwise repeatable


def test_spmm_rectangular_shard():
    """A user-row shard [rows x N] (what the multi-GPU path multiplies)."""
    from recad_b200 import ops
    g, (ptr, col, val, _, _) = _rand_graph(400, 150, 9000, seed=5)
    lo, hi = 100, 260
    sub_ptr = torch.as_tensor(ptr[lo:hi + 1] - ptr[lo], device=_dev())
    sub_col = torch.as_tensor(col[ptr[lo]:ptr[hi]], device=_dev(), dtype=torch.int32)
    sub_val = torch.as_tensor(val[ptr[lo]:ptr[hi]], device=_dev())
    gs = ops.Graph.from_csr(sub_ptr, sub_col, sub_val, n_cols=g.n_rows)
    X = torch.randn(g.n_rows, 64)
    ref = torch.sparse.mm(olg.csr_to_torch(ptr, col, val, g.n_rows), X)[lo:hi]
    Y = torch.empty((hi - lo, 64), device=_dev())
    ops.spmm(gs, X.to(_dev()), Y)
    assert torch.allclose(Y.cpu(), ref, rtol=1e-5, atol=1e-6)


def test_spmm_linearity_at_scale():
    """Size-independent property on a graph far beyond what the oracle checks element-wise:
    A(aX + bW) == a AX + b AW, and row sums of A X for X = 1 equal the row sums of A."""
    from recad_b200 import ops
    U, I, E = 200_000, 40_000, 4_000_000
    gen = torch.Generator(device=_dev()).manual_seed(0)
    u = torch.randint(0, U, (E,), device=_dev(), generator=gen)
    i = (torch.rand(E, device=_dev(), generator=gen) ** 3 * I).long().clamp_(max=I - 1)     # skewed popularity
    g = ops.Graph.from_edges(u, i, U, I)
    N, D = U + I, 64
    X = torch.randn(N, D, device=_dev(), generator=gen)
    W = torch.randn(N, D, device=_dev(), generator=gen)
    AX, AW, AM = (torch.empty_like(X) for _ in range(3))
    ops.spmm(g, X, AX)
    ops.spmm(g, W, AW)
    ops.spmm(g, 2.0 * X - 3.0 * W, AM)
    assert torch.allclose(AM, 2.0 * AX - 3.0 * AW, rtol=1e-4, atol=1e-5)
    ones = torch.ones(N, D, device=_dev())
    ops.spmm(g, ones, AX)
    rows = torch.repeat_interleave(torch.arange(N, device=_dev()), g.rowptr[1:] - g.rowptr[:-1])
    rs = torch.zeros(N, device=_dev(), dtype=torch.float64).index_add_(0, rows, g.vals.double())
    assert torch.allclose(AX[:, 0].double(), rs, rtol=1e-5, atol=1e-6)
    # structure: sorted (row, col), symmetric nnz, no duplicates
    key = rows * N + g.colidx.long()
    assert bool((key[1:] > key[:-1]).all())
    assert g.nnz % 2 == 0


def test_adjacency_cache_file_is_the_reference_format(tmp_path):
    """if_cache=True: the normalised adjacency is written as `adj_mat_<name>_<U>_<I>.npz` (scipy CSR, implicit.py:279-289)
    and a second dataset is built from that file; its value bits equal the reference's matrix (golden), so either
    implementation can read what the other cached."""
    import scipy.sparse as sp
    from recad_b200 import dataset
    DEV = "cuda:0"
    tr, va, te = util.dicts("dev")
    kw = dict(train_dict=tr, valid_dict=va, test_dict=te, need_graph=True, device=torch.device(DEV), if_cache=True,
              cache_dir=str(tmp_path))
    a = dataset.from_config("implicit", "dev", **kw)
    f = tmp_path / f"adj_mat_dev_{a.n_users}_{a.n_items}.npz"
    assert f.exists()
    m = sp.load_npz(str(f))
    z = util.load("dev_graph.npz")
    assert m.dtype == np.float32 and np.array_equal(m.indptr, z["crow"]) and np.array_equal(m.indices, z["col"])
    assert m.data.tobytes() == z["val"].astype(np.float32).tobytes()       # the reference-actual graph (graph_edges="reference")
    b = dataset.from_config("implicit", "dev", **kw)              # read back from the cache file
    pa, ca, va_ = a.Graph.to_numpy()
    pb, cb, vb = b.Graph.to_numpy()
    assert np.array_equal(pa, pb) and np.array_equal(ca, cb) and va_.tobytes() == vb.tobytes()
    X = torch.randn(a.n_users + a.n_items, 64, device=DEV)
    from recad_b200 import ops
    ya, _ = ops.spmm(a.Graph, X, torch.empty_like(X))
    yb, _ = ops.spmm(b.Graph, X, torch.empty_like(X))
    assert torch.equal(ya, yb)
    c = b.inject_data("explicit", np.full((3, a.n_items), 5.0), filter_num=4)      # a cached graph is rebuilt, not extended
    assert c.n_users == a.n_users + 3 and c.Graph.n_rows == a.Graph.n_rows + 3
