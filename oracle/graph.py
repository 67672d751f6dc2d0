"""Oracle: dataset flattening and the symmetric-normalised adjacency.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, for arbitrary sizes and without the dok/lil Python-speed detours:
  * ``ImplicitData.read_data``            recad/dataset/implicit.py:221-241
  * ``ImplicitData._init_data``           recad/dataset/implicit.py:166-219
  * ``ImplicitData.getSparseGraph``       recad/dataset/implicit.py:243-298
  * ``_convert_sp_mat_to_sp_tensor``      recad/dataset/implicit.py:320-326
  * ``fake_array2dict`` / ``inject_data`` recad/dataset/implicit.py:107-114,482-494

Pinned against the live reference on dev, game and an ml1m-shaped synthetic
graph by tests/golden/make_golden.py -> tests/test_oracle_golden.py.
"""
import numpy as np


def flatten_dict(data_dict):
    """read_data (implicit.py:221-241): dict {uid: [iid..]} -> flat (user[], item[]).

    Users with an empty list are skipped.  Returns (users int64, items int64,
    max_uid, max_iid) with max_* = -1 when there is no interaction."""
    users, items = [], []
    max_u, max_i = -1, -1
    for uid, iids in data_dict.items():
        if len(iids) == 0:
            continue
        users.extend([uid] * len(iids))
        items.extend(iids)
        max_u = max(max_u, int(uid))
        max_i = max(max_i, int(max(iids)))
    return (np.asarray(users, dtype=np.int64), np.asarray(items, dtype=np.int64), max_u, max_i)


def dataset_shape(train_dict, valid_dict, test_dict):
    """_init_data (implicit.py:166-195): n_users / n_items are max id + 1 over
    all three splits (the running max starts at 0, so an empty dataset has 1/1)."""
    mu, mi = 0, 0
    for d in (train_dict, valid_dict, test_dict):
        _, _, u, i = flatten_dict(d)
        mu, mi = max(mu, u), max(mi, i)
    return mu + 1, mi + 1


def graph_edges_reference(train_dict, valid_dict, test_dict):
    """The edge set the reference ACTUALLY feeds the graph with.

    read_data overwrites self.trainUser/self.trainItem on every call
    (implicit.py:233-235) and _init_data calls it train -> valid -> test
    (implicit.py:173-192), so UserItemNet (206-209) is built from the TEST split.
    """
    u, i, _, _ = flatten_dict(test_dict)
    return u, i


def unique_edges(users, items, n_users, n_items):
    """csr_matrix((ones, (u, i))) sums duplicate pairs (implicit.py:206-209):
    return the distinct (u, i) pairs in (u, i) order plus their multiplicity."""
    key = users.astype(np.int64) * np.int64(n_items) + items.astype(np.int64)
    uk, mult = np.unique(key, return_counts=True)
    return uk // n_items, uk % n_items, mult.astype(np.int64)


def norm_adj_csr(users, items, n_users, n_items):
    """A_hat = D^-1/2 [[0, R], [R^T, 0]] D^-1/2 as CSR over N = n_users + n_items.

    implicit.py:259-276: the adjacency is float32; rowsum is float32;
    ``d_inv = np.power(rowsum + 1e-14, -0.5)`` is evaluated by numpy IN float32 on
    an (N, 1) array; ``norm = D.dot(A).dot(D)`` => value = (d[r] * a) * d[c] as two
    successive float32 products.  coalesce() (implicit.py:296) orders entries by
    (row, col).  Returns (indptr int64[N+1], indices int64[nnz], data float32[nnz],
    d_inv float32[N], degree int64[N])."""
    U, I = int(n_users), int(n_items)
    N = U + I
    eu, ei, mult = unique_edges(users, items, U, I)
    # user rows (cols offset by U), already in (row, col) order
    r_top, c_top, m_top = eu, ei + U, mult
    # item rows: transpose, sorted by (item, user)
    order = np.lexsort((eu, ei))
    r_bot, c_bot, m_bot = ei[order] + U, eu[order], mult[order]
    rows = np.concatenate([r_top, r_bot])
    cols = np.concatenate([c_top, c_bot])
    m = np.concatenate([m_top, m_bot]).astype(np.float32)
    deg = np.bincount(rows, weights=m.astype(np.float64), minlength=N)
    rowsum = deg.astype(np.float32).reshape(-1, 1)
    d_inv = np.power(rowsum + 1e-14, -0.5).flatten()
    d_inv[np.isinf(d_inv)] = 0.0  # implicit.py:272 (never fires: 1e-14 ** -0.5 = 1e7)
    data = (d_inv[rows] * m) * d_inv[cols]
    indptr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=N), out=indptr[1:])
    return indptr, cols.astype(np.int64), data.astype(np.float32), d_inv, deg.astype(np.int64)


def d_inv_from_degree(deg):
    """The exact numpy expression of implicit.py:270-272 on a degree vector.
    Same shape / dtype path as the reference: float32 (N, 1) array."""
    rowsum = np.asarray(deg).astype(np.float32).reshape(-1, 1)
    d_inv = np.power(rowsum + 1e-14, -0.5).flatten()
    d_inv[np.isinf(d_inv)] = 0.0
    return d_inv


def all_pos(users, items, n_users, n_items):
    """getUserPosItems (implicit.py:339-343): per-user sorted distinct items of
    UserItemNet, as CSR (indptr int64[U+1], indices int64)."""
    eu, ei, _ = unique_edges(users, items, n_users, n_items)
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    np.cumsum(np.bincount(eu, minlength=n_users), out=indptr[1:])
    return indptr, ei


def fake_array_to_dict(fake_array, n_users, filter_num=4):
    """fake_array2dict (implicit.py:107-114): rows -> new user ids n_users + r,
    keeping the columns whose rating is STRICTLY greater than filter_num."""
    fake_array = np.asarray(fake_array)
    assert fake_array.ndim == 2
    r, c = np.where(fake_array > filter_num)
    out = {}
    for u, i in zip((r + n_users).tolist(), c.tolist()):
        out.setdefault(u, []).append(i)
    return out


def inject(train_dict, fake_array, n_users, filter_num=4):
    """inject_data (implicit.py:482-494): new train dict = old + fake rows.
    Injection onto an existing user id is an AssertionError (implicit.py:489-491)."""
    new = dict(train_dict)
    for k, v in fake_array_to_dict(fake_array, n_users, filter_num).items():
        assert k not in new, f"Injection to a exist user {k} is not allowed"
        new[k] = v
    return new
