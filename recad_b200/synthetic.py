"""Shape-matched synthetic interaction data (ml1m.zip / yelp.zip are absent from
the reference checkout and there is no network: SURVEY.md section 8c/8d).

All generators are deterministic functions of ``seed`` (numpy Generator PCG64).
The host generators return dicts in the reference's ``{uid: [iid, ...]}`` form
(recad/dataset/implicit.py:94-104); the large generator returns flat arrays.
"""
import numpy as np

ML1M = dict(n_users=5950, n_items=3702, train=468_649, valid=49_390, test=49_494)   # readme.md:168-176
YELP = dict(n_users=54_632, n_items=34_474, train=1_500_000, valid=150_000, test=150_000)  # data/readme.md:62; count chosen
SYNTH = dict(n_users=1_000_000, n_items=200_000, train=50_000_000)                  # BASELINE.json configs[3]


def _draw_unique_pairs(rng, n_users, n_items, total, zipf=0.8, sigma=1.0):
    """`total` distinct (u, i) pairs; user activity ~ lognormal, item popularity
    ~ rank^-zipf; every user gets at least one pair."""
    uw = rng.lognormal(0.0, sigma, n_users)
    uw /= uw.sum()
    iw = np.arange(1, n_items + 1, dtype=np.float64) ** (-zipf)
    iw /= iw.sum()
    perm = rng.permutation(n_items)          # popularity not correlated with the id
    ucdf, icdf = np.cumsum(uw), np.cumsum(iw)
    keys = np.arange(n_users, dtype=np.int64) * n_items + perm[
        np.minimum(np.searchsorted(icdf, rng.random(n_users)), n_items - 1)]
    keys = np.unique(keys)
    while len(keys) < total:
        need = int((total - len(keys)) * 1.25) + 1024
        u = np.minimum(np.searchsorted(ucdf, rng.random(need)), n_users - 1).astype(np.int64)
        i = perm[np.minimum(np.searchsorted(icdf, rng.random(need)), n_items - 1)].astype(np.int64)
        keys = np.unique(np.concatenate([keys, u * n_items + i]))
    if len(keys) > total:
        # drop surplus pairs, never a user's only pair
        u = keys // n_items
        first = np.ones(len(keys), dtype=bool)
        first[1:] = u[1:] != u[:-1]
        removable = np.flatnonzero(~first)
        drop = rng.choice(removable, size=len(keys) - total, replace=False)
        keep = np.ones(len(keys), dtype=bool)
        keep[drop] = False
        keys = keys[keep]
    return keys // n_items, keys % n_items


def _to_dict(users, items):
    order = np.argsort(users, kind="stable")
    users, items = users[order], items[order]
    cuts = np.flatnonzero(np.diff(users)) + 1
    starts = np.concatenate([[0], cuts])
    ends = np.concatenate([cuts, [len(users)]])
    return {int(users[s]): items[s:e].tolist() for s, e in zip(starts, ends)}


def make_splits(shape=ML1M, seed=0, zipf=0.8):
    """train/valid/test dicts with the interaction counts of ``shape``; the three
    splits are disjoint and every user has >= 1 train item."""
    rng = np.random.default_rng(seed)
    U, I = shape["n_users"], shape["n_items"]
    total = shape["train"] + shape["valid"] + shape["test"]
    u, i = _draw_unique_pairs(rng, U, I, total, zipf)
    first = np.ones(len(u), dtype=bool)
    first[1:] = u[1:] != u[:-1]          # pairs are sorted by (u, i): one guaranteed train pair per user
    rest = rng.permutation(np.flatnonzero(~first))
    n_tr_extra = shape["train"] - int(first.sum())
    tr = np.concatenate([np.flatnonzero(first), rest[:n_tr_extra]])
    va = rest[n_tr_extra:n_tr_extra + shape["valid"]]
    te = rest[n_tr_extra + shape["valid"]:]
    # time order inside a user's list is arbitrary in the reference (timestamp sort): shuffle it
    out = []
    for sel in (tr, va, te):
        sel = rng.permutation(sel)
        out.append(_to_dict(u[sel], i[sel]))
    return tuple(out)


def make_edges(n_users, n_items, n_edges, seed=0, zipf=0.8):
    """Flat distinct (u, i) int64 arrays, sorted by (u, i) (host; for mid sizes)."""
    rng = np.random.default_rng(seed)
    return _draw_unique_pairs(rng, n_users, n_items, n_edges, zipf)
