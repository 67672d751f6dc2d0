"""Oracle: the BPR (pairwise) and BCE (pointwise) training samplers.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Two restatements of each sampler:
  * ``*_numpy``: the reference's own sequence of legacy ``np.random`` calls
    (the arithmetic lives in numpy's RandomState, a third-party dependency that
    is present both here and on the GPU box: numpy 2.3);
  * ``*_stream``: the same result derived from the RAW MT19937 32-bit output
    stream with numpy's published masked-rejection rule, i.e. the recipe the
    C++ product sampler (recad_b200/csrc/sampler.cpp) implements.

Reference call sites:
  * ``pairwise_sample``   recad/dataset/implicit.py:50-74
  * ``pointwise_sample``  recad/dataset/implicit.py:77-91
  * ``shuffle``           recad/dataset/implicit.py:18-35
  * ``minibatch``         recad/dataset/implicit.py:38-47
  * seeding               recad/__init__.py:11-14 (np.random.seed(2023))
numpy algorithm restated (numpy 2.3, numpy/random/_mt19937.pyx `_legacy_seeding`,
src/distributions/distributions.c `buffered_bounded_masked_uint32`,
`random_interval`; mtrand.pyx RandomState.randint/choice/shuffle).
"""
import numpy as np


# --------------------------------------------------------------------------- #
# (1) restatement with the legacy numpy calls themselves
# --------------------------------------------------------------------------- #
def pairwise_sample_numpy(n_users, n_items, train_size, allpos_indptr, allpos_indices):
    """implicit.py:50-74.  Users drawn first as ONE vector call; users without a
    positive are dropped; negative = rejection until not in the user's positives."""
    users = np.random.randint(0, n_users, train_size)
    out = []
    for user in users:
        pos = allpos_indices[allpos_indptr[user]:allpos_indptr[user + 1]]
        if len(pos) == 0:
            continue
        positem = pos[np.random.randint(0, len(pos))]
        while True:
            negitem = np.random.randint(0, n_items)
            if negitem in pos:
                continue
            break
        out.append([user, positem, negitem])
    return np.array(out, dtype=np.int64).reshape(-1, 3)


def pointwise_sample_numpy(train_dict, n_items, negative_ratio):
    """implicit.py:77-91.  Per user in dict order: positives (label 1), then
    negative_ratio * |pos| negatives drawn WITH replacement from the complement
    (iterated ascending), label 0."""
    data = []
    full = set(range(n_items))
    for uid, iids in train_dict.items():
        data.extend([(uid, iid, 1) for iid in iids])
        left = list(full - set(iids))
        assert all(left[k] < left[k + 1] for k in range(len(left) - 1)), \
            "complement iteration is expected ascending (SURVEY.md a-S)"
        negs = np.random.choice(left, size=len(iids) * negative_ratio)
        data.extend([(uid, ni, 0) for ni in negs])
    return np.array(data, dtype=np.int64).reshape(-1, 3)


def shuffle_indices_numpy(n):
    """implicit.py:24-25: np.random.shuffle(np.arange(n))."""
    idx = np.arange(n)
    np.random.shuffle(idx)
    return idx


def minibatch_slices(n, batch_size):
    """implicit.py:38-47: [0:B], [B:2B], ... the last batch is ragged."""
    return [(s, min(s + batch_size, n)) for s in range(0, n, batch_size)]


# --------------------------------------------------------------------------- #
# (2) restatement from the raw MT19937 stream
# --------------------------------------------------------------------------- #
class MT19937:
    """Raw MT19937 with numpy's legacy state layout (key[624], pos)."""

    N, M = 624, 397

    def __init__(self, key, pos):
        self.key = np.array(key, dtype=np.uint32).copy()
        self.pos = int(pos)

    @classmethod
    def from_seed(cls, seed):
        """_legacy_seeding with an int: Knuth's init_genrand; pos = 624."""
        key = np.zeros(cls.N, dtype=np.uint64)
        key[0] = seed & 0xFFFFFFFF
        for i in range(1, cls.N):
            key[i] = (1812433253 * (int(key[i - 1]) ^ (int(key[i - 1]) >> 30)) + i) & 0xFFFFFFFF
        return cls(key.astype(np.uint32), cls.N)

    @classmethod
    def from_numpy_global(cls):
        st = np.random.get_state()
        assert st[0] == "MT19937"
        return cls(st[1], st[2])

    def to_numpy_global(self):
        st = np.random.get_state()
        np.random.set_state(("MT19937", self.key.copy(), self.pos, st[3], st[4]))

    def _twist(self):
        k = self.key.astype(np.uint64)
        N, M = self.N, self.M
        for i in range(N):
            y = (int(k[i]) & 0x80000000) | (int(k[(i + 1) % N]) & 0x7FFFFFFF)
            v = int(k[(i + M) % N]) ^ (y >> 1)
            if y & 1:
                v ^= 0x9908B0DF
            k[i] = v
        self.key = k.astype(np.uint32)
        self.pos = 0

    def next_uint32(self):
        if self.pos >= self.N:
            self._twist()
        y = int(self.key[self.pos])
        self.pos += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def masked(self, r):
        """Uniform integer in [0, r]: r == 0 consumes NOTHING; otherwise draw
        32-bit words, AND with the smallest all-ones mask >= r, reject while > r."""
        if r == 0:
            return 0
        mask = r
        for s in (1, 2, 4, 8, 16):
            mask |= mask >> s
        while True:
            v = self.next_uint32() & mask
            if v <= r:
                return v


def pairwise_sample_stream(mt, n_users, n_items, train_size, allpos_indptr, allpos_indices):
    users = [mt.masked(n_users - 1) for _ in range(train_size)]
    out = []
    for user in users:
        lo, hi = int(allpos_indptr[user]), int(allpos_indptr[user + 1])
        if hi == lo:
            continue
        pos = allpos_indices[lo:hi]
        positem = int(pos[mt.masked(hi - lo - 1)])
        posset = set(pos.tolist())
        while True:
            neg = mt.masked(n_items - 1)
            if neg not in posset:
                break
        out.append([user, positem, neg])
    return np.array(out, dtype=np.int64).reshape(-1, 3)


def pointwise_sample_stream(mt, user_ids, pos_indptr, pos_items, n_items, negative_ratio):
    """user_ids / pos_indptr / pos_items: train_dict in dict order, item lists in
    their stored order (NOT sorted)."""
    out = []
    for k, uid in enumerate(user_ids):
        iids = pos_items[pos_indptr[k]:pos_indptr[k + 1]].tolist()
        out.extend([(uid, i, 1) for i in iids])
        left = np.setdiff1d(np.arange(n_items), np.asarray(iids, dtype=np.int64))
        for _ in range(len(iids) * negative_ratio):
            out.append((uid, int(left[mt.masked(len(left) - 1)]), 0))
    return np.array(out, dtype=np.int64).reshape(-1, 3)


def shuffle_indices_stream(mt, n):
    """RandomState.shuffle on a 1-d array: Fisher-Yates from the top,
    j = random_interval(i) (masked rejection), swap(i, j)."""
    idx = list(range(n))
    for i in range(n - 1, 0, -1):
        j = mt.masked(i)
        idx[i], idx[j] = idx[j], idx[i]
    return np.array(idx, dtype=np.int64)
