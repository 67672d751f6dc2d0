"""Tensor-level wrappers over the C ABI (include/recad_b200.h).

PyTorch is used here for device memory and the current CUDA stream only; every
computation is a kernel of librecad_b200.so.  All functions raise RecadError on
failure -- there is no fallback path.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import RecadError, check

SEG_LEN = None  # SpMM plan: max stored entries one warp walks (multiple of 32); None = by size (auto_seg_len)


def auto_seg_len(nnz):
    """Measured on B200 (profiles/spmm_variant_sweep_r01.jsonl, spmm_sweep_ml1m_r01.jsonl): long segments
    amortise the per-segment prologue when there are far more segments than resident warps (synthetic,
    100 M entries: 256 is fastest); an L2-resident graph with ~1 segment per warp slot is latency-bound and
    wants more, shorter chains (ml1m-shaped, 0.94 M entries: 64 is 25 % faster than 256)."""
    return 64 if nnz <= 8_000_000 else 256


def _ptr(t):
    if t is None:
        return None
    assert t.is_contiguous(), "recad_b200 ops need contiguous tensors"
    return C.c_void_p(t.data_ptr())


def _stream(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RecadError("recad_b200 runs on CUDA tensors only (no CPU path); got a CPU tensor")


def device_info():
    sm, a, b = C.c_int(), C.c_int(), C.c_int()
    check(_lib.lib().recad_device_info(C.byref(sm), C.byref(a), C.byref(b)), "recad_device_info")
    return sm.value, a.value, b.value


# --------------------------------------------------------------------------- #
# graph
# --------------------------------------------------------------------------- #
def d_inv_from_degree(degree):
    """The reference's own host expression (recad/dataset/implicit.py:269-272):
    float32 rowsum as an (N, 1) array, ``np.power(rowsum + 1e-14, -0.5)``.  It is
    evaluated on the host with numpy on purpose: numpy's float32 pow is not
    correctly rounded and no device expression reproduces its bits."""
    rowsum = np.asarray(degree).astype(np.float32).reshape(-1, 1)
    d_inv = np.power(rowsum + 1e-14, -0.5).flatten()
    d_inv[np.isinf(d_inv)] = 0.0
    return d_inv


def auto_col_blocks(n_cols, row_bytes=256):
    """Column blocks for rows that gather from a table of n_cols rows: one block while the table fits the 126 MB L2
    next to the streamed matrix, else ~80 MB blocks (measured on B200: 2-4 blocks of the 256 MB user table are equally
    fast, profiles/spmm2_sweep_r02.jsonl)."""
    nbytes = int(n_cols) * row_bytes
    return 1 if nbytes <= 100e6 else max(2, int(round(nbytes / 80e6)))


class PackedPlan:
    """Work list of the SpMM (csrc/spmm.cu): one 16-byte entry per SEGMENT {lo (64 bit), row, (slot + 1) << 11 | count}.
    `groups` = [(row_lo, row_hi, col_lo, col_hi, n_blocks)] partitions the rows; the rows of a group gather columns in
    [col_lo, col_hi) and are cut at n_blocks equal column blocks, their segments ordered block by block, so that the
    gathered slice of X (an L2-sized block of the user table) stays cache resident while it is in use; inside a
    (group, block) the segments go longest first (no straggler warps at the tail of the persistent kernel).  A row
    with more than one segment gets consecutive partial-sum slots in (row, first entry) order; its last-arriving warp
    adds them in that order (deterministic).
    Plain index arithmetic on torch tensors (runs wherever the row pointer lives): executed once per graph, the hot
    path only reads its output."""

    MAX_SEG = 1024

    def __init__(self, meta, row_mseg, n_mrow, n_slot, group_segs):
        self.meta, self.row_mseg, self.n_mrow, self.n_slot, self.group_segs = meta, row_mseg, int(n_mrow), int(n_slot), group_segs
        self.n_seg = int(meta.shape[0])

    @classmethod
    def build(cls, rowptr, colidx, n_rows, seg_len, groups=None):
        if not (0 < seg_len <= cls.MAX_SEG):
            raise RecadError(f"PackedPlan: seg_len must be in (0, {cls.MAX_SEG}]")
        dev = rowptr.device
        i64 = dict(dtype=torch.int64, device=dev)
        rowptr = rowptr.long()
        if groups is None:
            groups = [(0, n_rows, 0, 1 << 31, 1)]
        assert groups[0][0] == 0 and groups[-1][1] == n_rows and all(a[1] == b[0] for a, b in zip(groups, groups[1:]))
        # pieces = (row, lo, hi, order key): one piece per (row, column block)
        p_row, p_lo, p_hi, p_grp = [], [], [], []
        gkey = 0
        for (r0, r1, c0, c1, nb) in groups:
            rows = torch.arange(r0, r1, **i64)
            if nb <= 1 or r1 == r0:
                p_row.append(rows); p_lo.append(rowptr[r0:r1]); p_hi.append(rowptr[r0 + 1:r1 + 1])
                p_grp.append(torch.full((r1 - r0,), gkey, **i64))
                gkey += 1
                continue
            e0, e1 = int(rowptr[r0]), int(rowptr[r1])
            erow = torch.repeat_interleave(rows, rowptr[r0 + 1:r1 + 1] - rowptr[r0:r1])
            key = (erow << 32) | colidx[e0:e1].long()                       # ascending: CSR is (row, col) sorted
            bounds = torch.tensor([c0 + (b * (c1 - c0)) // nb for b in range(1, nb)], **i64)
            q = (rows[:, None] << 32) | bounds[None, :]
            cut = torch.searchsorted(key, q.reshape(-1)).reshape(-1, nb - 1) + e0
            edges = torch.cat([rowptr[r0:r1, None], cut, rowptr[r0 + 1:r1 + 1, None]], 1)   # [rows, nb + 1]
            lo_b, hi_b = edges[:, :-1], edges[:, 1:]
            keep = hi_b > lo_b
            keep[:, 0] |= ~keep.any(1)                                       # an empty row is still written once
            grp_b = (gkey + torch.arange(nb, **i64))[None, :].expand_as(lo_b)
            p_row.append(rows[:, None].expand_as(lo_b)[keep]); p_lo.append(lo_b[keep]); p_hi.append(hi_b[keep])
            p_grp.append(grp_b[keep])
            gkey += nb
        p_row, p_lo, p_hi, p_grp = (torch.cat(x) for x in (p_row, p_lo, p_hi, p_grp))
        # pieces -> segments of at most seg_len entries
        nch = torch.clamp((p_hi - p_lo + seg_len - 1) // seg_len, min=1)
        first = torch.cumsum(nch, 0) - nch
        n_seg = int(nch.sum())
        piece = torch.repeat_interleave(torch.arange(p_row.numel(), **i64), nch)
        k = torch.arange(n_seg, **i64) - first[piece]
        s_row, s_grp = p_row[piece], p_grp[piece]
        s_lo = p_lo[piece] + k * seg_len
        s_cnt = torch.minimum(p_hi[piece] - s_lo, torch.full_like(s_lo, seg_len))
        # partial-sum slots of the rows with more than one segment, in (row, lo) order
        per_row = torch.bincount(s_row, minlength=n_rows)
        multi = per_row > 1
        slots_of = torch.where(multi, per_row, torch.zeros_like(per_row))
        slot0 = torch.cumsum(slots_of, 0) - slots_of
        n_slot, n_mrow = int(slots_of.sum()), int(multi.sum())
        if n_slot + 1 >= (1 << 21):
            raise RecadError(f"PackedPlan: {n_slot} partial slots exceed the packed field; use a longer seg_len")
        by_row = torch.argsort(s_row * (1 << 40) + s_lo)                    # lo < 2^40
        row_first = torch.cumsum(per_row, 0) - per_row
        rank = torch.empty(n_seg, **i64)
        rank[by_row] = torch.arange(n_seg, **i64) - row_first[s_row[by_row]]
        s_slot = torch.where(multi[s_row], slot0[s_row] + rank, torch.full_like(rank, -1))
        # execution order: group by group, block by block, longest first
        ex = torch.argsort(s_grp * (1 << 44) + (cls.MAX_SEG - s_cnt) * (1 << 32) + s_row, stable=True)
        meta = torch.stack([s_lo & 0xffffffff, s_lo >> 32, s_row, ((s_slot + 1) << 11) | s_cnt], 1)[ex]
        meta = torch.where(meta >= (1 << 31), meta - (1 << 32), meta).to(torch.int32).contiguous()
        row_mseg = torch.stack([slot0, per_row], 1).to(torch.int32).contiguous()
        # segment ranges of the caller's groups (profiling tools time them separately)
        g_of_key, gk = [], 0
        for gi, (r0, r1, c0, c1, nb) in enumerate(groups):
            n_keys = 1 if (nb <= 1 or r1 == r0) else nb
            g_of_key += [gi] * n_keys
            gk += n_keys
        cnt_key = torch.bincount(s_grp, minlength=gk).tolist()
        group_segs, at = [], 0
        for gi in range(len(groups)):
            n = sum(c for c, g in zip(cnt_key, g_of_key) if g == gi)
            group_segs.append((at, at + n))
            at += n
        return cls(meta, row_mseg, n_mrow, n_slot, group_segs)


class Graph:
    """Device-resident CSR of a (normalised) adjacency + its SpMM work plan."""

    def __init__(self, n_rows, n_cols, rowptr, colidx, vals, mult=None, degree=None, seg_len=SEG_LEN, split=None):
        """split = n_users for the symmetric bipartite adjacency (rows < split gather item rows of X, rows >= split
        gather user rows): the two halves become separate plan groups, each column-blocked when its gathered table
        exceeds L2 (PackedPlan)."""
        self.n_rows, self.n_cols = int(n_rows), int(n_cols)
        self.rowptr, self.colidx, self.vals, self.mult, self.degree = rowptr, colidx, vals, mult, degree
        self.nnz = int(colidx.numel())
        self.device = rowptr.device
        self.seg_len = seg_len or auto_seg_len(self.nnz)
        self.split = None if split is None else int(split)
        self._partials = {}
        self._structs = {}
        self._plan()

    # -- construction ------------------------------------------------------ #
    @classmethod
    def from_edges(cls, users, items, n_users, n_items, seg_len=SEG_LEN, degree_hook=None):
        """Symmetric-normalised bipartite adjacency (implicit.py:243-298) from an
        edge list on the device.  users / items: int64 CUDA tensors.
        degree_hook(degree int32 [N]) may complete the degrees in place before normalisation (the sharded
        path all-reduces the item block: a rank's rows hold only its own users' edges)."""
        _need_cuda(users, items)
        L = _lib.lib()
        dev = users.device
        users, items = users.contiguous().long(), items.contiguous().long()
        E, N = int(users.numel()), int(n_users) + int(n_items)
        with torch.cuda.device(dev):
            rowptr = torch.empty(N + 1, dtype=torch.int64, device=dev)
            colidx = torch.empty(max(2 * E, 1), dtype=torch.int32, device=dev)
            mult = torch.empty(max(2 * E, 1), dtype=torch.float32, device=dev)
            degree = torch.empty(N, dtype=torch.int32, device=dev)
            nbytes = L.recad_csr_build_scratch_bytes(E, n_users, n_items)
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            nnz = C.c_int64()
            check(L.recad_csr_build_structure(_ptr(users), _ptr(items), E, n_users, n_items, _ptr(rowptr), _ptr(colidx),
                                              _ptr(mult), _ptr(degree), C.byref(nnz), _ptr(scratch), nbytes, _stream(dev)),
                  "recad_csr_build_structure")
            del scratch
            colidx, mult = colidx[:nnz.value].clone(), mult[:nnz.value].clone()
            if degree_hook is not None:
                degree_hook(degree)
            vals = cls._normalize(rowptr, colidx, mult, degree, N)
        return cls(N, N, rowptr, colidx, vals, mult, degree, seg_len, split=n_users)

    @staticmethod
    def _normalize(rowptr, colidx, mult, degree, N):
        dev = rowptr.device
        d_inv = torch.from_numpy(d_inv_from_degree(degree.cpu().numpy())).to(dev)
        vals = torch.empty(colidx.numel(), dtype=torch.float32, device=dev)
        check(_lib.lib().recad_csr_normalize(_ptr(rowptr), _ptr(colidx), _ptr(mult), _ptr(d_inv), N, _ptr(vals), _stream(dev)),
              "recad_csr_normalize")
        return vals

    def append_users(self, n_users, n_items, fake_rowptr, fake_items, degree_hook=None):
        """In-place injection (implicit.py:482-494): returns the graph over
        n_users + F users with the fake rows appended; no sort, no Python dict.
        degree_hook: as in from_edges (a shard completes its item degrees before normalisation)."""
        L = _lib.lib()
        dev = self.device
        F = int(fake_rowptr.numel()) - 1
        nf = int(fake_items.numel())
        Nn = n_users + F + n_items
        with torch.cuda.device(dev):
            fake_rowptr = fake_rowptr.to(dev, torch.int64).contiguous()
            fake_items = fake_items.to(dev, torch.int32).contiguous()
            rowptr = torch.empty(Nn + 1, dtype=torch.int64, device=dev)
            colidx = torch.empty(max(self.nnz + 2 * nf, 1), dtype=torch.int32, device=dev)
            mult = torch.empty(max(self.nnz + 2 * nf, 1), dtype=torch.float32, device=dev)
            degree = torch.empty(Nn, dtype=torch.int32, device=dev)
            nbytes = L.recad_csr_append_scratch_bytes(n_users, n_items, F, nf)
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            check(L.recad_csr_append_users(_ptr(self.rowptr), _ptr(self.colidx), _ptr(self.mult), n_users, n_items, F,
                                           _ptr(fake_rowptr), _ptr(fake_items), nf, _ptr(rowptr), _ptr(colidx), _ptr(mult),
                                           _ptr(degree), _ptr(scratch), nbytes, _stream(dev)), "recad_csr_append_users")
            colidx, mult = colidx[:self.nnz + 2 * nf], mult[:self.nnz + 2 * nf]
            if degree_hook is not None:
                degree_hook(degree)
            vals = self._normalize(rowptr, colidx, mult, degree, Nn)
        return Graph(Nn, Nn, rowptr, colidx, vals, mult, degree, self.seg_len, split=n_users + F)

    def renormalized(self, degree):
        """The same structure with values recomputed from new degrees (a shard whose own rows did not
        change while fake users elsewhere changed the item degrees)."""
        with torch.cuda.device(self.device):
            vals = self._normalize(self.rowptr, self.colidx, self.mult, degree, self.n_rows)
        return Graph(self.n_rows, self.n_cols, self.rowptr, self.colidx, vals, self.mult, degree, self.seg_len, split=self.split)

    @classmethod
    def from_csr(cls, rowptr, colidx, vals, n_cols, seg_len=SEG_LEN, split=None):
        """Wrap an existing (possibly rectangular) CSR, e.g. a user-row shard or a cached adjacency (split = n_users for
        the symmetric bipartite matrix).  Without multiplicities / degrees such a graph cannot be extended in place."""
        _need_cuda(rowptr, colidx, vals)
        return cls(rowptr.numel() - 1, n_cols, rowptr.contiguous().long(), colidx.contiguous().int(),
                   vals.contiguous().float(), None, None, seg_len, split=split)

    # -- plan ---------------------------------------------------------------- #
    def _plan(self):
        if self.split is not None and 0 < self.split < self.n_rows:
            U = self.split
            groups = [(0, U, U, self.n_cols, auto_col_blocks(self.n_cols - U)), (U, self.n_rows, 0, U, auto_col_blocks(U))]
        else:
            groups = [(0, self.n_rows, 0, self.n_cols, auto_col_blocks(self.n_cols))]
        self.plan = PackedPlan.build(self.rowptr, self.colidx, self.n_rows, self.seg_len, groups)
        self.n_seg, self.n_mrow, self.n_slot = self.plan.n_seg, self.plan.n_mrow, self.plan.n_slot
        with torch.cuda.device(self.device):
            self.cv = torch.empty((max(self.nnz, 1), 2), dtype=torch.int32, device=self.device)
            check(_lib.lib().recad_spmm_pack_cv(_ptr(self.colidx), _ptr(self.vals), self.nnz, _ptr(self.cv), _stream(self.device)),
                  "recad_spmm_pack_cv")
            self.row_cnt = torch.zeros(self.n_rows, dtype=torch.int32, device=self.device)

    def struct(self, D):
        """ctypes recad_csr for embedding width D.  The partial-sum scratch is per D and is never freed while a
        struct handed out for that D may still be referenced by a victim."""
        if D not in self._structs:
            self._partials[D] = torch.empty(max(self.n_slot, 1) * D, dtype=torch.float32, device=self.device)
            s = _lib.CSR()
            s.n_rows, s.n_cols, s.nnz = self.n_rows, self.n_cols, self.nnz
            s.rowptr, s.colidx, s.vals, s.cv = self.rowptr.data_ptr(), self.colidx.data_ptr(), self.vals.data_ptr(), self.cv.data_ptr()
            s.n_seg, s.seg_meta = self.n_seg, self.plan.meta.data_ptr()
            s.n_mrow, s.row_mseg, s.row_cnt = self.n_mrow, self.plan.row_mseg.data_ptr(), self.row_cnt.data_ptr()
            s.partials = self._partials[D].data_ptr()
            self._structs[D] = s
        return self._structs[D]

    def with_values(self, vals=None):
        """A matrix with this one's structure and PLAN but values of its own (graph dropout: the kept entries rescaled, the
        dropped ones zero).  Only vals / cv and the partial-sum scratch are new; call `set_values` to refill them in place
        (the structs handed out stay valid).  Shares row_cnt with its parent: never launch both at the same time."""
        g = object.__new__(Graph)
        g.__dict__.update(self.__dict__)
        g._partials, g._structs = {}, {}
        with torch.cuda.device(self.device):
            g.vals = torch.empty_like(self.vals)
            g.cv = torch.empty_like(self.cv)
        g.set_values(self.vals if vals is None else vals)
        return g

    def transpose_permutation(self):
        """T with entry k = (r, c) -> the position of (c, r): vals[T] are the values of the transposed matrix on the SAME
        structure (square, structurally symmetric matrices only -- the bipartite adjacency).  Cached."""
        if getattr(self, "_tperm", None) is None:
            if self.n_rows != self.n_cols:
                raise RecadError("transpose_permutation: the matrix is not square")
            with torch.cuda.device(self.device):
                rows = torch.repeat_interleave(torch.arange(self.n_rows, device=self.device), self.rowptr[1:] - self.rowptr[:-1])
                cols = self.colidx.long()
                keys = rows * self.n_cols + cols                              # ascending: CSR order with sorted columns
                T = torch.searchsorted(keys, cols * self.n_cols + rows)
                if self.nnz and not bool((keys[T.clamp_(max=self.nnz - 1)] == cols * self.n_cols + rows).all()):
                    raise RecadError("transpose_permutation: the structure is not symmetric")
            self._tperm = T
        return self._tperm

    def set_values(self, vals):
        with torch.cuda.device(self.device):
            self.vals.copy_(vals)
            check(_lib.lib().recad_spmm_pack_cv(_ptr(self.colidx), _ptr(self.vals), self.nnz, _ptr(self.cv), _stream(self.device)),
                  "recad_spmm_pack_cv")

    # -- export --------------------------------------------------------------- #
    def to_numpy(self):
        return (self.rowptr.cpu().numpy(), self.colidx.cpu().numpy().astype(np.int64), self.vals.cpu().numpy())

    def to_torch_sparse(self):
        """torch sparse COO view with the reference's layout (implicit.py:295-296)."""
        rows = torch.repeat_interleave(torch.arange(self.n_rows, device=self.device), self.rowptr[1:] - self.rowptr[:-1])
        return torch.sparse_coo_tensor(torch.stack([rows, self.colidx.long()]), self.vals, (self.n_rows, self.n_cols)).coalesce()

    def algorithmic_bytes(self, D):
        """SURVEY.md section 8(d) no-reuse gather model for one SpMM over this matrix."""
        return self.nnz * (8 + 4 * D) + self.n_rows * 4 * D + (self.n_rows + 1) * 4


def spmm(graph, X, Y=None, C_=None, Z=None, alpha=1.0):
    """Y = A X; Z = alpha * (C + A X).  Returns (Y, Z)."""
    _need_cuda(X, Y, C_, Z)
    D = X.shape[1]
    with torch.cuda.device(X.device):
        check(_lib.lib().recad_spmm(C.byref(graph.struct(D)), _ptr(X), _ptr(Y), _ptr(C_), _ptr(Z), float(alpha), D,
                                    _stream(X.device)), "recad_spmm")
    return Y, Z


def axpby(z, a, x, b, y):
    """z = a * x + b * y (z may alias x or y)."""
    _need_cuda(z, x, y)
    with torch.cuda.device(z.device):
        check(_lib.lib().recad_axpby(_ptr(z), float(a), _ptr(x), float(b), _ptr(y), z.numel(), _stream(z.device)), "recad_axpby")
    return z


def bpr_fwd_bwd(O, E, n_users, n_items, samples, perm, grad_scale, gO, cnt, loss_acc, B_norm=None):
    """samples: int64 [n, 3] (user, pos, neg); perm: int64 [B] rows of this batch or None (= all rows in order).
    B_norm: batch size used for the 1/B normalisation (default: the number of rows processed)."""
    _need_cuda(O, E, samples, perm, gO, cnt, loss_acc)
    B = perm.numel() if perm is not None else samples.shape[0]
    with torch.cuda.device(O.device):
        check(_lib.lib().recad_bpr_fwd_bwd(_ptr(O), _ptr(E), n_users, n_items, _ptr(samples), _ptr(perm), B,
                                           int(B_norm or B), float(grad_scale), _ptr(gO), _ptr(cnt), _ptr(loss_acc), O.shape[1], _stream(O.device)),
              "recad_bpr_fwd_bwd")


def adam(p, g, m, v, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, cnt=None, reg_scale=0.0):
    _need_cuda(p, g, m, v, cnt)
    D = p.shape[1] if p.dim() == 2 else 1
    with torch.cuda.device(p.device):
        check(_lib.lib().recad_adam(_ptr(p), _ptr(g), _ptr(cnt), float(reg_scale), _ptr(m), _ptr(v), p.numel(), D, lr, b1, b2, eps,
                                    int(step), _stream(p.device)), "recad_adam")


def gemm_tn(A, B, bias=None, relu=False):
    """C = A @ B.T (+ bias) (relu) on the tensor cores (tcgen05, 3xTF32 = fp32-accurate).  A [M, K], B [N, K]."""
    _need_cuda(A, B, bias)
    M, K = A.shape
    N = B.shape[0]
    ld = (K + 3) // 4 * 4
    out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    with torch.cuda.device(A.device):
        scratch = torch.empty(2 * (M + N) * ld, dtype=torch.float32, device=A.device)
        check(_lib.lib().recad_gemm_tn_tf32x3(_ptr(A.contiguous()), _ptr(B.contiguous()), M, N, K, _ptr(bias), int(relu), _ptr(out),
                                              _ptr(scratch), _stream(A.device)), "recad_gemm_tn_tf32x3")
    return out


def dot_scores(O, n_users, users, items):
    _need_cuda(O, users, items)
    out = torch.empty(users.numel(), dtype=torch.float32, device=O.device)
    with torch.cuda.device(O.device):
        check(_lib.lib().recad_dot_scores(_ptr(O), n_users, _ptr(users), _ptr(items), users.numel(), O.shape[1], _ptr(out),
                                          _stream(O.device)), "recad_dot_scores")
    return out


# --------------------------------------------------------------------------- #
# evaluation
# --------------------------------------------------------------------------- #
def transpose_items(item_emb):
    """[I, D] -> [D, ld] with ld = I rounded up to 64, zero padded."""
    _need_cuda(item_emb)
    I, D = item_emb.shape
    ld = (I + 63) // 64 * 64
    out = torch.zeros((D, ld), dtype=torch.float32, device=item_emb.device)
    with torch.cuda.device(item_emb.device):
        check(_lib.lib().recad_transpose_items(_ptr(item_emb.contiguous()), I, D, _ptr(out), ld, _stream(item_emb.device)),
              "recad_transpose_items")
    return out


def _eval_outputs(n, T, K, dev):
    return (torch.empty((n, K), dtype=torch.int32, device=dev), torch.empty((n, K), dtype=torch.float32, device=dev),
            torch.empty((n, max(T, 1)), dtype=torch.int32, device=dev), torch.empty((n, max(T, 1)), dtype=torch.float32, device=dev))


EVAL_PRECISION = os.environ.get("RECAD_EVAL_PRECISION", "tf32x3")   # "tf32x3" (tensor cores) | "exact" (fp32 FMA chain)


def fullrank_eval(user_emb, item_emb, user_ids, train_rowptr, train_col, targets, K, item_T=None, item_bias=None,
                  precision=None):
    """Fused score + mask + top-K + target rank (normal.py:57-93 without the score matrix).
    score(u, i) = <user_emb[u], item_emb[i]> (+ item_bias[i]).  Returns (topk_idx, topk_val, target_rank,
    target_score).  precision "tf32x3": tcgen05 tensor-core kernel (needs D <= 64, K <= 32; fp32-accurate scores,
    ranks may differ from "exact" between near-tied items only); "exact": CUDA-core kernel whose scores are the
    bit-exact ascending-d FMA chain."""
    _need_cuda(user_emb, item_emb, user_ids, train_rowptr, train_col, item_bias)
    dev = user_emb.device
    I, D = item_emb.shape
    targets_t = torch.as_tensor(list(targets), dtype=torch.int32, device=dev)
    T = int(targets_t.numel())
    if T and (int(targets_t.min()) < 0 or int(targets_t.max()) >= I):
        raise RecadError(f"target item id out of range [0, {I})")
    n = int(user_ids.numel())
    topi, topv, trank, tscore = _eval_outputs(n, T, K, dev)
    precision = precision or EVAL_PRECISION
    if precision not in ("tf32x3", "exact"):
        raise RecadError(f"unknown evaluation precision {precision!r}")
    if precision == "tf32x3" and D <= 64 and K <= 32:
        L = _lib.lib()
        nfl = L.recad_fullrank_tc_scratch_floats(n, I)
        with torch.cuda.device(dev):
            scratch = torch.empty(nfl, dtype=torch.float32, device=dev)
            check(L.recad_fullrank_eval_tc(_ptr(user_emb.contiguous()), _ptr(item_emb.contiguous()), I, D,
                                           _ptr(user_ids.contiguous().long()), n, _ptr(train_rowptr), _ptr(train_col),
                                           _ptr(targets_t), T, K, _ptr(item_bias), _ptr(topi), _ptr(topv), _ptr(trank),
                                           _ptr(tscore), _ptr(scratch), nfl, _stream(dev)), "recad_fullrank_eval_tc")
        return topi, topv, trank[:, :T], tscore[:, :T]
    if item_bias is not None:        # exact kernel: the bias rides along as one more embedding dimension
        user_emb = torch.cat([user_emb, torch.ones((user_emb.shape[0], 1), device=dev)], 1)
        item_emb = torch.cat([item_emb, item_bias.view(-1, 1)], 1)
        D, item_T = D + 1, None
    if item_T is None:
        item_T = transpose_items(item_emb)
    with torch.cuda.device(dev):
        check(_lib.lib().recad_fullrank_eval(_ptr(user_emb.contiguous()), _ptr(item_T), item_T.shape[1], I, D,
                                             _ptr(user_ids.contiguous().long()), n, _ptr(train_rowptr), _ptr(train_col),
                                             _ptr(targets_t), T, K, _ptr(topi), _ptr(topv), _ptr(trank), _ptr(tscore),
                                             _stream(dev)), "recad_fullrank_eval")
    return topi, topv, trank[:, :T], tscore[:, :T]


def rank_from_scores(scores, user_ids, train_rowptr, train_col, targets, K):
    _need_cuda(scores, user_ids, train_rowptr, train_col)
    dev = scores.device
    n, I = scores.shape
    targets_t = torch.as_tensor(list(targets), dtype=torch.int32, device=dev)
    T = int(targets_t.numel())
    if T and (int(targets_t.min()) < 0 or int(targets_t.max()) >= I):
        raise RecadError(f"target item id out of range [0, {I})")
    topi, topv, trank, tscore = _eval_outputs(n, T, K, dev)
    with torch.cuda.device(dev):
        check(_lib.lib().recad_rank_from_scores(_ptr(scores.contiguous()), n, I, _ptr(user_ids.contiguous().long()),
                                                _ptr(train_rowptr), _ptr(train_col), _ptr(targets_t), T, K, _ptr(topi),
                                                _ptr(topv), _ptr(trank), _ptr(tscore), _stream(dev)), "recad_rank_from_scores")
    return topi, topv, trank[:, :T], tscore[:, :T]


def eligible_users(train_rowptr, train_col, n_users, n_items, targets, is_key=None):
    """Candidate users of the evaluation (normal.py:133-143) as an ascending int64 device tensor."""
    _need_cuda(train_rowptr, train_col, is_key)
    dev = train_rowptr.device
    L = _lib.lib()
    with torch.cuda.device(dev):
        t = torch.as_tensor(list(targets), dtype=torch.int32, device=dev)
        out = torch.empty(n_users, dtype=torch.int64, device=dev)
        nbytes = L.recad_eligible_users_scratch_bytes(n_users)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        n = C.c_int64()
        check(L.recad_eligible_users(_ptr(train_rowptr), _ptr(train_col), n_users, n_items, _ptr(t), int(t.numel()), _ptr(is_key),
                                     _ptr(out), C.byref(n), _ptr(scratch), nbytes, _stream(dev)), "recad_eligible_users")
    return out[:n.value]


def recall_ndcg(topk_idx, user_ids, gt_rowptr, gt_col):
    """Sums of Recall@K / NDCG@K and the number of users with ground truth."""
    _need_cuda(topk_idx, user_ids, gt_rowptr, gt_col)
    out = torch.zeros(3, dtype=torch.float64, device=topk_idx.device)
    with torch.cuda.device(topk_idx.device):
        check(_lib.lib().recad_recall_ndcg(_ptr(topk_idx), topk_idx.shape[0], topk_idx.shape[1], _ptr(user_ids.contiguous().long()),
                                           _ptr(gt_rowptr), _ptr(gt_col), _ptr(out), _stream(topk_idx.device)), "recad_recall_ndcg")
    return out


# --------------------------------------------------------------------------- #
# samplers (host; advance numpy's legacy global MT19937 state in place)
# --------------------------------------------------------------------------- #
def _np_state():
    st = np.random.get_state()
    if st[0] != "MT19937":
        raise RecadError("np.random global state is not MT19937")
    return st, np.ascontiguousarray(st[1], dtype=np.uint32).copy(), [int(st[2])]


def _np_state_commit(st, key, pos):
    np.random.set_state((st[0], key, int(pos[0]), st[3], st[4]))


def _np(a, dtype=np.int64):
    return np.ascontiguousarray(a, dtype=dtype)


# *_raw: explicit MT19937 state -- key: uint32[624] (advanced in place), pos: one-element list.
# The ctypes calls release the GIL, so a raw sampler can run on a background thread.
def mt_pairwise_raw(key, pos, n_users, n_items, train_size, allpos_rowptr, allpos_col, out=None):
    rp, col = _np(allpos_rowptr), _np(allpos_col, np.int32)
    if out is None:
        out = np.empty((max(train_size, 1), 3), dtype=np.int64)
    assert out.dtype == np.int64 and out.flags.c_contiguous and out.shape[0] >= train_size
    n_out, cpos = C.c_int64(), C.c_int32(pos[0])
    if train_size >= FAST_SAMPLER_MIN:
        filt, ext = pairwise_filter(rp, col, n_users)
        check(_lib.lib().recad_mt19937_pairwise_fast(key.ctypes.data, C.byref(cpos), n_users, n_items, train_size, rp.ctypes.data,
                                                     col.ctypes.data, filt.ctypes.data, ext.ctypes.data, min(os.cpu_count() or 1, 32),
                                                     out.ctypes.data, C.byref(n_out)), "recad_mt19937_pairwise_fast")
    else:
        check(_lib.lib().recad_mt19937_pairwise(key.ctypes.data, C.byref(cpos), n_users, n_items, train_size, rp.ctypes.data,
                                                col.ctypes.data, out.ctypes.data, C.byref(n_out)), "recad_mt19937_pairwise")
    pos[0] = cpos.value
    return out[:n_out.value]


def mt_pairwise_epoch_raw(key, pos, n_users, n_items, train_size, allpos_rowptr, allpos_col, out, j_out):
    """Sampler + the draws of the epoch shuffle as one call (large epochs only): -> (samples [n, 3], j [n]); the draws run
    on the calling thread while the other threads still write the rows.  Same stream consumption as
    mt_pairwise_raw followed by mt_permutation_draw_raw."""
    if train_size < FAST_SAMPLER_MIN:
        S = mt_pairwise_raw(key, pos, n_users, n_items, train_size, allpos_rowptr, allpos_col, out=out)
        return S, mt_permutation_draw_raw(key, pos, len(S), j_out)
    rp, col = _np(allpos_rowptr), _np(allpos_col, np.int32)
    assert out.dtype == np.int64 and out.flags.c_contiguous and out.shape[0] >= train_size
    assert j_out.dtype == np.uint32 and j_out.flags.c_contiguous and j_out.shape[0] >= train_size
    filt, ext = pairwise_filter(rp, col, n_users)
    n_out, cpos = C.c_int64(), C.c_int32(pos[0])
    check(_lib.lib().recad_mt19937_pairwise_epoch(key.ctypes.data, C.byref(cpos), n_users, n_items, train_size, rp.ctypes.data,
                                                  col.ctypes.data, filt.ctypes.data, ext.ctypes.data, min(os.cpu_count() or 1, 32),
                                                  out.ctypes.data, C.byref(n_out), j_out.ctypes.data), "recad_mt19937_pairwise_epoch")
    pos[0] = cpos.value
    return out[:n_out.value], j_out[:n_out.value]


def mt_pairwise_soa_raw(key, pos, n_users, n_items, train_size, allpos_rowptr, allpos_col, users, rel, negs, j_out=None):
    """The large-epoch sampler (csrc/sampler.cpp, recad_mt19937_pairwise_soa): same samples and stream consumption as
    mt_pairwise_raw (+ mt_permutation_draw_raw when j_out is given), written as three uint32 arrays
    (user, index of the positive inside the user's row, negative).  -> number of samples."""
    rp, col = _np(allpos_rowptr), _np(allpos_col, np.int32)
    for a in (users, rel, negs) + ((j_out,) if j_out is not None else ()):
        assert a.dtype == np.uint32 and a.flags.c_contiguous and a.shape[0] >= train_size
    filt, ext = pairwise_filter(rp, col, n_users)
    n_out, cpos = C.c_int64(), C.c_int32(pos[0])
    check(_lib.lib().recad_mt19937_pairwise_soa(key.ctypes.data, C.byref(cpos), n_users, n_items, train_size, rp.ctypes.data,
                                                col.ctypes.data, filt.ctypes.data, ext.ctypes.data, min(os.cpu_count() or 1, 32),
                                                users.ctypes.data, rel.ctypes.data, negs.ctypes.data, C.byref(n_out),
                                                j_out.ctypes.data if j_out is not None else None), "recad_mt19937_pairwise_soa")
    pos[0] = cpos.value
    return n_out.value


def permutation_apply32(j, out):
    n = int(j.shape[0])
    assert out.dtype == np.int32 and out.flags.c_contiguous and out.shape[0] >= n
    check(_lib.lib().recad_permutation_apply32(n, j.ctypes.data, out.ctypes.data), "recad_permutation_apply32")
    return out[:n]


def samples_expand(allpos_rowptr, allpos_col, users, rel, negs, rows):
    """rows[k] = (users[k], allpos_col[allpos_rowptr[users[k]] + rel[k]], negs[k]) on the device (int32 [n, 3])."""
    _need_cuda(allpos_rowptr, allpos_col, users, rel, negs, rows)
    n = int(users.shape[0])
    assert rows.dtype == torch.int32 and rows.shape[0] >= n and allpos_rowptr.dtype == torch.int64 and allpos_col.dtype == torch.int32
    with torch.cuda.device(rows.device):
        check(_lib.lib().recad_samples_expand(_ptr(allpos_rowptr), _ptr(allpos_col), _ptr(users), _ptr(rel), _ptr(negs), n,
                                              _ptr(rows), _stream(rows.device)), "recad_samples_expand")
    return rows


def host_empty(shape, dtype):
    """np.empty whose pages are requested as transparent huge pages (must be called before first touch)."""
    a = np.empty(shape, dtype=dtype)
    if a.nbytes >= (8 << 20):
        _lib.lib().recad_host_advise_huge(a.ctypes.data, a.nbytes)
    return a


def to_host(t, dtype=None):
    """Device tensor -> numpy array backed by huge pages (for arrays the samplers access at random)."""
    out = host_empty(tuple(t.shape), dtype or {torch.int64: np.int64, torch.int32: np.int32, torch.float32: np.float32}[t.dtype])
    torch.from_numpy(out).copy_(t)
    return out


FAST_SAMPLER_MIN = 1 << 20      # samples per epoch from which the filter-based parser pays off


class _DerivedCache:
    """Per-dataset host structures derived from the positives (filter blocks, sorted lists), keyed by the identity of
    the source arrays.  The entry keeps those arrays alive, so an address can never be reused by another dataset while
    its derived data is cached; a lock makes lookup-or-build atomic (the epochs of two datasets -- the clean one's
    speculative prefetch and the attacked one's first draw -- do run concurrently)."""

    def __init__(self, capacity):
        import threading
        from collections import OrderedDict
        self.capacity, self.lock, self.entries = capacity, threading.Lock(), OrderedDict()

    def get(self, key, keepalive, build):
        with self.lock:
            if key in self.entries:
                self.entries.move_to_end(key)
                return self.entries[key][0]
            value = build()
            self.entries[key] = (value, keepalive)
            while len(self.entries) > self.capacity:
                self.entries.popitem(last=False)
            return value


_filters = _DerivedCache(3)
_sorted_lists = _DerivedCache(4)


_filter_parents = {}     # (new col address, new rowptr address) -> (parent rowptr, parent col, parent n_users): injected datasets


def filter_parent_hint(new_rowptr, new_col, old_rowptr, old_col, n_users_old):
    """The positives (new_rowptr, new_col) are (old_rowptr, old_col) with rows APPENDED (dataset injection): the sampler's
    filter blocks of the old users can be copied instead of rebuilt."""
    if len(_filter_parents) > 8:
        _filter_parents.clear()
    _filter_parents[(new_col.ctypes.data, new_rowptr.ctypes.data)] = (old_rowptr, old_col, int(n_users_old))


def _filter_key(rowptr, col, n_users):
    return (col.ctypes.data, rowptr.ctypes.data, len(col), int(n_users))


def pairwise_filter(allpos_rowptr, allpos_col, n_users):
    """Per-user 128-byte blocks (row bounds + two-level membership filter, first line decisive) for mt_pairwise_raw(fast);
    built once per positives array -- or, for a dataset derived by injection from one whose blocks are still cached,
    copied from the parent and completed for the appended users only."""
    def build():
        filt = host_empty(int(n_users) * 16, np.uint64)
        ext = host_empty(max(len(allpos_col), 1), np.uint32)
        u_lo = 0
        parent = _filter_parents.pop((allpos_col.ctypes.data, allpos_rowptr.ctypes.data), None)
        if parent is not None:
            prp, pcol, pn = parent
            got = _filters.entries.get(_filter_key(prp, pcol, pn))        # build() runs under the cache's lock
            if got is not None and pn <= n_users:
                _filters.entries.move_to_end(_filter_key(prp, pcol, pn))    # a parent in use is not the eviction candidate
                (pf, pe), _ = got
                filt[:pn * 16] = pf[:pn * 16]
                ext[:len(pe)] = pe
                u_lo = pn
        check(_lib.lib().recad_pairwise_filter_build_range(allpos_rowptr.ctypes.data, allpos_col.ctypes.data, u_lo, n_users,
                                                           filt.ctypes.data, ext.ctypes.data, os.cpu_count() or 1),
              "recad_pairwise_filter_build_range")
        return filt, ext
    return _filters.get(_filter_key(allpos_rowptr, allpos_col, n_users), (allpos_rowptr, allpos_col), build)


def mt_pointwise_raw(key, pos, user_ids, pos_rowptr, pos_items, n_items, ratio, out=None):
    uid, rp, items = _np(user_ids), _np(pos_rowptr), _np(pos_items)

    def build():                                # once per dataset, not once per epoch (a 0.1 s lexsort at ml1m size)
        rows = np.repeat(np.arange(len(uid), dtype=np.int64), np.diff(rp))
        return items[np.lexsort((items, rows))]   # each user's list sorted ascending, lists kept in place
    srt = _sorted_lists.get((items.ctypes.data, rp.ctypes.data, len(items)), (items, rp), build)
    n = len(items) * (1 + ratio)
    if out is None:
        out = np.empty((max(n, 1), 3), dtype=np.int64)
    assert out.dtype == np.int64 and out.flags.c_contiguous and out.shape[0] >= n
    cpos = C.c_int32(pos[0])
    check(_lib.lib().recad_mt19937_pointwise(key.ctypes.data, C.byref(cpos), len(uid), uid.ctypes.data, rp.ctypes.data,
                                             items.ctypes.data, srt.ctypes.data, n_items, ratio, out.ctypes.data),
          "recad_mt19937_pointwise")
    pos[0] = cpos.value
    return out[:n]


def mt_permutation_raw(key, pos, n, out=None):
    perm = np.empty(max(n, 1), dtype=np.int64) if out is None else out
    assert perm.dtype == np.int64 and perm.flags.c_contiguous and perm.shape[0] >= n
    cpos = C.c_int32(pos[0])
    check(_lib.lib().recad_mt19937_permutation(key.ctypes.data, C.byref(cpos), n, perm.ctypes.data), "recad_mt19937_permutation")
    pos[0] = cpos.value
    return perm[:n]


def mt_permutation_draw_raw(key, pos, n, j_out=None):
    """The stream-consuming half of the shuffle: j[i] = random_interval(i), i = n-1 .. 1 (uint32 [n])."""
    j = host_empty(max(n, 1), np.uint32) if j_out is None else j_out
    assert j.dtype == np.uint32 and j.flags.c_contiguous and j.shape[0] >= n
    cpos = C.c_int32(pos[0])
    check(_lib.lib().recad_mt19937_permutation_draw(key.ctypes.data, C.byref(cpos), n, j.ctypes.data), "recad_mt19937_permutation_draw")
    pos[0] = cpos.value
    return j[:n]


def permutation_apply(j, out=None):
    """The generator-free half: perm = arange(n) with swap(perm[i], perm[j[i]]) applied for i = n-1 .. 1."""
    n = int(j.shape[0])
    perm = np.empty(max(n, 1), dtype=np.int64) if out is None else out
    assert perm.dtype == np.int64 and perm.flags.c_contiguous and perm.shape[0] >= n
    check(_lib.lib().recad_permutation_apply(n, j.ctypes.data, perm.ctypes.data), "recad_permutation_apply")
    return perm[:n]


def mt_pairwise(n_users, n_items, train_size, allpos_rowptr, allpos_col, out=None):
    """== pairwise_sample (implicit.py:50-74) on the global np.random stream."""
    st, key, pos = _np_state()
    S = mt_pairwise_raw(key, pos, n_users, n_items, train_size, allpos_rowptr, allpos_col, out)
    _np_state_commit(st, key, pos)
    return S


def mt_pointwise(user_ids, pos_rowptr, pos_items, n_items, ratio, out=None):
    """== pointwise_sample (implicit.py:77-91) on the global np.random stream."""
    st, key, pos = _np_state()
    S = mt_pointwise_raw(key, pos, user_ids, pos_rowptr, pos_items, n_items, ratio, out)
    _np_state_commit(st, key, pos)
    return S


def mt_permutation(n, out=None):
    """== np.random.shuffle(np.arange(n)) (implicit.py:24-25)."""
    st, key, pos = _np_state()
    perm = mt_permutation_raw(key, pos, n, out)
    _np_state_commit(st, key, pos)
    return perm
