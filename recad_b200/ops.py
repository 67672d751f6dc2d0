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


class Graph:
    """Device-resident CSR of a (normalised) adjacency + its SpMM work plan."""

    def __init__(self, n_rows, n_cols, rowptr, colidx, vals, mult=None, degree=None, seg_len=SEG_LEN):
        self.n_rows, self.n_cols = int(n_rows), int(n_cols)
        self.rowptr, self.colidx, self.vals, self.mult, self.degree = rowptr, colidx, vals, mult, degree
        self.nnz = int(colidx.numel())
        self.device = rowptr.device
        self.seg_len = seg_len or auto_seg_len(self.nnz)
        self._partials = None
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
        return cls(N, N, rowptr, colidx, vals, mult, degree, seg_len)

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
        return Graph(Nn, Nn, rowptr, colidx, vals, mult, degree, self.seg_len)

    def renormalized(self, degree):
        """The same structure with values recomputed from new degrees (a shard whose own rows did not
        change while fake users elsewhere changed the item degrees)."""
        with torch.cuda.device(self.device):
            vals = self._normalize(self.rowptr, self.colidx, self.mult, degree, self.n_rows)
        return Graph(self.n_rows, self.n_cols, self.rowptr, self.colidx, vals, self.mult, degree, self.seg_len)

    @classmethod
    def from_csr(cls, rowptr, colidx, vals, n_cols, seg_len=SEG_LEN):
        """Wrap an existing (possibly rectangular) CSR, e.g. a user-row shard."""
        _need_cuda(rowptr, colidx, vals)
        return cls(rowptr.numel() - 1, n_cols, rowptr.contiguous().long(), colidx.contiguous().int(),
                   vals.contiguous().float(), None, None, seg_len)

    # -- plan ---------------------------------------------------------------- #
    def _plan(self):
        L = _lib.lib()
        dev = self.device
        with torch.cuda.device(dev):
            cap = L.recad_spmm_plan_max_segments(self.n_rows, self.nnz, self.seg_len)
            seg_row = torch.empty(cap, dtype=torch.int32, device=dev)
            seg_lo = torch.empty(cap, dtype=torch.int64, device=dev)
            seg_slot = torch.empty(cap, dtype=torch.int32, device=dev)
            mcap = self.nnz // self.seg_len + 2
            mrow = torch.empty(mcap, dtype=torch.int32, device=dev)
            mrow_lo = torch.empty(mcap + 1, dtype=torch.int32, device=dev)
            nbytes = L.recad_spmm_plan_scratch_bytes(self.n_rows)
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            counts = (C.c_int64 * 3)()
            check(L.recad_spmm_plan(_ptr(self.rowptr), self.n_rows, self.seg_len, _ptr(seg_row), _ptr(seg_lo), _ptr(seg_slot),
                                    _ptr(mrow), _ptr(mrow_lo), counts, _ptr(scratch), nbytes, _stream(dev)), "recad_spmm_plan")
        self.n_seg, self.n_mrow, self.n_slot = (int(c) for c in counts)
        self.seg_row, self.seg_lo, self.seg_slot = seg_row[:self.n_seg], seg_lo[:self.n_seg], seg_slot[:self.n_seg]
        self.mrow, self.mrow_lo = mrow[:max(self.n_mrow, 1)], mrow_lo[:self.n_mrow + 1]
        self._struct = None

    def struct(self, D):
        """ctypes recad_csr for embedding width D (allocates the partial-sum scratch lazily)."""
        need = max(self.n_slot, 1) * D
        if self._partials is None or self._partials.numel() < need:
            self._partials = torch.empty(need, dtype=torch.float32, device=self.device)
            self._struct = None
        if self._struct is None:
            s = _lib.CSR()
            s.n_rows, s.nnz = self.n_rows, self.nnz
            s.rowptr, s.colidx, s.vals = self.rowptr.data_ptr(), self.colidx.data_ptr(), self.vals.data_ptr()
            s.n_seg, s.seg_len = self.n_seg, self.seg_len
            s.seg_row, s.seg_lo, s.seg_slot = self.seg_row.data_ptr(), self.seg_lo.data_ptr(), self.seg_slot.data_ptr()
            s.n_mrow, s.mrow, s.mrow_lo = self.n_mrow, self.mrow.data_ptr(), self.mrow_lo.data_ptr()
            s.partials = self._partials.data_ptr()
            self._struct = s
        return self._struct

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


def pairwise_filter(allpos_rowptr, allpos_col, n_users):
    """Per-user 128-byte blocks (row bounds + two-level membership filter, first line decisive) for mt_pairwise_raw(fast); built once per positives array."""
    def build():
        filt = host_empty(int(n_users) * 16, np.uint64)
        ext = host_empty(max(len(allpos_col), 1), np.uint32)
        check(_lib.lib().recad_pairwise_filter_build(allpos_rowptr.ctypes.data, allpos_col.ctypes.data, n_users,
                                                     filt.ctypes.data, ext.ctypes.data, os.cpu_count() or 1),
              "recad_pairwise_filter_build")
        return filt, ext
    return _filters.get((allpos_col.ctypes.data, allpos_rowptr.ctypes.data, len(allpos_col), int(n_users)),
                        (allpos_rowptr, allpos_col), build)


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
