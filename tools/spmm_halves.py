"""Time the user-row and item-row halves of the propagation SpMM separately (L2 locality study)."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from recad_b200 import ops
w = bench.WORKLOADS["synthetic"]
dev = torch.device("cuda:0")
eu, ei = bench.synth_edges(w, dev)
U, I, D = w["n_users"], w["n_items"], 64
g = ops.Graph.from_edges(eu, ei, U, I)
del eu, ei
nnz_u = int(g.rowptr[U])
g_user = ops.Graph.from_csr(g.rowptr[:U + 1], g.colidx[:nnz_u] - U, g.vals[:nnz_u], n_cols=I)
g_item = ops.Graph.from_csr(g.rowptr[U:] - nnz_u, g.colidx[nnz_u:], g.vals[nnz_u:], n_cols=U)
X = torch.randn(U + I, D, device=dev) * 0.1
Y = torch.empty_like(X)
def t(fn, n=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
print(json.dumps({"full_ms": t(lambda: ops.spmm(g, X, Y)),
                  "user_rows_ms (gathers 51 MB item table)": t(lambda: ops.spmm(g_user, X[U:], Y[:U])),
                  "item_rows_ms (gathers 256 MB user table)": t(lambda: ops.spmm(g_item, X[:U], Y[U:])),
                  "n_seg": [g.n_seg, g_user.n_seg, g_item.n_seg], "n_mrow": [g.n_mrow, g_user.n_mrow, g_item.n_mrow]}))
# column-blocked item rows: 4 user blocks, each gathered from an L2-sized slice
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4
rows = torch.repeat_interleave(torch.arange(I, device=dev), g_item.rowptr[1:] - g_item.rowptr[:-1])
tot = 0.0
for b in range(nb):
    lo, hi = b * U // nb, (b + 1) * U // nb
    sel = (g_item.colidx >= lo) & (g_item.colidx < hi)
    r = rows[sel]
    ptr = torch.zeros(I + 1, dtype=torch.int64, device=dev)
    ptr[1:] = torch.cumsum(torch.bincount(r, minlength=I), 0)
    gb = ops.Graph.from_csr(ptr, (g_item.colidx[sel] - lo).contiguous(), g_item.vals[sel].contiguous(), n_cols=hi - lo)
    ms = t(lambda: ops.spmm(gb, X[lo:hi], Y[U:]))
    tot += ms
    print(json.dumps({"block": b, "nnz": int(sel.sum()), "ms": ms, "n_seg": gb.n_seg}))
print(json.dumps({"item_rows_column_blocked_total_ms": tot, "blocks": nb}))
