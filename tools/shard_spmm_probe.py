"""One rank's two SpMM halves of the user-row sharded path, timed on ONE GPU (no exchange): rank 0's shard of the synthetic
graph for a given world size -- g_user (own user rows gathering the item replica) and g_item (all item rows restricted to
the rank's users = the partial sums that are pushed to their owners).  Tells what the sharded epoch's SpMM phases cost by
themselves, next to the time the entry count alone would take at the single-GPU rate.

    python tools/shard_spmm_probe.py [--world 8] [--reps 20]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from recad_b200 import ops  # noqa: E402
from recad_b200.dist import user_range  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    w = bench.WORKLOADS["synthetic"]
    dev = torch.device("cuda:0")
    eu, ei = bench.synth_edges(w, dev)
    U, I, D = w["n_users"], w["n_items"], w["D"]
    lo, hi = user_range(U, 0, a.world)
    Ug = hi - lo
    mine = (eu >= lo) & (eu < hi)
    g = ops.Graph.from_edges((eu[mine] - lo).contiguous(), ei[mine].contiguous(), Ug, I)
    nnz_u = int(g.rowptr[Ug])
    g_user = ops.Graph.from_csr(g.rowptr[:Ug + 1], g.colidx[:nnz_u] - Ug, g.vals[:nnz_u], n_cols=I)
    g_item = ops.Graph.from_csr(g.rowptr[Ug:] - nnz_u, g.colidx[nnz_u:], g.vals[nnz_u:], n_cols=Ug)
    Xu, Xi = torch.randn(Ug, D, device=dev), torch.randn(I, D, device=dev)
    Yu, Yi = torch.empty_like(Xu), torch.empty_like(Xi)

    def t_ms(fn):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.reps
    per_entry_ms = 2.21 / 100e6                        # the single-GPU product: 2.21 ms for 100 M entries
    out = {"world": a.world, "users_local": Ug, "entries_per_half": nnz_u,
           "user_rows": {"rows": Ug, "segments": g_user.n_seg, "ms": round(t_ms(lambda: ops.spmm(g_user, Xi, Yu)), 4)},
           "item_rows": {"rows": I, "segments": g_item.n_seg, "multi_segment_rows": g_item.n_mrow,
                         "ms": round(t_ms(lambda: ops.spmm(g_item, Xu, Yi)), 4)},
           "ms_by_entry_count_at_the_single_gpu_rate": round(nnz_u * per_entry_ms, 4)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
