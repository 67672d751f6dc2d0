"""Sweep of the second-generation SpMM (csrc/spmm2.cu) on the synthetic graph: gather width x unroll x register cap x
plan (segment length, column blocks of the item rows).  Checks every plan against the first kernel, then times the
full product and its two halves (user rows gather the L2-resident item table; item rows gather the 256 MB user table).

    python tools/spmm2_sweep.py [--workload synthetic] [--variants 0,1,4,5] [--blocks 1,4,6] [--seg-lens 256]
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from recad_b200 import _lib, ops  # noqa: E402


def t_ms(fn, n=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="synthetic")
    ap.add_argument("--variants", default="0,1,4,5,8,9,16,17,21,32,33,36,37,52,53")
    ap.add_argument("--blocks", default="1,4,6")
    ap.add_argument("--seg-lens", default="256")
    ap.add_argument("--order", default="row", help="row | long (longest segments first inside each group)")
    ap.add_argument("--hot", default="800", help="hot-row counts tried by the variants >= 100")
    a = ap.parse_args()
    w = bench.WORKLOADS[a.workload]
    dev = torch.device("cuda:0")
    eu, ei = bench.synth_edges(w, dev)
    U, I, D = w["n_users"], w["n_items"], w["D"]
    g = ops.Graph.from_edges(eu, ei, U, I)
    del eu, ei
    N = U + I
    X = torch.randn(N, D, device=dev) * 0.1
    Y0, Z0 = torch.empty_like(X), torch.empty_like(X)
    ops.spmm(g, X, Y0, X, Z0, 0.5)
    old_ms = t_ms(lambda: ops.spmm(g, X, Y0, X, Z0, 0.5))
    print(json.dumps({"kernel": "spmm_seg_kernel (round 1)", "ms": round(old_ms, 4), "n_seg": g.n_seg, "n_mrow": g.n_mrow}), flush=True)
    cv = ops.pack_cv(g.colidx, g.vals)
    lib = _lib.lib()
    nnz_u = int(g.rowptr[U])

    def hot_copy(H):          # the H most popular items live in shared memory: their columns become slot ids (bit 31 set)
        top = torch.topk(g.degree[U:], H).indices
        slot_of = torch.full((N,), -1, dtype=torch.int64, device=dev)
        slot_of[U + top] = torch.arange(H, device=dev)
        cvh = cv.clone()
        c = cvh[:nnz_u, 0].long()
        sl = slot_of[c]
        cvh[:nnz_u, 0] = torch.where(sl >= 0, sl - (1 << 31), c).to(torch.int32)
        share = float((sl >= 0).float().mean())
        return cvh, (U + top).to(torch.int32).contiguous(), share
    hots = {H: hot_copy(H) for H in [int(h) for h in a.hot.split(",") if h]}
    Y, Z = torch.empty_like(X), torch.empty_like(X)
    for seg_len in [int(s) for s in a.seg_lens.split(",")]:
        for nb in [int(b) for b in a.blocks.split(",")]:
            plan = ops.PackedPlan.build(g.rowptr, g.colidx, N, seg_len, first_blocked_row=U if nb > 1 else None,
                                        n_col_blocks=nb, n_block_cols=U, order=a.order)
            partials = torch.empty(max(plan.n_slot, 1) * D, dtype=torch.float32, device=dev)
            for v, H in [(int(x), h) for x in a.variants.split(",") for h in (hots if int(x) >= 100 else [0])]:
                cvv, hot_rows, share = hots[H] if H else (cv, None, 0.0)

                def run(meta=plan.meta, n_mrow=plan.n_mrow):
                    _lib.check(lib.recad_spmm_packed(meta.data_ptr(), meta.shape[0], cvv.data_ptr(), n_mrow, plan.mrow.data_ptr(),
                                                     plan.mrow_lo.data_ptr(), partials.data_ptr(), X.data_ptr(), Y.data_ptr(),
                                                     X.data_ptr(), Z.data_ptr(), 0.5, D, v, hot_rows.data_ptr() if H else None, H,
                                                     ops._stream(dev)), "recad_spmm_packed")
                Y.zero_(); Z.zero_()
                run()
                torch.cuda.synchronize()
                err = max(float((Y - Y0).abs().max()), float((Z - Z0).abs().max()))
                ms = t_ms(run)
                nu = plan.n_first                       # segments of the unblocked (user) rows
                ms_u = t_ms(lambda: run(plan.meta[:nu], 0))
                ms_i = t_ms(lambda: run(plan.meta[nu:], 0)) if plan.meta.shape[0] > nu else 0.0
                print(json.dumps({"variant": v, "seg_len": seg_len, "col_blocks": nb, "ms": round(ms, 4), "user_rows_ms": round(ms_u, 4),
                                  "item_rows_ms": round(ms_i, 4), "max_abs_err_vs_r1": err, "n_seg": int(plan.meta.shape[0]),
                                  "n_mrow": plan.n_mrow, "n_slot": plan.n_slot, "order": a.order, "hot": H, "hot_share_of_user_row_gathers": round(share, 4)}), flush=True)


if __name__ == "__main__":
    main()
