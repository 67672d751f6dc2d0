"""The propagation SpMM on the synthetic graph: whole product and its two halves (user rows gather the L2-resident
item table; item rows gather the 256 MB user table, column-blocked), timed with CUDA events -- or, with --ncu, launched
a fixed number of times so that `ncu -k regex:spmm_kernel` sees [full, user rows, item rows] x reps in that order.

    python tools/spmm_probe.py [--workload synthetic] [--ncu] [--reps 10]
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="synthetic")
    ap.add_argument("--ncu", action="store_true")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--check", action="store_true", help="compare with torch.sparse.mm (fp64 accumulate)")
    ap.add_argument("--dim", type=int, default=0, help="table width (default: the workload's D); 8 / 16 / 32 = a column shard")
    a = ap.parse_args()
    w = bench.WORKLOADS[a.workload]
    dev = torch.device("cuda:0")
    eu, ei = bench.synth_edges(w, dev)
    U, I, D = w["n_users"], w["n_items"], a.dim or w["D"]
    g = ops.Graph.from_edges(eu, ei, U, I)
    del eu, ei
    N = U + I
    X = torch.randn(N, D, device=dev) * 0.1
    Y, Z = torch.empty_like(X), torch.empty_like(X)
    lib = _lib.lib()
    s_full = g.struct(D)

    def part(lo, hi):                      # the same matrix restricted to plan segments [lo, hi)
        s = _lib.CSR.from_buffer_copy(s_full)
        s.n_seg = hi - lo
        s.n_rows = min(s.n_rows, s.n_seg)          # only the argument check reads it
        s.seg_meta = g.plan.meta.data_ptr() + lo * 16
        return s
    (u0, u1), (i0, i1) = g.plan.group_segs
    structs = {"full": s_full, "user_rows": part(u0, u1), "item_rows": part(i0, i1)}

    def run(s):
        _lib.check(lib.recad_spmm(C.byref(s), X.data_ptr(), Y.data_ptr(), X.data_ptr(), Z.data_ptr(), 0.25, D, ops._stream(dev)), "recad_spmm")

    def t_ms(s, n):
        for _ in range(3):
            run(s)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            run(s)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    if a.ncu:
        for _ in range(2):
            for s in structs.values():
                run(s)
        torch.cuda.synchronize()
        return
    out = {"workload": a.workload, "D": D, "nnz": g.nnz, "n_seg": g.n_seg, "n_mrow": g.n_mrow, "n_slot": g.n_slot, "seg_len": g.seg_len,
           "group_segs": g.plan.group_segs, "algorithmic_bytes": g.algorithmic_bytes(D) + 2 * N * 4 * D}
    for k, s in structs.items():
        out[k + "_ms"] = round(t_ms(s, a.reps), 4)
    if a.check:
        run(s_full)
        A = torch.sparse_csr_tensor(g.rowptr, g.colidx.long(), g.vals.double(), (N, N))
        ref = torch.sparse.mm(A, X.double())
        out["max_rel_err_vs_fp64_sparse_mm"] = float(((Y.double() - ref).abs().max() / ref.abs().max()).item())
        out["max_rel_err_Z"] = float(((Z.double() - 0.25 * (X.double() + ref)).abs().max() / ref.abs().max()).item())
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
