"""Time the fused SpMM under every RECAD_SPMM_VARIANT on the synthetic graph (one process per variant).
    python tools/spmm_sweep.py [--workload synthetic] [--variants 0,1,2,...]"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(workload, seg_len):
    import torch
    import bench
    from recad_b200 import ops
    w = bench.WORKLOADS[workload]
    dev = torch.device("cuda:0")
    eu, ei = bench.synth_edges(w, dev)
    g = ops.Graph.from_edges(eu, ei, w["n_users"], w["n_items"], seg_len=seg_len)
    del eu, ei
    N, D = g.n_rows, w["D"]
    X = torch.randn(N, D, device=dev) * 0.1
    Y, Z = torch.empty_like(X), torch.empty_like(X)
    for _ in range(3):
        ops.spmm(g, X, Y, X, Z, 1.0)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        ops.spmm(g, X, Y, X, Z, 1.0)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    alg = g.algorithmic_bytes(D) + 2 * N * 4 * D
    print(json.dumps({"variant": int(os.environ.get("RECAD_SPMM_VARIANT", -1)), "seg_len": seg_len, "ms": round(ms, 4),
                      "alg_GBps": round(alg / ms / 1e6, 1), "n_seg": g.n_seg, "n_mrow": g.n_mrow}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="synthetic")
    ap.add_argument("--variants", default="0,1,2,3,4,5,6,7,9,13")
    ap.add_argument("--seg-lens", default="256")
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--seg-len", type=int, default=256)
    a = ap.parse_args()
    if a.child:
        child(a.workload, a.seg_len)
    else:
        for sl in a.seg_lens.split(","):
            for v in a.variants.split(","):
                env = dict(os.environ, RECAD_SPMM_VARIANT=v)
                r = subprocess.run([sys.executable, __file__, "--child", "--workload", a.workload, "--seg-len", sl], env=env,
                                   capture_output=True, text=True)
                print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:], flush=True)
