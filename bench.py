"""Headline benchmark of the victim hot path (BASELINE.json): one LightGCN BPR epoch + one
full-ranking evaluation on the synthetic 1M x 200k x 50M graph (D=64, L=3), plus the SpMM roofline.

    python bench.py --gpus 1 --steps 2 --warmup 3                  # this implementation
    python bench.py --impl reference --gpus 1 --steps 1 --warmup 0 # reference arm: the reference's own
                                                                   # CPU algorithm (oracle port) on host cores

A "step" = one BPR epoch (all batches: propagate, BPR, Horner backward, dense Adam) followed by one
full-rank evaluation of every user (top-20 + target rank, Recall/NDCG@20).  `value` is timed with
the epoch's samples already resident in HBM; `e2e` goes through the public API (`train_step()`,
`evaluate.*`): host sampler, pinned host->device copy of the samples, loss / metric read back.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[3] / SURVEY.md 8(d) input 4
    "synthetic": dict(n_users=1_000_000, n_items=200_000, n_edges=50_000_000, D=64, L=3, batch=1_048_576),
    # BASELINE.json configs[1] shape (ml1m.zip is absent: shape-matched synthetic), default batch of the reference
    "ml1m": dict(n_users=5950, n_items=3702, n_edges=468_649, D=64, L=3, batch=1024),
    # quick functional check of this script
    "tiny": dict(n_users=20_000, n_items=5_000, n_edges=400_000, D=64, L=3, batch=65_536),
}
METRIC = "lightgcn_bpr_epoch_plus_fullrank_eval_seconds"
L2_CAP_BYTES_PER_CLK = 6300            # full-chip L2 -> SM throughput cap (B300_MICROARCH.md, "LTS throughput cap")


def ncu_spmm_traffic(workload):
    """DRAM / L2 bytes of ONE SpMM launch from the committed ncu capture -- valid only for the kernel source it was
    taken from (profiles/ncu_spmm_r02.json records the hash of csrc/spmm.cu); None when the kernel changed since."""
    import hashlib
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "ncu_spmm_r02.json")))
        sha = hashlib.sha256(open(os.path.join(ROOT, "recad_b200", "csrc", "spmm.cu"), "rb").read()).hexdigest()[:16]
    except OSError:
        return None
    if rec.get("workload") != workload or rec.get("spmm_cu_sha16") != sha:
        return None
    return rec


def synth_edges(w, device, seed=0):
    """Distinct (u, i) pairs: user activity ~ lognormal, item popularity ~ Zipf(0.8), every user >= 1 item.
    Generated with torch on `device` (a 50M-edge graph takes seconds on the GPU)."""
    U, I, E = w["n_users"], w["n_items"], w["n_edges"]
    g = torch.Generator(device=device).manual_seed(seed)
    uw = torch.exp(torch.randn(U, device=device, generator=g, dtype=torch.float64))
    ucdf = torch.cumsum(uw / uw.sum(), 0)
    iw = torch.arange(1, I + 1, device=device, dtype=torch.float64) ** -0.8
    icdf = torch.cumsum(iw / iw.sum(), 0)
    perm = torch.randperm(I, device=device, generator=g)

    def draw(n):
        u = torch.searchsorted(ucdf, torch.rand(n, device=device, generator=g, dtype=torch.float64)).clamp_(max=U - 1)
        i = perm[torch.searchsorted(icdf, torch.rand(n, device=device, generator=g, dtype=torch.float64)).clamp_(max=I - 1)]
        return u * I + i
    first = torch.arange(U, device=device) * I + perm[
        torch.searchsorted(icdf, torch.rand(U, device=device, generator=g, dtype=torch.float64)).clamp_(max=I - 1)]
    keys = torch.unique(first)
    while keys.numel() < E:
        need = int((E - keys.numel()) * 1.3) + 1024
        keys = torch.unique(torch.cat([keys, draw(need)]))
    if keys.numel() > E:       # drop surplus at random, never a user's first pair
        u = keys // I
        is_first = torch.ones_like(u, dtype=torch.bool)
        is_first[1:] = u[1:] != u[:-1]
        cand = torch.nonzero(~is_first).flatten()
        drop = cand[torch.randperm(cand.numel(), device=device, generator=g)[: keys.numel() - E]]
        keep = torch.ones_like(u, dtype=torch.bool)
        keep[drop] = False
        keys = keys[keep]
    return keys // I, keys % I


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=lambda: self.rows.extend(self.proc.stdout.readlines()), daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6 or not f[0].isdigit():
                continue
            sm.append(int(f[0]))
            mx = max(mx, int(f[1]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# this implementation
# ----------------------------------------------------------------------------------------------
def run_b200(args, w):
    import torch.distributed as dist
    from recad_b200 import dataset, model, ops

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"        # the version banner goes to stdout, which carries ONE JSON line
        # N rank processes share the host with rank 0's sampler threads: a rank waiting for its GPU sleeps instead of spinning
        import ctypes
        try:
            ctypes.CDLL("libcudart.so.12").cudaSetDeviceFlags(4)        # cudaDeviceScheduleBlockingSync, before the context exists
        except OSError:
            pass
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        from recad_b200 import dist as rdist
        return rdist.bench_sharded(args, w, dev, rank, world, METRIC, synth_edges, ClockSampler, peaks,
                                   workload_string(args.workload, w, args.batch or w["batch"]))

    ev = lambda: torch.cuda.Event(enable_timing=True)     # noqa: E731
    U, I, D, L, B = w["n_users"], w["n_items"], w["D"], w["L"], args.batch or w["batch"]
    t0 = time.time()
    eu, ei = synth_edges(w, dev)
    torch.cuda.synchronize()
    t_gen = time.time() - t0
    t0 = time.time()
    graph = ops.Graph.from_edges(eu, ei, U, I)
    torch.cuda.synchronize()
    t_graph = time.time() - t0
    del eu, ei
    gtest = torch.Generator().manual_seed(1)          # one held-out item per user: ground truth of Recall/NDCG@20
    test = (np.arange(U, dtype=np.int64), torch.randint(0, I, (U,), generator=gtest).numpy())
    data = dataset.ArrayImplicitData(args.workload, U, I, None, dev, batch_size=B, graph=graph, test=test, prefetch=False)
    torch.manual_seed(2023)
    m = model.from_config("victim", "lightgcn", latent_dim_rec=D, lightGCN_n_layers=L, device=dev).I(dataset=data)
    rp, rc = data.train_csr(dev)
    users_all = torch.arange(U, device=dev)
    target = [0]
    # two resident sample sets, alternated between steps (the exact C++ MT19937 sampler produces them)
    np.random.seed(2023)
    t0 = time.time()
    sets = [data.epoch_samples(dev) for _ in range(2)]
    torch.cuda.synchronize()
    t_sampler = (time.time() - t0) / 2
    n = int(sets[0][0].shape[0])
    n_batches = (n + B - 1) // B
    gt_ptr, gt_col = data.ground_truth_csr("test", dev)
    def epoch(samples):
        rows, perm = samples
        m.run_epoch(rows, perm, B)
        m._steps += (int(rows.shape[0]) + B - 1) // B
        m._O_valid = False

    def full_eval():
        topi, topv, rank_, score, _ = m.full_rank(users_all, target, 20, rp, rc)
        return ops.recall_ndcg(topi, users_all, gt_ptr, gt_col), rank_

    for k in range(args.warmup):
        epoch(sets[k % 2])
        full_eval()
    torch.cuda.synchronize()
    marks = []
    with ClockSampler(local) as clocks:
        for k in range(args.steps):
            a, b, c = ev(), ev(), ev()
            a.record()
            epoch(sets[k % 2])
            b.record()
            full_eval()
            c.record()
            marks.append((a, b, c))
        torch.cuda.synchronize()
    ep_ms = [a.elapsed_time(b) for a, b, _ in marks]
    evl_ms = [b.elapsed_time(c) for _, b, c in marks]
    step_ms = float(np.mean(ep_ms) + np.mean(evl_ms))
    loss = m._read_loss(m.loss_acc, n_batches)[0]

    # dominant kernel: the fused SpMM (6 per batch).  Algorithmic bytes: SURVEY 8(d) no-reuse gather model.
    X, Y, Z = m.E, m.X0, m.X1
    for _ in range(3):
        ops.spmm(graph, X, Y, X, Z, 1.0)
    reps = 10 if args.workload == "synthetic" else 50
    a, b = ev(), ev()
    a.record()
    for _ in range(reps):
        ops.spmm(graph, X, Y, X, Z, 1.0)
    b.record()
    torch.cuda.synchronize()
    spmm_ms = a.elapsed_time(b) / reps
    # parity at full size, outside every timed region: one product against torch.sparse.mm accumulating in fp64
    spmm_check = None
    try:
        rows_chk = torch.randint(0, graph.n_rows, (200_000,), device=dev)
        Aref = torch.sparse_csr_tensor(graph.rowptr, graph.colidx.long(), graph.vals.double(), (graph.n_rows, graph.n_rows))
        ref = torch.sparse.mm(Aref, X.double())
        ops.spmm(graph, X, Y, X, Z, 0.25)
        e1 = float(((Y.double() - ref).abs().max() / ref.abs().max()).item())
        e2 = float(((Z.double() - 0.25 * (X.double() + ref)).abs().max() / ref.abs().max()).item())
        spmm_check = {"vs": "torch.sparse.mm (CSR, fp64 accumulate) on the same graph and input", "nnz": graph.nnz,
                      "max_abs_err_over_max_abs_Y": e1, "max_abs_err_over_max_abs_Z(fused mean epilogue)": e2, "ok": bool(e1 < 1e-5 and e2 < 1e-5)}
        del Aref, ref, rows_chk
        torch.cuda.empty_cache()
    except Exception as e:   # noqa: BLE001 -- informative only
        spmm_check = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
    pk, pk_src = peaks()
    alg = graph.algorithmic_bytes(D) + graph.n_rows * 4 * D * 2       # + read C, write Z of the fused epilogue
    achieved = alg / (spmm_ms * 1e-3) / 1e9
    ncu = ncu_spmm_traffic(args.workload)
    traffic = (ncu["dram_bytes_read"] + ncu["dram_bytes_write"]) if ncu else None

    # second kernel of the step: the fused full-rank evaluation (tcgen05, 3xTF32 = 3 tensor-core passes per score tile)
    fr_ms, roofline_eval = None, None
    if ops.EVAL_PRECISION == "tf32x3" and D <= 64:
        Ou, Oi = m.O[:U], m.O[U:]
        ops.fullrank_eval(Ou, Oi, users_all, rp, rc, target, 20)
        a, b = ev(), ev()
        a.record()
        for _ in range(3):
            ops.fullrank_eval(Ou, Oi, users_all, rp, rc, target, 20)
        b.record()
        torch.cuda.synchronize()
        fr_ms = a.elapsed_time(b) / 3
        issued = 3 * 2.0 * U * I * D                                   # tf32 MMA flops issued (hi*hi + hi*lo + lo*hi)
        tf32_peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"]) / 2      # tf32 runs at half the bf16 rate; measured bf16 peak / 2
        roofline_eval = {"bound": "tensor", "kernel": "fullrank_tc_kernel", "ms_per_launch": round(fr_ms, 3),
                         "achieved": round(issued / (fr_ms * 1e-3) / 1e12, 1), "peak": round(tf32_peak, 1), "unit": "TFLOP/s",
                         "frac": round(issued / (fr_ms * 1e-3) / 1e12 / tf32_peak, 4),
                         "useful_tflops": round(2.0 * U * I * D / (fr_ms * 1e-3) / 1e12, 1),
                         "model": "issued tf32 flops = 3 x 2 U I D (3xTF32 split keeps fp32-accurate scores); peak = measured "
                                  "sustained bf16 cuBLAS rate / 2 (tf32 : bf16 = 1 : 2 on the tcgen05 pipe); the kernel also masks "
                                  "train items, ranks the target and keeps a top-20 per user in its epilogue"}

    # end to end through the public API: host sampler + pinned H2D + epoch + loss D2H, then evaluation + metric D2H
    # (the next epoch's samples are drawn on a background thread while the GPU works: dataset._EpochPipe)
    data.config["prefetch"] = True
    data._pipe.prefetch = True
    e2e, metrics = [], None
    prime = 4                                          # untimed: fill the depth-2 prefetch queue (steady state of an epoch loop)
    for k in range(prime + max(1, args.steps)):
        torch.cuda.synchronize()
        t0 = time.time()
        loss_e2e = m.train_step()[0]
        sums, rank_ = full_eval()
        sums = sums.cpu().numpy()
        hr20 = float((rank_[:, 0] < 20).float().mean().item())
        torch.cuda.synchronize()
        if k >= prime:
            e2e.append(time.time() - t0)
        metrics = {"loss": loss_e2e, "recall@20": sums[0] / max(sums[2], 1), "ndcg@20": sums[1] / max(sums[2], 1), "HR@20(target 0)": hr20}
    data._pipe._flush()        # let the speculative epochs finish before their buffers go away
    h2d = sum(int(t.numel()) * t.element_size() for t in sets[0]) + (n * 4 if sets[0][0].dtype == torch.int32 else 0)   # users, rel, negs, perm (int32) or rows + perm (int64)
    d2h = 4 * 8 + 3 * 8 + 4

    sm_mhz = clocks.summary()["sm_mhz"]
    out = {
        "metric": METRIC, "value": round(step_ms / 1e3, 6), "unit": "s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(step_ms, 3), "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(args.workload, w, B),
                   "l2": "inputs exceed L2 (graph + tables > 126 MB)" if graph.nnz * 8 > 126e6 else "L2-resident workload; absolute times only",
                   "parallelism": "single GPU"},
        "epoch_s": round(float(np.mean(ep_ms)) / 1e3, 6), "eval_s": round(float(np.mean(evl_ms)) / 1e3, 6),
        "eval_users_per_s": round(U / (float(np.mean(evl_ms)) / 1e3), 1), "epoch_loss": loss,
        "graph_build_s": round(t_graph, 4), "edge_gen_s": round(t_gen, 3), "host_sampler_s": round(t_sampler, 3),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": round(achieved / pk["hbm_gbs"], 4),
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full (profiles/ncu_spmm_r02.md);
                     # null when csrc/spmm.cu changed after the capture
                     "traffic": traffic,
                     "frac_dram": round(traffic / (spmm_ms * 1e-3) / 1e9 / pk["hbm_gbs"], 4) if traffic else None,
                     "frac_l2": round(ncu["l2_to_sm_bytes"] / (spmm_ms * 1e-3) / (L2_CAP_BYTES_PER_CLK * sm_mhz * 1e6), 4)
                     if ncu and sm_mhz else None,
                     "kernel": f"spmm_kernel<{D},4,4,0>",
                     "ms_per_launch": round(spmm_ms, 4), "algorithmic_bytes": alg, "peak_source": pk_src,
                     "model": "no-reuse gather: nnz*(8+4D) + 3*N*4D + (N+1)*4 (SURVEY 8d); the gathers are served by L2, so "
                              "`frac` above 1 is NOT an HBM fraction: `frac_dram` = measured DRAM bytes / time / peak, `frac_l2` = "
                              "measured L2->SM bytes / time / (6300 B/clk x SM clock), the bound this kernel runs at "
                              "(profiles/ncu_spmm_r02.md)"},
        "e2e": {"value": round(float(np.mean(e2e)), 6), "unit": "s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "includes": "steady state of an epoch loop (4 untimed epochs fill the prefetch queue): train_step(): exact C++ MT19937 "
                            "sampler + shuffle on host (the next two epochs are drawn on background threads), pinned H2D of "
                            "samples + permutation, epoch, loss D2H; then full-rank eval of all users, Recall/NDCG/HR D2H",
                "per_step_s": [round(t, 3) for t in e2e], "metrics_last_step": metrics},
        "gpu_launches": args.steps * (n_batches * (2 * L + 4) + 3 + 4),
        "clocks": clocks.summary(),
    }
    if roofline_eval:
        out["roofline_eval"] = roofline_eval
    out["spmm_check_full_size"] = spmm_check
    if not args.no_gpu_baseline:
        try:
            out["gpu_library_baseline"] = torch_gpu_reference(graph, U, I, D, L, B, n_batches, dev)
        except Exception as e:   # noqa: BLE001 -- informative only, never fail the bench line
            out["gpu_library_baseline"] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        torch.cuda.empty_cache()
    if not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        out["cpu_baseline"] = (reference_lightgcn(w, B, None, 1, csr=graph.to_numpy()) or cpu_reference(w, B, graph, m, data, steps=1))
    return out


def torch_gpu_reference(graph, U, I, D, L, B, n_batches, dev):
    """The reference's own formulation (lightgcn.py:82-172: coalesced COO + torch.sparse.mm + autograd +
    torch.optim.Adam; evaluation as batched getUsersRating + topk, lightgcn.py:115-120) written with stock
    torch ops and run on the SAME B200 -- the GPU-library baseline of SURVEY.md 8(d).  Two training steps are
    timed and extrapolated to the epoch; evaluation is timed on 16384 users and extrapolated."""
    ev = lambda: torch.cuda.Event(enable_timing=True)     # noqa: E731
    N = U + I
    A = graph.to_torch_sparse()
    gen = torch.Generator(device=dev).manual_seed(0)
    E = torch.nn.Parameter(torch.randn(N, D, device=dev, generator=gen) * 0.1)
    opt = torch.optim.Adam([E], lr=1e-3)
    us, ps, ns = (torch.randint(0, hi, (B,), device=dev, generator=gen) for hi in (U, I, I))

    def computer():
        X, acc = E, E
        for _ in range(L):
            X = torch.sparse.mm(A, X)
            acc = acc + X
        return acc / (L + 1)

    def step():
        O = computer()
        u, p, q = O[us], O[U + ps], O[U + ns]
        u0, p0, q0 = E[us], E[U + ps], E[U + ns]
        loss = torch.nn.functional.softplus((u * q).sum(1) - (u * p).sum(1)).mean() + 1e-4 * 0.5 * (
            u0.norm(2).pow(2) + p0.norm(2).pow(2) + q0.norm(2).pow(2)) / B
        opt.zero_grad()
        loss.backward()
        opt.step()

    step()
    a, b, c, d = ev(), ev(), ev(), ev()
    a.record()
    for _ in range(2):
        step()
    b.record()
    with torch.no_grad():
        X = E.detach()
        torch.sparse.mm(A, X)
        c.record()
        for _ in range(3):
            torch.sparse.mm(A, X)
        d.record()
        O = computer()
        ne, chunk = min(16384, U), 4096
        e0, e1 = ev(), ev()
        e0.record()
        for lo in range(0, ne, chunk):
            torch.topk(O[lo:lo + chunk] @ O[U:].t(), 20)
        e1.record()
    torch.cuda.synchronize()
    step_s, spmm_ms, eval_user = a.elapsed_time(b) / 2e3, c.elapsed_time(d) / 3, e0.elapsed_time(e1) / 1e3 / ne
    epoch_s, eval_s = step_s * n_batches, eval_user * U
    return {"value": round(epoch_s + eval_s, 4), "unit": "s", "kind": "port (stock torch ops on the same GPU)",
            "epoch_s": round(epoch_s, 4), "eval_s": round(eval_s, 4), "step_s": round(step_s, 5), "spmm_ms": round(spmm_ms, 3),
            "sample": f"2 steps of batch {B} x {n_batches} batches (torch.sparse.mm COO fwd + autograd bwd + torch.optim.Adam); "
                      f"eval: {ne} users, matmul + topk without train-item masking, x {U} users"}


# ----------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU algorithm (oracle port: torch CPU restatement of
# recad/model/victim/lightgcn.py) on the host cores, bounded sample, extrapolated to the metric's unit
# ----------------------------------------------------------------------------------------------
def cpu_reference(w, B, graph=None, model_=None, data=None, steps=1, host_edges=None):
    from oracle import lightgcn as olg
    U, I, D, L = w["n_users"], w["n_items"], w["D"], w["L"]
    N = U + I
    if graph is not None:
        ptr, col, val = graph.to_numpy()
    else:
        from oracle import graph as og
        ptr, col, val, _, _ = og.norm_adj_csr(host_edges[0], host_edges[1], U, I)
    A = olg.csr_to_torch_coo(ptr, col, val, N)      # the reference holds a coalesced COO tensor (implicit.py:295-296)
    nnz = len(col)
    threads = torch.get_num_threads()
    g = torch.Generator().manual_seed(0)
    E = torch.randn(N, D, generator=g) * 0.1
    n_samples = min(w["n_edges"], int(nnz // 2))
    n_batches = (n_samples + B - 1) // B
    t_spmm, t_bpr, t_eval = [], [], []
    for _ in range(steps):
        t0 = time.time()
        X = torch.sparse.mm(A, E)                                  # lightgcn.py:99-108, one of the 2L per batch
        t_spmm.append(time.time() - t0)
        # one batch of the BPR head + dense Adam without the propagate (lightgcn.py:122-168)
        Bs = min(B, n_samples)
        us, ps, ns = (torch.randint(0, hi, (Bs,), generator=g) for hi in (U, I, I))
        P = torch.nn.Parameter(X.clone())
        opt = torch.optim.Adam([P], lr=1e-3)
        t0 = time.time()
        u, p, q = P[us], P[U + ps], P[U + ns]
        loss = torch.nn.functional.softplus((u * q).sum(1) - (u * p).sum(1)).mean() + 1e-4 * 0.5 * (
            u.norm(2).pow(2) + p.norm(2).pow(2) + q.norm(2).pow(2)) / Bs
        opt.zero_grad()
        loss.backward()
        opt.step()
        t_bpr.append(time.time() - t0)
        # evaluation sample: batched getUsersRating + topk (lightgcn.py:115-120) for 1000 users
        ne = min(1000, U)
        t0 = time.time()
        torch.topk(X[:ne] @ X[U:].t(), 20)
        t_eval.append((time.time() - t0) / ne)
    spmm, bpr, ev_user = float(np.mean(t_spmm)), float(np.mean(t_bpr)), float(np.mean(t_eval))
    epoch_s = n_batches * (2 * L * spmm + bpr)
    eval_s = ev_user * U
    return {"value": round(epoch_s + eval_s, 3), "unit": "s", "cores": threads, "kind": "port",
            "sample": f"full-size graph (nnz {nnz}): {steps} x [1 torch.sparse.mm of the 2L per batch ({spmm:.3f} s), 1 BPR-head+Adam batch "
                      f"of {min(B, n_samples)} ({bpr:.3f} s), batched matmul+topk eval of 1000 users ({ev_user * 1e3:.3f} ms/user)]; "
                      f"extrapolated: {n_batches} batches x (2L x spmm + bpr) + n_users x eval; the reference's own per-user eval "
                      f"(normal.py:62-71: one propagate per user = {L * spmm:.2f} s/user) is NOT charged",
            "epoch_s": round(epoch_s, 3), "eval_s": round(eval_s, 3), "spmm_s": round(spmm, 4)}


def workload_string(name, w, B):
    U, I, D, L = w["n_users"], w["n_items"], w["D"], w["L"]
    n = w["n_edges"]            # the pairwise sampler draws one sample per interaction (implicit.py:56-57)
    return (f"{name}: LightGCN {U} users x {I} items x {w['n_edges']} interactions, D={D}, L={L}, "
            f"BPR batch {B} ({(n + B - 1) // B} batches/epoch, {n} samples), full-rank eval of all {U} users K=20")


class _StubDataset:
    """The slice of the dataset interface the reference's LightGCN touches (lightgcn.py:32-34, 136): shapes, the
    normalised adjacency as a coalesced torch sparse COO tensor (implicit.py:295-296) and pre-sampled batches.  The
    reference's own dataset code cannot build a 50 M-interaction graph (Python dok / lil + per-user loops)."""
    dataset_name = "stub"

    def __init__(self, U, I, graph, batches):
        self.n_users, self.n_items, self.graph, self.batches = U, I, graph, batches

    def info_describe(self):
        return {"n_users": self.n_users, "n_items": self.n_items, "graph": self.graph}

    def generate_batch(self, **kw):
        yield from self.batches

    def mode(self):
        return "train"

    def switch_mode(self, mode):
        pass


def reference_lightgcn(w, B, host_edges, steps, csr=None):
    """The reference's OWN model class (recad.model.victim.LightGCN from baseline/_ref, unmodified) on the host cores:
    `train_step()` over a bounded number of batches and `getUsersRating` + topk over a bounded number of users,
    extrapolated to one epoch + one evaluation of every user."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "recad")):
        return None
    cwd = os.getcwd()
    import tempfile
    os.chdir(tempfile.mkdtemp())                     # the reference creates ./generated relative to cwd at import
    sys.path.insert(0, ref)
    try:
        import recad
        from oracle import graph as og
        from oracle import lightgcn as olg
        U, I, D, L = w["n_users"], w["n_items"], w["D"], w["L"]
        N = U + I
        if csr is not None:                           # the device-built matrix: bit-equal to the reference's (tests/test_gpu_graph_spmm.py)
            ptr, col, val = csr
        else:
            ptr, col, val, _, _ = og.norm_adj_csr(host_edges[0], host_edges[1], U, I)   # scipy fast path of implicit.py:243-298
        A = olg.csr_to_torch_coo(ptr, col, val, N)
        nnz = len(col)
        del ptr, col, val
        n = w["n_edges"]
        n_batches = (n + B - 1) // B
        g = torch.Generator().manual_seed(0)
        Bs = min(B, n)

        def batch():
            return {"users": torch.randint(0, U, (Bs,), generator=g), "positive_items": torch.randint(0, I, (Bs,), generator=g),
                    "negative_items": torch.randint(0, I, (Bs,), generator=g)}
        data = _StubDataset(U, I, A, [batch()])
        torch.manual_seed(2023)
        m = recad.model.from_config("victim", "lightgcn", latent_dim_rec=D, lightGCN_n_layers=L,
                                    device=torch.device("cpu")).I(dataset=data)
        assert type(m).__module__ == "recad.model.victim.lightgcn"
        t_step, t_comp, t_rate = [], [], []
        ne = min(2048, U)
        for _ in range(steps):
            data.batches = [batch()]
            t0 = time.time()
            m.train_step()                                           # lightgcn.py:132-172, ONE batch of the epoch
            t_step.append(time.time() - t0)
            with torch.no_grad():
                m.eval()
                t0 = time.time()
                users_emb, items_emb = m.computer()                  # what getUsersRating starts with (lightgcn.py:116)
                t_comp.append(time.time() - t0)
                t0 = time.time()
                torch.topk(m.f(torch.matmul(users_emb[:ne], items_emb.t())), 20)    # lightgcn.py:117-120 + top-20
                t_rate.append((time.time() - t0) / ne)
        step_s, comp_s, rate_s = float(np.mean(t_step)), float(np.mean(t_comp)), float(np.mean(t_rate))
        epoch_s, eval_s = n_batches * step_s, comp_s + rate_s * U
        return {"value": round(epoch_s + eval_s, 3), "unit": "s", "cores": torch.get_num_threads(), "kind": "reference",
                "sample": f"unmodified recad.model.victim.LightGCN (baseline/_ref) on a stub dataset over the full-size graph (nnz {nnz}): "
                          f"{steps} x [train_step() over 1 of the {n_batches} batches of {Bs} ({step_s:.2f} s), computer() ({comp_s:.2f} s), "
                          f"getUsersRating-style matmul + sigmoid + topk of {ne} users ({rate_s * 1e3:.3f} ms/user)]; extrapolated: "
                          f"{n_batches} x train_step + computer + n_users x rating; the reference's own evaluation loop "
                          f"(normal.py:62-71: one computer() PER USER = {comp_s:.1f} s/user) is NOT charged",
                "epoch_s": round(epoch_s, 3), "eval_s": round(eval_s, 3), "batch_step_s": round(step_s, 3), "computer_s": round(comp_s, 3)}
    finally:
        sys.path.remove(ref)
        os.chdir(cwd)


def run_reference(args, w):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return None
    torch.set_num_threads(os.cpu_count() or 1)       # torchrun exports OMP_NUM_THREADS=1: the reference gets every host core
    B = args.batch or w["batch"]
    t0 = time.time()
    # same generator as the GPU arm, on the CPU torch device (seeded identically; streams differ by device)
    eu, ei = synth_edges(w, torch.device("cpu"))
    edges = (eu.numpy(), ei.numpy())
    t_gen = time.time() - t0
    steps = max(1, min(args.steps, 2))               # each step is ~1 min of host work at the synthetic size
    base = reference_lightgcn(w, B, edges, steps) or cpu_reference(w, B, host_edges=edges, steps=steps)
    return {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": "s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": base["value"] * 1e3, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(args.workload, w, B),
                   "l2": "inputs exceed L2 (graph + tables > 126 MB)" if w["n_edges"] * 16 > 126e6 else "L2-resident workload; absolute times only",
                   "parallelism": f"{base['cores']} host threads (torch CPU)"},
        "cpu_baseline": base, "edge_gen_s": round(t_gen, 2), "timed_steps": steps,
        "e2e": {"value": base["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the stock-torch-on-the-same-GPU comparison")
    ap.add_argument("--workload", default="synthetic", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    out = run_reference(args, w) if args.impl == "reference" else run_b200(args, w)
    if out is not None and int(os.environ.get("RANK", 0)) == 0:
        print(json.dumps(out), flush=True)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
