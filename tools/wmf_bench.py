"""WMF surrogate retrain of the attack loop at the ml1m shape: one `fit_adv` call = 50 epochs of
batch-16 dense-Adam steps over a [(5950 + 50) x 3702] rating matrix, the last epoch unrolled, + the reverse pass to the
fake rows.  CUDA path (csrc/wmf.cu: persistent cluster kernels) next to the oracle's torch-CPU restatement of the
reference loop (bounded sample: 2 plain epochs + 1 unrolled epoch with backward, extrapolated).

    python tools/wmf_bench.py [--epochs 50] [--reps 3] [--no-cpu]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recad_b200 import surrogate  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=5950)
    ap.add_argument("--items", type=int, default=3702)
    ap.add_argument("--fake", type=int, default=50)
    ap.add_argument("--epochs", type=int, default=50)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    n_rows = a.users + a.fake
    data = ((rng.random((n_rows, a.items)) < 0.021) * rng.integers(1, 6, (n_rows, a.items))).astype(np.float32)   # ~78 ratings per user
    data[-a.fake:] = (rng.random((a.fake, a.items)) * 5 * (rng.random((a.fake, a.items)) < 0.03)).astype(np.float32)
    up = torch.tensor(rng.standard_normal((n_rows, a.items)).astype(np.float32), device=dev)
    tr = surrogate.WMFTrainer(n_users=n_rows, n_items=a.items, device=dev, hidden_dim=16, lr=1e-2, weight_decay=1e-5, batch_size=16,
                              weight_pos=1.0, weight_neg=0.0)
    d = torch.tensor(data, device=dev, requires_grad=True)
    times = []
    for r in range(a.reps + 1):
        torch.manual_seed(r); np.random.seed(r)
        d.grad = None
        torch.cuda.synchronize()
        t0 = time.time()
        pred = tr.fit_adv(d, a.epochs, 1)
        torch.cuda.synchronize()
        t1 = time.time()
        (pred * up).sum().backward()
        torch.cuda.synchronize()
        t2 = time.time()
        if r:
            times.append((t1 - t0, t2 - t1))
    spe = (n_rows + 15) // 16
    out = {"workload": f"WMF surrogate fit_adv: {n_rows} rows x {a.items} items, dim 16, batch 16, {a.epochs} epochs ({spe * a.epochs} Adam steps), 1 unrolled",
           "fit_s": round(float(np.mean([t[0] for t in times])), 4), "backward_s": round(float(np.mean([t[1] for t in times])), 4),
           "us_per_step": round(float(np.mean([t[0] for t in times])) / (spe * a.epochs) * 1e6, 2),
           "fake_row_grad_abs_max": float(d.grad[-a.fake:].abs().max())}
    if not a.no_cpu:
        from oracle import wmf as owmf
        torch.set_num_threads(os.cpu_count() or 1)
        dc = torch.tensor(data, requires_grad=True)
        torch.manual_seed(0); np.random.seed(0)
        t0 = time.time()
        owmf.fit_adv(dc, 2, 0)
        t1 = time.time()
        pred, _, _ = owmf.fit_adv(dc, 1, 1)
        (pred * up.cpu()).sum().backward()
        t2 = time.time()
        plain, unrolled = (t1 - t0) / 2, t2 - t1
        out["cpu_baseline"] = {"kind": "port", "cores": torch.get_num_threads(), "value": round(plain * (a.epochs - 1) + unrolled, 2), "unit": "s",
                               "sample": f"oracle/wmf.py (torch CPU): 2 plain epochs ({plain:.2f} s each) + 1 unrolled epoch with autograd "
                                         f"backward ({unrolled:.2f} s); extrapolated to {a.epochs - 1} plain + 1 unrolled"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
