"""AUSH attacker training at the ml1m shape (BASELINE config 2's attacker phase): `train_step` =
one pass over the eligible users in batches of 256 (aush.py:78-180), 5 950 users x 3 702 items, ~79 ratings per user,
filler_num 36, one selected item.  CUDA path (recad_b200/attacker.py + csrc/aush.cu) with its host / device split, next to
 * the UNMODIFIED reference attacker on the reference's own ExplicitData (baseline/_ref, if present) on the host cores and
   on the same GPU through its own torch code (the GPU-library baseline), and
 * the oracle's numpy restatement (bounded sample: 1 epoch).

    python tools/aush_bench.py [--epochs 10] [--no-ref]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from recad_b200 import dataset, model  # noqa: E402


def timed(fn, n):
    ts = []
    for _ in range(n):
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return ts, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=5950)
    ap.add_argument("--items", type=int, default=3702)
    ap.add_argument("--epochs", type=int, default=10)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--no-ref", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    mat = ((rng.random((a.users, a.items)) < 0.0213) * rng.integers(1, 6, (a.users, a.items))).astype(np.float32)
    mat[-1, -1] = 3                                   # pins n_users / n_items of the datasets built from the rows
    targets = [0]
    uu, ii = np.nonzero(mat)
    kvr = np.stack([uu, ii, mat[uu, ii]], 1).astype(np.float64)
    torch.manual_seed(2023); np.random.seed(2023)
    ds = dataset.from_config("explicit", "ml1m", device=dev, batch_size=a.batch, train_dict=kvr.copy(), valid_dict=kvr[:8].copy(),
                             test_dict=kvr[:8].copy())
    assert np.array_equal(ds.train_mat, mat)
    att = model.from_config("attacker", "aush", device=dev).I(dataset=ds)
    G0, D0 = att.netG_state(), att.netD_state()
    t_first, _ = timed(lambda: att.train_step(target_id_list=targets), 1)         # includes the one-off candidate lists
    ts, loss = timed(lambda: att.train_step(target_id_list=targets), a.epochs)
    n_elig = len(att._eligible(ds.train_mat, targets))
    nb = (n_elig + a.batch - 1) // a.batch
    # device share: the epoch call alone, CUDA events on the launching stream
    import ctypes as C
    from recad_b200 import _lib, ops
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    orig = _lib.lib().recad_aush_train_epoch
    dev_ms = []

    def wrapped(*args):
        ev0.record(torch.cuda.current_stream(dev))
        rc = orig(*args)
        ev1.record(torch.cuda.current_stream(dev))
        torch.cuda.synchronize()
        dev_ms.append(ev0.elapsed_time(ev1))
        return rc
    _lib.lib().recad_aush_train_epoch = wrapped
    for _ in range(3):
        att.train_step(target_id_list=targets)
    _lib.lib().recad_aush_train_epoch = orig
    tg, fake = timed(lambda: att.generate_fake(target_id_list=targets), 3)
    out = {"workload": f"AUSH train_step: {a.users} users x {a.items} items, {n_elig} eligible rows, batch {a.batch} ({nb} batches), "
                       f"filler_num 36, 1 selected item",
           "epoch_s": round(float(np.median(ts)), 5), "first_epoch_s": round(t_first[0], 4), "device_ms_per_epoch": round(float(np.median(dev_ms)), 3),
           "device_us_per_batch": round(float(np.median(dev_ms)) / nb * 1e3, 2), "gpu_launches_per_epoch": 5 * nb + 1,
           "generate_fake_s": round(float(np.median(tg)), 5), "loss": [round(x, 6) for x in loss], "fake_sum": float(fake.sum())}
    if not a.no_ref:
        ref = os.path.join(ROOT, "baseline", "_ref")
        if os.path.isdir(os.path.join(ref, "recad")):
            sys.path.insert(0, ref)
            import recad
            from recad.model.attacker.aush import Aush as RefAush
            torch.set_num_threads(os.cpu_count() or 1)
            os.chdir(__import__("tempfile").mkdtemp())
            for name, d in (("reference_cpu", torch.device("cpu")), ("reference_torch_cuda", dev)):
                dsr = recad.dataset.from_config("explicit", "ml1m", device=d, download=False, batch_size=a.batch, train_dict=kvr.copy(),
                                                valid_dict=kvr[:8].copy(), test_dict=kvr[:8].copy())
                assert type(dsr).__module__.startswith("recad.")
                torch.manual_seed(2023); np.random.seed(2023)
                r = recad.model.from_config("attacker", "aush", device=d).I(dataset=dsr)
                assert isinstance(r, RefAush) or type(r).__module__.startswith("recad.")
                r.netG.load_state_dict({k: v.to(d) for k, v in G0.items()})
                r.netD.load_state_dict({k: v.to(d) for k, v in D0.items()})
                r.train_step(target_id_list=targets)
                tr, lr = timed(lambda: r.train_step(target_id_list=targets), 2)
                out[name + "_epoch_s"] = round(float(np.median(tr)), 4)
                out[name + "_loss"] = [round(float(x), 6) for x in lr]
            out["host_threads"] = torch.get_num_threads()
        from oracle import aush as oa
        o = oa.AushOracle(mat, {k: v.cpu().numpy() for k, v in G0.items()}, {k: v.cpu().numpy() for k, v in D0.items()}, batch_size=a.batch)
        np.random.seed(2023)
        t0 = time.perf_counter()
        lo = o.train_step(targets)
        out["oracle_numpy_epoch_s"] = round(time.perf_counter() - t0, 3)
        # same seed, same start: the first CUDA epoch must reproduce the oracle's first epoch
        torch.manual_seed(2023); np.random.seed(2023)
        att2 = model.from_config("attacker", "aush", device=dev).I(dataset=ds)
        att2.load_netG_state(G0); att2.load_netD_state(D0)
        np.random.seed(2023)
        l2 = att2.train_step(target_id_list=targets)
        out["first_epoch_rel_err_vs_oracle"] = [float(abs(x - y) / abs(y)) for x, y in zip(l2, lo)]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
