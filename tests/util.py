"""Shared fixture loaders for the test-suite."""
import hashlib
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def meta():
    with open(os.path.join(GOLDEN, "meta.json")) as f:
        return json.load(f)


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def dicts(name):
    """(train, valid, test) dicts of a stored dataset, in the reference's dict order."""
    z = load(f"{name}_dicts.npz")
    out = []
    for split in ("train", "valid", "test"):
        keys, indptr, items = (z[f"{split}_{k}"].astype(np.int64) for k in ("keys", "indptr", "items"))
        out.append({int(k): items[indptr[j]:indptr[j + 1]].tolist() for j, k in enumerate(keys)})
    return tuple(out)


def split_batches(z, names):
    """Re-cut the concatenated recorded batches of a fixture into per-batch tuples."""
    sizes = z["batch_sizes"]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    return [tuple(z[n][offs[k]:offs[k + 1]] for n in names) for k in range(len(sizes))]


def aush_case(case):
    """(train_mat, G0, D0, hyper-parameters, golden archive) of tests/golden/aush_synth.npz (make_golden_aush.py)."""
    z = load("aush_synth.npz")
    hp = z[f"{case}_hp"]
    tr = z["train"].astype(np.int64)
    mat = np.zeros((int(hp[0]), int(hp[1])), dtype=np.float32)
    mat[tr[:, 0], tr[:, 1]] = tr[:, 2]
    G = {k[len(case) + 5:]: z[k] for k in z.files if k.startswith(f"{case}_G0__")}
    D = {k[len(case) + 5:]: z[k] for k in z.files if k.startswith(f"{case}_D0__")}
    kw = dict(selected_ids=z[f"{case}_selected_ids"].tolist(), filler_num=int(hp[3]), attack_num=int(hp[4]), ZR_ratio=float(hp[5]))
    return mat, G, D, kw, int(hp[2]), z[f"{case}_targets"].tolist(), z
