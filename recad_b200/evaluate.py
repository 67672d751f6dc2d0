"""Full-ranking evaluation, batched on the device (drop-in for
recad/workflow/normal.py:57-93 `user_item_model_generate` and 111-160 `normal_evaluate`).

The reference scores one user at a time (one full LightGCN propagation per user, normal.py:62-71),
copies every score to the host and ranks with a pandas sort per user.  Here the eligible users go
through ONE fused kernel (score tile x train mask x target rank x top-K) per model and only
[n_users, T] ranks and scores come back.

Tie rule (the reference's pandas quicksort leaves ties undefined, normal.py:86-88): an item
outranks the target iff score > s_t, or score == s_t and item id < target id.
"""
from collections import OrderedDict

import numpy as np
import torch


def eligible_users(dataset, target_id_list):
    """normal.py:133-143: train users (non-empty train list) with NO target item in their train set
    and at least one candidate item, ascending (pandas groupby order, normal.py:79)."""
    ptr, col = dataset.train_csr()
    n_items = dataset.n_items
    has = np.diff(ptr) > 0
    rows = np.repeat(np.arange(len(ptr) - 1), np.diff(ptr))
    hit = np.zeros(len(ptr) - 1, dtype=bool)
    hit[rows[np.isin(col, np.asarray(list(target_id_list), dtype=col.dtype))]] = True
    full = np.diff(ptr) >= n_items                      # no candidate left: skipped (normal.py:63-64)
    train_dict = dataset.info_describe().get("train_dict")
    if train_dict is not None:                          # every KEY of train_dict is evaluated, even with an empty list
        keys = np.fromiter((int(k) for k in train_dict), dtype=np.int64, count=len(train_dict))
        has = np.zeros(len(ptr) - 1, dtype=bool)
        has[keys] = True
    return np.flatnonzero(has & ~hit & ~full).astype(np.int64)


def model_rows(model, dataset, target_id_list, topks, K=20, users=None):
    """== user_item_model_generate: float64 [n_users * T, 2 + len(topks)] rows
    [uid, score_target, 1[target in top-k] ...], INCLUDING the reference's row-index quirk for
    more than one target (normal.py:92: row = user index + target index, later users overwrite).
    Also returns the device outputs of the fused kernel."""
    dev = model._dev
    users = eligible_users(dataset, target_id_list) if users is None else np.asarray(users, dtype=np.int64)
    T = len(target_id_list)
    n = len(users)
    rows = np.zeros((n * T, 2 + len(topks)), dtype=np.float64)
    if n == 0:
        return rows, None
    rowptr, col = dataset.train_csr(dev)
    uid = torch.from_numpy(users).to(dev)
    topi, topv, trank, tscore, offset = model.full_rank(uid, target_id_list, K, rowptr, col)
    rank = trank.cpu().numpy().astype(np.int64)
    score = tscore.cpu().numpy().astype(np.float64) + offset
    # quirk-faithful fill: for idx ascending, bias ascending: rows[idx + bias] = line(idx, bias)
    final = {}
    for bias in range(T):
        # rows idx + bias for idx in [0, n): later idx overwrite earlier ones; process in write order
        lines = np.concatenate([users[:, None].astype(np.float64), score[:, bias:bias + 1],
                                np.stack([(rank[:, bias] < k) for k in topks], 1).astype(np.float64)], 1)
        final[bias] = lines
    if T == 1:
        rows[:] = final[0]
    else:
        for idx in range(n):                   # exact emulation of normal.py:80-92 write order (T is tiny)
            for bias in range(T):
                rows[idx + bias] = final[bias][idx]
    return rows, (uid, topi, topv, trank, tscore)


def normal_evaluate(model, model_fake, dataset, target_id_list, topks, verbose=True):
    """== Normal.normal_evaluate (normal.py:111-160): pred_shift and HR@k before / after attack.
    `dataset` is the CLEAN dataset (the fake users are not evaluated)."""
    for m in (model, model_fake):
        fwd = m.input_describe()["forward"]
        assert len(fwd) == 2, "Expect forward only need two inputs"
        assert "users" in fwd, "Expect to have the users input in forward method"
        assert "items" in fwd, "Expect to have the items input in forward method"
    users = eligible_users(dataset, target_id_list)
    pred_results, _ = model_rows(model, dataset, target_id_list, topks, users=users)
    pred_results_fake, _ = model_rows(model_fake, dataset, target_id_list, topks, users=users)
    assert np.allclose(pred_results[:, 0], pred_results_fake[:, 0]), "Users are not aligned"
    results = OrderedDict()
    results["pred_shift"] = np.mean(pred_results_fake[:, 1] - pred_results[:, 1])
    for i, k in enumerate(topks):
        results[f"HR@{k}"] = np.mean(pred_results[:, 2 + i])
        results[f"HR@{k} after attack"] = np.mean(pred_results_fake[:, 2 + i])
    if verbose:
        try:
            from tabulate import tabulate
            print(tabulate(list(results.items()), headers="firstrow", tablefmt="fancy_grid"))
        except ImportError:
            for k, v in results.items():
                print(f"{k}: {v}")
    return results


def recall_ndcg(model, dataset, K=20, split="test", users=None):
    """Recall@K / NDCG@K over the users of a held-out split, train items masked (the consumer the
    reference's test-mode batches and getUsersRating were written for: implicit.py:461-476,
    lightgcn.py:115-120).  PARITY UNPINNED in the reference; definition in
    oracle/evaluate.py.  Returns {'recall', 'ndcg', 'n_users'}."""
    from . import ops
    dev = model._dev
    gt_ptr, gt_col = dataset.ground_truth_csr(split, dev)
    if users is None:
        users = torch.nonzero(gt_ptr[1:] > gt_ptr[:-1]).flatten()
    else:
        users = torch.as_tensor(users, dtype=torch.int64, device=dev)
    rowptr, col = dataset.train_csr(dev)
    topi, _, _, _, _ = model.full_rank(users, [], K, rowptr, col)
    s = ops.recall_ndcg(topi, users, gt_ptr, gt_col).cpu().numpy()
    n = max(s[2], 1.0)
    return {"recall": s[0] / n, "ndcg": s[1] / n, "n_users": int(s[2])}
