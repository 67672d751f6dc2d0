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


def eligible_users(dataset, target_id_list, device=None):
    """normal.py:133-143: train users (non-empty train list) with NO target item in their train set
    and at least one candidate item, ascending (pandas groupby order, normal.py:79).
    With a CUDA `device` the set is built there (recad_eligible_users: flag -> scan -> compaction over the device copy
    of the train rows) and returned as an int64 device tensor; without, the same on the host as a numpy array."""
    train_dict = dataset.info_describe().get("train_dict")
    if device is not None and torch.device(device).type == "cuda":
        from . import ops
        ptr, col = dataset.train_csr(device)
        is_key = None
        if train_dict is not None:                      # every KEY of train_dict is evaluated, even with an empty list
            is_key = getattr(dataset, "_train_key_mask", None)
            if is_key is None or is_key.device != ptr.device or is_key.numel() != ptr.numel() - 1:
                mask = np.zeros(ptr.numel() - 1, dtype=np.uint8)
                mask[np.fromiter((int(k) for k in train_dict), dtype=np.int64, count=len(train_dict))] = 1
                is_key = torch.from_numpy(mask).to(ptr.device)
                try:
                    dataset._train_key_mask = is_key
                except AttributeError:
                    pass
        return ops.eligible_users(ptr, col, ptr.numel() - 1, dataset.n_items, [int(t) for t in target_id_list], is_key)
    ptr, col = dataset.train_csr()
    n_items = dataset.n_items
    has = np.diff(ptr) > 0
    rows = np.repeat(np.arange(len(ptr) - 1), np.diff(ptr))
    hit = np.zeros(len(ptr) - 1, dtype=bool)
    hit[rows[np.isin(col, np.asarray(list(target_id_list), dtype=col.dtype))]] = True
    full = np.diff(ptr) >= n_items                      # no candidate left: skipped (normal.py:63-64)
    if train_dict is not None:
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
    if users is None:
        users = eligible_users(dataset, target_id_list, device=dev)
    uid = users.to(dev) if torch.is_tensor(users) else torch.from_numpy(np.asarray(users, dtype=np.int64)).to(dev)
    users = uid.cpu().numpy()                          # the uid column of the result table (host by contract)
    T = len(target_id_list)
    n = len(users)
    rows = np.zeros((n * T, 2 + len(topks)), dtype=np.float64)
    if n == 0:
        return rows, None
    rowptr, col = dataset.train_csr(dev)
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
    users = eligible_users(dataset, target_id_list, device=model._dev)      # built once on the device, shared by both passes
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


def recall_ndcg_batches(model, dataset, K=20, split="test"):
    """The same metric driven by the reference's test-mode batch interface (implicit.py:461-476): the dataset is
    switched to `split`, every batch hands over `users`, their train items (`positive_items`, the mask) and
    `ground_truth`; each batch is one fused full-rank launch.  Restores the dataset's mode."""
    from . import ops
    dev = model._dev
    prev = dataset.mode()
    dataset.switch_mode("validate" if split in ("valid", "validate") else "test")
    sums = np.zeros(3, dtype=np.float64)
    try:
        rowptr, col = dataset.train_csr(dev)
        for batch in dataset.generate_batch():
            users = batch["users"].to(dev).long()
            if users.numel() == 0:
                continue
            gt = batch["ground_truth"]
            lens = np.fromiter((len(set(g)) for g in gt), dtype=np.int64, count=len(gt))
            gptr = np.zeros(int(users.max().item()) + 2, dtype=np.int64)
            uh = users.cpu().numpy()
            gptr_rows = np.zeros(len(gptr) - 1, dtype=np.int64)
            gptr_rows[uh] = lens
            np.cumsum(gptr_rows, out=gptr[1:])
            gcol = np.zeros(int(lens.sum()), dtype=np.int32)
            for u, g in zip(uh, gt):
                g = np.unique(np.asarray(g, dtype=np.int32))
                gcol[gptr[u]:gptr[u] + len(g)] = g
            topi, _, _, _, _ = model.full_rank(users, [], K, rowptr, col)
            sums += ops.recall_ndcg(topi, users, torch.from_numpy(gptr).to(dev), torch.from_numpy(gcol).to(dev)).cpu().numpy()
    finally:
        dataset.switch_mode(prev)
    n = max(sums[2], 1.0)
    return {"recall": sums[0] / n, "ndcg": sums[1] / n, "n_users": int(sums[2])}


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
