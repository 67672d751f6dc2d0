"""World-size-2 gloo test (CPU) of the multi-GPU host logic: user partition, routing of the global batches to
the owner of each user, and the cross-rank loss / degree aggregation identities the sharded path relies on."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from recad_b200.dist import batch_rows, route_epoch, user_range


def _worker(rank, world, port, U, I, n, batch, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)               # identical "global" data on every rank
    samples = torch.stack([torch.randint(0, U, (n,), generator=g), torch.randint(0, I, (n,), generator=g),
                           torch.randint(0, I, (n,), generator=g)], 1)
    perm = torch.randperm(n, generator=g)
    lo, hi = user_range(U, rank, world)
    local, ptr = route_epoch(samples, perm, batch, lo, hi)
    S = samples[perm]
    n_batches = (n + batch - 1) // batch
    assert len(ptr) == n_batches + 1 and ptr[-1] == local.shape[0]
    for b in range(n_batches):                         # every batch: exactly the rows of my users, in order
        glob = S[b * batch:(b + 1) * batch]
        mine = glob[(glob[:, 0] >= lo) & (glob[:, 0] < hi)].clone()
        mine[:, 0] -= lo
        assert torch.equal(local[ptr[b]:ptr[b + 1]], mine)
    # counts over ranks add up to the global batch sizes (the 1/B normalisation uses the global size)
    counts = torch.tensor([ptr[b + 1] - ptr[b] for b in range(n_batches)])
    dist.all_reduce(counts)
    expect = torch.tensor([min(batch, n - b * batch) for b in range(n_batches)])
    assert torch.equal(counts, expect)
    # item degrees: local partial counts all-reduce to the global degree
    edges_u, edges_i = samples[:, 0], samples[:, 1]
    part = torch.bincount(edges_i[(edges_u >= lo) & (edges_u < hi)], minlength=I)
    dist.all_reduce(part)
    assert torch.equal(part, torch.bincount(edges_i, minlength=I))
    # partial (sum softplus, sum sq) per batch all-reduce to the single-process totals
    x = (S[:, 1] - S[:, 2]).double() / I
    sp_local = torch.zeros(n_batches, dtype=torch.float64)
    for b in range(n_batches):
        rows = local[ptr[b]:ptr[b + 1]]
        sp_local[b] = torch.nn.functional.softplus((rows[:, 1] - rows[:, 2]).double() / I).sum()
    dist.all_reduce(sp_local)
    ref = torch.stack([torch.nn.functional.softplus(x[b * batch:(b + 1) * batch]).sum() for b in range(n_batches)])
    assert torch.allclose(sp_local, ref, rtol=1e-12)
    out.put((rank, lo, hi, int(local.shape[0])))
    dist.destroy_process_group()


def test_batch_rows_partition_every_global_batch():
    """Data-parallel MF / NCF: the ranks' row sets of a batch are disjoint, balanced and cover it (so the
    all-reduced gradient sums every row exactly once); a short last batch may leave a rank empty."""
    perm = torch.randperm(1000, generator=torch.Generator().manual_seed(0))
    for B, world in ((256, 2), (256, 8), (3, 8), (1000, 3)):
        for b0 in range(0, 1000, B):
            Bg = min(B, 1000 - b0)
            parts = [batch_rows(perm, b0, Bg, r, world) for r in range(world)]
            assert sorted(torch.cat(parts).tolist()) == sorted(perm[b0:b0 + Bg].tolist())
            sizes = [p.numel() for p in parts]
            assert max(sizes) - min(sizes) <= 1 and all(p.is_contiguous() for p in parts)


def test_user_range_is_a_balanced_partition():
    for U, W in ((10, 3), (1_000_000, 8), (7, 8), (128, 2)):
        r = [user_range(U, k, W) for k in range(W)]
        assert r[0][0] == 0 and r[-1][1] == U and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1


def test_routing_and_aggregation_world_size_2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    U, I, n, batch = 1001, 300, 20_000, 4096
    procs = [ctx.Process(target=_worker, args=(r, 2, port, U, I, n, batch, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    got = sorted(out.get() for _ in range(2))
    assert got[0][1] == 0 and got[0][2] == got[1][1] and got[1][2] == U
    assert got[0][3] + got[1][3] == n
