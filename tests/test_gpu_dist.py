"""Multi-GPU parity (needs >= 2 GPUs; run with `gpurun --gpus 2`): the user-sharded NCCL path reproduces the
single-GPU epoch (loss, tables) and evaluation up to fp32 summation order."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from recad_b200 import dist as rdist, ops, synthetic
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    U, I, E, D, L, B = 3001, 1200, 60_000, 64, 3, 8192
    u, i = synthetic.make_edges(U, I, E, seed=3)
    eu, ei = torch.as_tensor(u, device=dev), torch.as_tensor(i, device=dev)
    g = torch.Generator().manual_seed(1)
    init_u, init_i = torch.randn(U, D, generator=g) * 0.1, torch.randn(I, D, generator=g) * 0.1
    n = 30_000
    samples = torch.stack([torch.randint(0, U, (n,), generator=g), torch.randint(0, I, (n,), generator=g),
                           torch.randint(0, I, (n,), generator=g)], 1).to(dev)
    perm = torch.randperm(n, generator=g).to(dev)
    runs = {}
    for fused in (False, True):       # NCCL all-reduce of the item block / partial rows sent to their owner from the SpMM epilogue
        m = rdist.ShardedLightGCN(U, I, (eu, ei), D=D, n_layers=L, batch=B, device=dev, init_user=init_u.to(dev),
                                  init_item=init_i.to(dev), fused=fused)
        assert (m.peer is not None) == fused
        losses = [m.train_epoch(samples, perm) for _ in range(2)]
        eu_all, ei_all = m.gather_tables()
        topi, topv, rank_, score = m.full_rank([5], 20)
        assert (m.n_fused > 0) == fused
        runs[fused] = (losses, eu_all, ei_all, rank_, score)
    # equal shards (3000 users / 2): the owner's slice is stored with ONE NVSwitch multicast store when the box has NVLS
    U2 = 3000
    keep = eu < U2
    s3 = samples.clone()
    s3[:, 0] %= U2
    mc = {}
    for fused in (False, True):
        m3 = rdist.ShardedLightGCN(U2, I, (eu[keep], ei[keep]), D=D, n_layers=L, batch=B, device=dev,
                                   init_user=init_u[:U2].to(dev), init_item=init_i.to(dev), fused=fused)
        l3 = m3.train_epoch(s3, perm)
        mc[fused] = (l3, *m3.gather_tables(), bool(m3.peer.multicast) if fused else None)
    ok_mc = bool(np.isclose(mc[True][0], mc[False][0], rtol=1e-5)) and all(
        torch.allclose(a, b, rtol=1e-4, atol=2e-6) for a, b in zip(mc[True][1:3], mc[False][1:3]))
    used_multicast = mc[True][3]
    # the host sampler's 32-bit arrays (user, index of the positive inside the user's row, negative) + int32 permutation
    # give the same epoch as the equivalent (user, pos, neg) rows
    keys = torch.unique(eu * I + ei)
    ap_ptr = torch.zeros(U + 1, dtype=torch.int64, device=dev)
    ap_ptr[1:] = torch.cumsum(torch.bincount(keys // I, minlength=U), 0)
    ap_col = keys % I
    su = samples[:, 0]
    deg = ap_ptr[su + 1] - ap_ptr[su]
    rel = (torch.randint(0, 1 << 30, (n,), generator=g).to(dev) % deg.clamp(min=1))
    rows_eq = torch.stack([su, ap_col[ap_ptr[su] + rel], samples[:, 2]], 1).contiguous()
    soa = {}
    for kind in ("soa", "rows"):
        m4 = rdist.ShardedLightGCN(U, I, (eu, ei), D=D, n_layers=L, batch=B, device=dev, init_user=init_u.to(dev),
                                   init_item=init_i.to(dev), fused=True)
        tr = {}
        if kind == "soa":
            l4 = m4.train_epoch_soa(su.int(), rel.int(), samples[:, 2].int().contiguous(), perm.int(), trace=tr)
            assert set(tr) >= {"bpr", "adam", "barriers"} and all(v >= 0 for v in tr.values())
        else:
            l4 = m4.train_epoch(rows_eq, perm)
        soa[kind] = (l4, *m4.gather_tables())
    ok_soa = bool(np.isclose(soa["soa"][0], soa["rows"][0], rtol=1e-6)) and all(
        torch.allclose(a, b, rtol=1e-5, atol=1e-7) for a, b in zip(soa["soa"][1:], soa["rows"][1:]))
    # attacked model: 7 fake users appended to the last shard, fresh tables, one epoch that also samples them
    F = 7
    fake = (torch.rand(F, I, generator=g) < 0.03).float() * 5.0          # rating 5 on ~3 % of the items
    fake[:, 5] = 5.0
    init_u2, init_i2 = torch.randn(U + F, D, generator=g) * 0.1, torch.randn(I, D, generator=g) * 0.1
    samples2 = torch.stack([torch.randint(0, U + F, (n,), generator=g), torch.randint(0, I, (n,), generator=g),
                            torch.randint(0, I, (n,), generator=g)], 1).to(dev)
    from recad_b200.dataset import ArrayImplicitData
    frp, fit = ArrayImplicitData.fake_rows(fake.numpy(), 4)
    m2 = m.inject(frp, fit, init_user=init_u2.to(dev), init_item=init_i2.to(dev))
    loss2 = m2.train_epoch(samples2, perm)
    eu2, ei2 = m2.gather_tables()
    if rank == 0:
        # single-GPU run of the same thing
        from recad_b200 import dataset, model
        data = dataset.ArrayImplicitData("t", U, I, (eu, ei), dev, batch_size=B, prefetch=False)
        ref = model.from_config("victim", "lightgcn", latent_dim_rec=D, lightGCN_n_layers=L, device=dev).I(dataset=data)
        ref.embedding_user.weight.data.copy_(init_u)
        ref.embedding_item.weight.data.copy_(init_i)
        data.epoch_samples = lambda device=None: (samples, perm)
        ref_losses = [ref.train_step()[0] for _ in range(2)]
        rp, rc = data.train_csr(dev)
        rtopi, rtopv, rrank, rscore, _ = ref.full_rank(torch.arange(m.Ug, device=dev), [5], 20, rp, rc)
        ok = True
        for fused, (losses, eu_all, ei_all, rank_, score) in runs.items():
            ok &= np.allclose(losses, ref_losses, rtol=1e-5)
            ok &= torch.allclose(eu_all, ref.embedding_user.weight, rtol=1e-4, atol=2e-6)
            ok &= torch.allclose(ei_all, ref.embedding_item.weight, rtol=1e-4, atol=2e-6)
            ok &= torch.allclose(score, rscore, rtol=1e-4, atol=1e-6)
            ok &= float((rank_ != rrank).float().mean()) < 0.01
        data2 = data.inject_data("explicit", fake.numpy(), filter_num=4)
        assert data2.n_users == U + F
        ref2 = model.from_config("victim", "lightgcn", latent_dim_rec=D, lightGCN_n_layers=L, device=dev).I(dataset=data2)
        ref2.embedding_user.weight.data.copy_(init_u2)
        ref2.embedding_item.weight.data.copy_(init_i2)
        data2.epoch_samples = lambda device=None: (samples2, perm)
        ref_loss2 = ref2.train_step()[0]
        ok2 = np.allclose(loss2, ref_loss2, rtol=1e-5)
        ok2 &= torch.allclose(eu2, ref2.embedding_user.weight, rtol=1e-4, atol=2e-6)
        ok2 &= torch.allclose(ei2, ref2.embedding_item.weight, rtol=1e-4, atol=2e-6)
        q.put((bool(ok) and bool(ok2) and ok_mc and ok_soa, losses + [loss2, ("multicast", used_multicast, ok_mc), ("soa", ok_soa)],
               ref_losses + [ref_loss2]))
    dist.barrier()
    dist.destroy_process_group()


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    from recad_b200 import dataset, dist as rdist, model, synthetic
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    U, I, B, n = 1500, 700, 1000, 7300                                   # last batch: 300 rows
    u, i = synthetic.make_edges(U, I, 20_000, seed=4)
    data = dataset.ArrayImplicitData("t", U, I, (torch.as_tensor(u, device=dev), torch.as_tensor(i, device=dev)), dev,
                                     sample="pointwise", batch_size=B, need_graph=False, prefetch=False)
    g = torch.Generator().manual_seed(2)
    samples = torch.stack([torch.randint(0, U, (n,), generator=g), torch.randint(0, I, (n,), generator=g),
                           torch.randint(0, 2, (n,), generator=g)], 1).to(dev)
    perm = torch.randperm(n, generator=g).to(dev)
    results = {}
    for name, kw in (("mf", dict(embedding_size=32)), ("ncf", dict(factor_num=8, num_layers=3, tower_precision="fp32"))):
        torch.manual_seed(11)
        v = model.from_config("victim", name, device=dev, **kw).I(dataset=data)
        dp = rdist.DataParallelVictim(v)
        losses = [dp.train_epoch(samples, perm, batch=B) for _ in range(2)]
        if rank == 0:
            torch.manual_seed(11)
            ref = model.from_config("victim", name, device=dev, **kw).I(dataset=data)
            data.epoch_samples = lambda device=None: (samples, perm)
            ref_losses = [ref.train_step()[0] for _ in range(2)]
            diff = (v.flat - ref.flat).abs()
            ok = diff <= 2e-6 + 1e-4 * ref.flat.abs()
            # NCF: the state of a ReLU unit sitting at zero depends on the order of the atomic gradient adds (and of
            # the all-reduce): nearly all elements at the bar, none further off than a few Adam steps
            close = bool(ok.all()) if name == "mf" else bool(ok.float().mean() >= 0.98 and diff.max() <= 3.5e-3)
            results[name] = (losses, ref_losses, close, float(diff.max()), v._steps == ref._steps)
    if rank == 0:
        q.put(results)
    dist.barrier()
    dist.destroy_process_group()


def _spawn(worker, world=2):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    return q.get()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_data_parallel_mf_ncf_match_single_gpu():
    """MF / NCF with the rows of every batch split over 2 GPUs and the flat gradient all-reduced: same
    epoch losses and tables as one GPU up to fp32 summation order (NCF with the exact fp32 tower so that no
    ReLU mask can flip)."""
    results = _spawn(_dp_worker)
    for name, (losses, ref_losses, close, max_abs, same_steps) in results.items():
        assert np.allclose(losses, ref_losses, rtol=1e-5), (name, losses, ref_losses)
        assert close, (name, max_abs)
        assert same_steps


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_sharded_epoch_matches_single_gpu():
    ok, losses, ref_losses = _spawn(_worker)
    print("sharded run:", losses[-1])
    assert ok, (losses, ref_losses)
