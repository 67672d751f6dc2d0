"""Attack-loop proxy (SURVEY.md 8d input 5): ITERS x (inject 50 fake users -> 1 epoch on the attacked graph ->
full-rank evaluation of every user) on the synthetic 1M x 200k x 50M graph -- the retrain-and-evaluate cycle of
Normal.execute (recad/workflow/normal.py:193-225) with the device-resident injection instead of the reference's
full dataset rebuild (implicit.py:482-494, base.py:108-118).

    python tools/attack_loop.py [--iters 10] [--workload synthetic]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/attack_loop.py    # sharded

Prints one JSON line (rank 0) with the per-phase wall-clock means (device work bracketed by synchronize).
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
import bench  # noqa: E402


def fake_profiles(rng, n_fake, n_items, target, fillers=36):
    """Random-attack shaped profiles: `fillers` random items + the target, all rated 5 (explicit ratings)."""
    fake = np.zeros((n_fake, n_items), dtype=np.float32)
    for r in range(n_fake):
        fake[r, rng.choice(n_items, size=fillers, replace=False)] = 5.0
        fake[r, target] = 5.0
    return fake


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--workload", default="synthetic")
    ap.add_argument("--fake", type=int, default=50)
    ap.add_argument("--device-init", action="store_true", help="fresh models drawn on the GPU instead of replaying the CPU init stream")
    args = ap.parse_args()
    w = bench.WORKLOADS[args.workload]
    U, I, D, L, B = w["n_users"], w["n_items"], w["D"], w["L"], w["batch"]
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from recad_b200 import dataset, model, ops
    target = [0]
    rng = np.random.default_rng(7)
    np.random.seed(2023)
    eu, ei = bench.synth_edges(w, dev)
    sync = torch.cuda.synchronize
    phases = {"inject_s": [], "sample_s": [], "epoch_s": [], "eval_s": []}
    hrs = []

    if world == 1:
        clean = dataset.ArrayImplicitData(args.workload, U, I, (eu, ei), dev, batch_size=B, prefetch=False)
        ops.pairwise_filter(*clean._allpos, U)                           # the clean dataset's sampler filter, built once
        del eu, ei
        for it in range(args.iters):
            fake = fake_profiles(rng, args.fake, I, target[0])
            sync(); t0 = time.time()
            data = clean.inject_data("explicit", fake, filter_num=4)
            victim = model.from_config("victim", "lightgcn", latent_dim_rec=D, lightGCN_n_layers=L, device=dev,
                                       init_on_device=args.device_init).I(dataset=data)
            sync(); t1 = time.time()
            samples = data.epoch_samples(dev)
            sync(); t2 = time.time()
            data.epoch_samples = lambda device=None, s=samples: s
            loss = victim.train_step()[0]
            sync(); t3 = time.time()
            rp, rc = data.train_csr(dev)
            users = torch.arange(U, device=dev)                       # the genuine users (normal.py:133-143)
            _, _, rank_, _, _ = victim.full_rank(users, target, 20, rp, rc)
            hr = float((rank_[:, 0] < 20).float().mean().item())
            sync(); t4 = time.time()
            for k, v in zip(phases, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
                phases[k].append(v)
            hrs.append(hr)
            del data, victim, samples
    else:
        import torch.distributed as dist
        from recad_b200 import dist as rdist
        dist.init_process_group("nccl", device_id=dev)
        base = rdist.ShardedLightGCN(U, I, (eu, ei), D=D, n_layers=L, batch=B, device=dev)
        n = int(eu.numel())
        if rank == 0:                                                 # the sampler needs every user's positives
            keys = torch.unique(eu * I + ei)
            ptr = torch.zeros(U + 1, dtype=torch.int64, device=dev)
            ptr[1:] = torch.cumsum(torch.bincount(keys // I, minlength=U), 0)
            ap_ptr, ap_col = ops.to_host(ptr), ops.to_host((keys % I).int())
            ops.pairwise_filter(ap_ptr, ap_col, U)                       # the clean dataset's filter blocks, built once
            del keys, ptr
        del eu, ei
        epoch_bufs = None
        for it in range(args.iters):
            fake = fake_profiles(rng, args.fake, I, target[0])        # same generator state on every rank
            frp, fit = dataset.ArrayImplicitData.fake_rows(fake, 4)
            F = len(frp) - 1
            sync(); dist.barrier(); t0 = time.time()
            m = base.inject(frp, fit)
            sync(); dist.barrier(); t1 = time.time()
            n2 = n + len(fit)
            if epoch_bufs is None or epoch_bufs.numel() < 4 * n2:
                epoch_bufs = torch.empty(4 * (n2 + 64 * args.fake), dtype=torch.int32, device=dev)
            ep = epoch_bufs[:4 * n2].view(4, n2)
            if rank == 0:
                # positives of the attacked dataset = the clean ones + the appended fake rows; the sampler's per-user filter
                # blocks of the genuine users are copied from the clean dataset's (ops.filter_parent_hint)
                ptr2 = ops.host_empty(len(ap_ptr) + F, np.int64)
                ptr2[:len(ap_ptr)] = ap_ptr
                ptr2[len(ap_ptr):] = ap_ptr[-1] + frp[1:]
                col2 = ops.host_empty(len(ap_col) + len(fit), np.int32)
                col2[:len(ap_col)] = ap_col
                col2[len(ap_col):] = fit
                ops.filter_parent_hint(ptr2, col2, ap_ptr, ap_col, U)
                st, key, pos = ops._np_state()
                host = [ops.host_empty(n2, np.uint32) for _ in range(4)]
                perm = ops.host_empty(n2, np.int32)
                cnt = ops.mt_pairwise_soa_raw(key, pos, U + F, I, n2, ptr2, col2, host[0], host[1], host[2], host[3])
                assert cnt == n2
                ops._np_state_commit(st, key, pos)
                ops.permutation_apply32(host[3][:cnt], perm)
                for k, a in enumerate((host[0], host[1], host[2], perm)):
                    ep[k].copy_(torch.from_numpy(a.view(np.int32)))
            dist.broadcast(ep, 0)
            sync(); dist.barrier(); t2 = time.time()
            loss = m.train_epoch_soa(ep[0], ep[1], ep[2], ep[3])
            sync(); dist.barrier(); t3 = time.time()
            genuine = torch.arange(min(m.Ug, max(0, U - m.lo)), device=dev)
            _, _, rank_, _ = m.full_rank(target, 20, users_local=genuine)
            hits = (rank_[:, 0] < 20).sum().double().view(1)
            dist.all_reduce(hits)
            hr = float(hits.item()) / U
            sync(); dist.barrier(); t4 = time.time()
            for k, v in zip(phases, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
                phases[k].append(v)
            hrs.append(hr)
            del m
    if rank == 0:
        out = {"tool": "attack_loop", "workload": args.workload, "n_gpus": world, "iters": args.iters, "fake_users": args.fake,
               "loss_last": loss, "HR@20_target_last": hrs[-1]}
        for k, v in phases.items():
            out[k] = round(float(np.mean(v[1:] if len(v) > 1 else v)), 4)      # first iteration warms allocators / JIT-free
            out[k + "_first"] = round(float(v[0]), 4)
        out["iter_s"] = round(sum(out[k] for k in phases), 4)
        out["note"] = ("inject = append rows + re-normalise + host copy of the positives + fresh model; sample = exact MT19937 epoch "
                       "draw on the host + H2D" + (" + NCCL broadcast" if world > 1 else "") + "; not overlapped here")
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
