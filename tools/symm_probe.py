"""Probe what the box offers for peer-memory kernels: torchrun --nproc-per-node N tools/symm_probe.py"""
import os

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
I, D = 200000, 64


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


x = torch.randn(I, D, device=dev)
t_nccl = timeit(lambda: dist.all_reduce(x))
if rank == 0:
    print(f"nccl all_reduce {I}x{D} fp32: {t_nccl:.3f} ms", flush=True)
small = torch.randn(I, device=dev)
t_small = timeit(lambda: dist.all_reduce(small))
if rank == 0:
    print(f"nccl all_reduce {I} fp32: {t_small:.3f} ms", flush=True)
try:
    t = symm.empty((I, D), dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD)
    if rank == 0:
        print("symm ok: multicast ptr", hex(int(getattr(hdl, "multicast_ptr", 0) or 0)), "bufs", [hex(p) for p in hdl.buffer_ptrs],
              "signal pad", getattr(hdl, "signal_pad_size", None), flush=True)
    t.normal_()
    tb = timeit(lambda: hdl.barrier(channel=0))
    if rank == 0:
        print(f"symm barrier: {tb * 1e3:.1f} us", flush=True)
    gname = dist.group.WORLD.group_name
    for opname in ("multimem_all_reduce_", "one_shot_all_reduce", "two_shot_all_reduce_"):
        try:
            op = getattr(torch.ops.symm_mem, opname)
            tt = timeit(lambda: op(t, "sum", gname))
            if rank == 0:
                print(f"symm_mem.{opname}: {tt:.3f} ms", flush=True)
        except Exception as e:   # noqa: BLE001
            if rank == 0:
                print(f"symm_mem.{opname}: FAILED {type(e).__name__}: {str(e)[:200]}", flush=True)
except Exception as e:   # noqa: BLE001
    print(f"[{rank}] symmetric memory FAILED {type(e).__name__}: {str(e)[:300]}", flush=True)
dist.destroy_process_group()
