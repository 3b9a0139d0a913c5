"""Cost of the sharded NVLink optimiser step in isolation (run under torchrun with N ranks):
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_dp_step.py"""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from egogen_b200 import _lib
from egogen_b200.crowd_env import default_cfg
from egogen_b200.runtime import build_policy

pol, _ = build_policy(default_cfg(), dev)
pol.flat_grads.normal_(0, 1e-3)
dp = pol._dp
lib, st = _lib.lib(), _lib.stream_ptr(dev)


def timed(fn, n=50):
    for _ in range(5):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) * 1e3


res = {"optimizer_step_us": timed(pol.optimizer_step)}
if dp is not None:
    bar = dp["hdl"]["grads"].barrier
    res["three_barriers_us"] = timed(lambda: (bar(channel=0), bar(channel=0), bar(channel=0)))
    W, n_pad = dp["world"], dp["n_pad"]

    def kernels_only():
        lib.eg_dp_reduce_norm(dp["ptrs"]["grads"], C.c_void_p(dp["mc"]["grads"] or None), W, rank, n_pad, pol.n_actor_critic,
                              _lib.ptr(dp["gred"]), dp["ptrs"]["scratch"], _lib.ptr(dp["work"]), st)
        lib.eg_dp_adamw_gather(dp["ptrs"]["params"], C.c_void_p(dp["mc"]["params"] or None), W, rank, n_pad, pol.n_actor_critic,
                               _lib.ptr(dp["gred"]), _lib.ptr(dp["scratch"]), _lib.ptr(pol.exp_avg), _lib.ptr(pol.exp_avg_sq),
                               0.1, 3e-4, 0.9, 0.999, 1e-8, 0.01, 7, st)
    res["two_kernels_no_barrier_us"] = timed(kernels_only)
    res["path"] = "multicast" if dp["mc"]["grads"] else "peer"
t = torch.zeros(pol.n_params, device=dev)
res["nccl_allreduce_52MB_us"] = timed(lambda: dist.all_reduce(t))
small = torch.zeros(12, dtype=torch.float64, device=dev)
res["nccl_allreduce_small_us"] = timed(lambda: dist.all_reduce(small))
if rank == 0:
    print(world, res)
dist.destroy_process_group()
