"""Worker of tests/test_gpu_multi.py: rank `rank` of a `world`-rank NCCL job runs two PPO minibatch updates on its
shard of a fixed batch and writes its final parameters; rank 0 first computes the single-process result on the whole batch."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def make_batch(n, seed, dev):
    from egogen_b200.ppo_policy import Batch
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    obs = {"state": (r(n, 2, 402) * 0.3).to(dev), "egosensing": torch.rand(n, 2, 32, generator=g).to(dev),
           "dist": torch.rand(n, generator=g).to(dev), "time": torch.rand(n, generator=g).to(dev)}
    return Batch(obs=obs, act=r(n, 128).to(dev), logp_old=(r(n) * 0.1 - 180.0).to(dev), adv=r(n).to(dev), returns=r(n).to(dev))


def shard(b, lo, hi):
    from egogen_b200.ppo_policy import Batch
    return Batch(obs={k: v[lo:hi].contiguous() for k, v in b.obs.items()}, act=b.act[lo:hi].contiguous(),
                 logp_old=b.logp_old[lo:hi].contiguous(), adv=b.adv[lo:hi].contiguous(), returns=b.returns[lo:hi].contiguous())


def build(dev):
    from egogen_b200.crowd_env import default_cfg
    from egogen_b200.runtime import build_policy
    torch.manual_seed(7)
    np.random.seed(7)
    pol, _ = build_policy(default_cfg(), dev)
    pol.train()
    return pol


def synth_grad(n, r):
    g = torch.Generator().manual_seed(500 + r)
    return torch.randn(n, generator=g) * 1e-3


def run(rank, world, port, out_dir, per_rank, env):
    os.environ.update(env)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    if rank == 0:                                  # single-process results on the concatenated batch (before the group exists)
        pol = build(dev)
        pol.loss_backward(make_batch(per_rank * world, 100, dev))
        g1 = pol.flat_grads.clone()
        pol.learn_minibatch(make_batch(per_rank * world, 100, dev))        # same gradient again + optimiser step
        p1 = pol.flat_params.clone()
        pol.flat_grads.copy_(sum(synth_grad(pol.n_params, r) for r in range(world)).to(dev))
        pol.optimizer_step()
        torch.save({"n_ac": pol.n_actor_critic, "max_norm": float(pol._grad_norm), "grads": g1.cpu(), "params1": p1.cpu(), "params2": pol.flat_params.cpu(), "stats": pol._stats.cpu()},
                   os.path.join(out_dir, "single.pt"))
        del pol
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    pol = build(dev)
    used = "nccl-allreduce" if pol._dp is None else ("multicast" if pol._dp["mc"]["grads"] else "peer")
    if pol._dp is not None and os.environ.get("EG_DP_OVERLAP", "0") == "1":
        used += "+overlap"
    full = make_batch(per_rank * world, 100, dev)
    mine = shard(full, rank * per_rank, (rank + 1) * per_rank)
    pol.loss_backward(mine)
    g_local = pol.flat_grads.clone()
    stats = pol.learn_minibatch(mine)              # the product path: (overlapped) gradient sum + sharded optimiser step
    p1 = pol.flat_params.clone()
    pol.flat_grads.copy_(synth_grad(pol.n_params, rank).to(dev))
    pol.optimizer_step()
    sd = pol.export_optim_state()
    torch.cuda.synchronize(dev)
    torch.save({"grads_local": g_local.cpu(), "params1": p1.cpu(), "params2": pol.flat_params.cpu(), "stats": stats.cpu(),
                "path": used, "exp_avg0": sd["state"][0]["exp_avg"].cpu()}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()
