#!/usr/bin/env python3
"""crowd_ppo training / evaluation entrypoint - keeps the flags, checkpoint layout and logging keys of the
reference's motion/crowd_ppo/main_ppo.py (get_args :40-92, main :95-243) with the environments, policy and
optimiser running on the egogen_b200 CUDA library. tianshou's onpolicy_trainer / Collector are replaced by the
loop below (SURVEY.md Appendix A5 semantics).

  python -m egogen_b200.main_ppo [--training-num 256 --step-per-collect 1024 ...]
  torchrun --nproc-per-node 8 -m egogen_b200.main_ppo ...        # envs sharded over ranks, one NCCL
                                                                 # gradient allreduce per optimiser step
"""
import argparse
import datetime
import os
import pprint

import numpy as np
import torch


def get_args(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--task", type=str, default="collision-avoidance")
    parser.add_argument("--seed", type=int, default=0)
    parser.add_argument("--buffer-size", type=int, default=4096)
    parser.add_argument("--lr", type=float, default=3e-4)
    parser.add_argument("--gamma", type=float, default=0.99)
    parser.add_argument("--epoch", type=int, default=3000)
    parser.add_argument("--step-per-epoch", type=int, default=20000)
    parser.add_argument("--step-per-collect", type=int, default=1024)
    parser.add_argument("--repeat-per-collect", type=int, default=1)
    parser.add_argument("--batch-size", type=int, default=256)
    parser.add_argument("--training-num", type=int, default=256)
    parser.add_argument("--test-num", type=int, default=20)
    parser.add_argument("--rew-norm", type=int, default=False)
    parser.add_argument("--vf-coef", type=float, default=1.0)
    parser.add_argument("--ent-coef", type=float, default=0.01)
    parser.add_argument("--weight-kld", type=float, default=0)
    parser.add_argument("--gae-lambda", type=float, default=0.95)
    parser.add_argument("--bound-action-method", type=str, default="clip")
    parser.add_argument("--max-grad-norm", type=float, default=0.1)
    parser.add_argument("--eps-clip", type=float, default=0.1)
    parser.add_argument("--dual-clip", type=float, default=None)
    parser.add_argument("--value-clip", type=int, default=0)
    parser.add_argument("--norm-adv", type=int, default=1)
    parser.add_argument("--recompute-adv", type=int, default=0)
    parser.add_argument("--logdir", type=str, default="./log")
    parser.add_argument("--render", type=float, default=0.0)
    parser.add_argument("--device", type=str, default="cuda")
    parser.add_argument("--resume-path", type=str, default=None)
    parser.add_argument("--resume-buffer", type=str, default=None)
    parser.add_argument("--resume-id", type=str, default=None)
    parser.add_argument("--finetune", default=False, action="store_true")
    parser.add_argument("--deterministic-eval", default=False, action="store_true")
    parser.add_argument("--logger", type=str, default="tensorboard", choices=["tensorboard", "wandb"])
    parser.add_argument("--save-interval", type=int, default=2)
    parser.add_argument("--wandb-project", type=str, default="mujoco.benchmark")
    parser.add_argument("--watch", default=False, action="store_true",
                        help="watch the play of pre-trained policy only")
    # additions (asset locations; the reference hard-codes them relative to motion/)
    parser.add_argument("--body-model-path", type=str, default=None)
    parser.add_argument("--scene-sdf", type=str, default=None, help="data/room0_sdf.pkl; synthetic box scene if absent")
    parser.add_argument("--scene-poly", type=str, default=None, help="data/replica_room0_shapely.pkl")
    parser.add_argument("--sdf-res", type=int, default=256)
    parser.add_argument("--motion-results-root", type=str, default="results/crowd_ppo",
                        help="root of the reference's GAMMA checkpoints (<root>/MPVAE_samp20_2frame_rollout/checkpoints/epoch-400.ckp, "
                             "<root>/MoshRegressor_v3_male/checkpoints/epoch-100.ckp, primitive_model.py:56-72); used when present")
    parser.add_argument("--predictor-dir", type=str, default=None, help="overrides the predictor checkpoint directory")
    parser.add_argument("--regressor-dir", type=str, default=None, help="overrides the regressor checkpoint directory")
    parser.add_argument("--vposer-dir", type=str, default=None,
                        help="VPoser v1.0 expr dir (<body-model-path>/vposer_v1_0 in the reference, main_ppo.py:259)")
    parser.add_argument("--synthetic-assets", default=False, action="store_true",
                        help="run on seeded synthetic motion-model / VPoser weights without looking for checkpoints")
    return parser.parse_args(argv)


def resolve_asset_dirs(args):
    """Checkpoint directories of the motion model and VPoser: explicit flags win; otherwise the reference's relative
    locations are used when they exist. A directory that was NAMED but lacks its checkpoint raises (build_world)."""
    from .runtime import motion_checkpoint_dirs
    if getattr(args, "synthetic_assets", False):
        return None, None, None
    pdir, rdir = args.predictor_dir, args.regressor_dir
    if not (pdir and rdir):
        cp, cr = motion_checkpoint_dirs(args.motion_results_root)
        explicit_root = args.motion_results_root != "results/crowd_ppo"
        if explicit_root or (os.path.isdir(cp) and os.path.isdir(cr)):
            pdir, rdir = pdir or cp, rdir or cr
    vdir = args.vposer_dir
    if vdir is None and args.body_model_path and os.path.isdir(os.path.join(args.body_model_path, "vposer_v1_0")):
        vdir = os.path.join(args.body_model_path, "vposer_v1_0")
    return pdir, rdir, vdir


def shard(total, world, rank, what):
    if total % world:
        raise ValueError(f"--{what}={total} must be a multiple of the number of ranks ({world})")
    return total // world


def main(args=None):
    args = args or get_args()
    import torch.distributed as dist
    from torch.utils.tensorboard import SummaryWriter
    from . import assets
    from .collector import Collector
    from .runtime import build_world
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("egogen_b200 has no CPU path: a CUDA device is required")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    np.random.seed(args.seed + rank)
    torch.manual_seed(args.seed + rank)

    scene_sdf = scene_rings = None
    if args.scene_sdf and os.path.exists(args.scene_sdf):
        scene_sdf = assets.load_scene_sdf(args.scene_sdf)
        scene_rings = assets.load_wkb_polygon(args.scene_poly)
    n_train = shard(args.training_num, world, rank, "training-num")
    n_collect = shard(args.step_per_collect, world, rank, "step-per-collect")
    n_batch = shard(args.batch_size, world, rank, "batch-size")
    pdir, rdir, vdir = resolve_asset_dirs(args)
    w = build_world(dev, n_train, seed=args.seed + rank, sdf_res=args.sdf_res, finetuning=args.finetune,
                    body_model_path=args.body_model_path, scene_sdf=scene_sdf, scene_rings=scene_rings, args=args,
                    box_mode=getattr(args, "box_mode", False), predictor_dir=pdir, regressor_dir=rdir, vposer_dir=vdir)
    policy, optim, train_collector = w["policy"], w["optim"], w["collector"]
    tw = build_world(dev, args.test_num, seed=args.seed + 1000 + rank, sdf_res=args.sdf_res, finetuning=args.finetune,
                     body_model_path=args.body_model_path, scene_sdf=None if getattr(args, "box_mode", False) else w["scene_sdf"],
                     scene_rings=None if getattr(args, "box_mode", False) else w["scene_rings"], with_policy=False,
                     box_mode=getattr(args, "box_mode", False), predictor_dir=pdir, regressor_dir=rdir, vposer_dir=vdir,
                     capture_rollout=True)
    # the reference's CrowdEnv writes a rollout pickle at every episode end (save_rollout=True by default,
    # crowd_env_2f.py:35,305-309); here the evaluation envs do, into the same folder
    tw["venv"].save_rollout = True
    tw["venv"].rollout_dir = os.path.join(args.logdir, "eval_results") if args.logdir != "./log" else "./log/eval_results/"
    test_collector = Collector(policy, tw["venv"])
    w["venv"].seed(args.seed + rank)
    tw["venv"].seed(args.seed + rank)

    if args.resume_path:
        ckpt = torch.load(args.resume_path, map_location=dev)
        policy.load_state_dict(ckpt["model"])      # optimiser state is NOT restored, like main_ppo.py:165-171
        print("Loaded agent from: ", args.resume_path)

    now = datetime.datetime.now().strftime("%y%m%d-%H%M%S")
    log_path = os.path.join(args.logdir, args.task, "ppo", str(args.seed), now)
    writer = None
    if rank == 0:
        writer = SummaryWriter(log_path)
        writer.add_text("args", str(args))

    def save_best_fn():
        torch.save({"model": policy.state_dict(), "optim": policy.export_optim_state()}, os.path.join(log_path, "policy.pth"))

    def save_checkpoint_fn(epoch):
        p = os.path.join(log_path, f"checkpoint_{epoch}.pth")
        torch.save({"model": policy.state_dict(), "optim": policy.export_optim_state()}, p)
        return p

    result = {}
    if not args.watch:
        policy.train()
        train_collector.reset()
        env_step = gradient_step = 0
        best = -float("inf")
        policy.eval()                                  # tianshou's trainer evaluates once before the first epoch
        tr0 = test_collector.collect_episodes(args.test_num)
        policy.train()
        best = tr0["rew"]
        if rank == 0:
            print(f"Epoch #0: test_reward: {tr0['rew']:.6f} +/- {tr0['rew_std']:.6f}")
            if writer:
                writer.add_scalar("test/reward", tr0["rew"], 0)
                writer.add_scalar("test/length", tr0["len"], 0)
        for epoch in range(1, args.epoch + 1):
            epoch_step = 0
            while epoch_step < args.step_per_epoch:
                batch, st = train_collector.collect(n_collect)
                losses = policy.learn(batch, n_batch, args.repeat_per_collect)
                epoch_step += args.step_per_collect
                env_step += args.step_per_collect
                gradient_step += len(losses["loss"])
                if writer:
                    if st.get("n/ep", 0) > 0:
                        writer.add_scalar("train/reward", st["rew"], env_step)
                        writer.add_scalar("train/length", st["len"], env_step)
                        writer.add_scalar("train/episode", train_collector.collect_episode, env_step)
                    for k, v in losses.items():
                        writer.add_scalar("update/" + k, float(np.mean(v)), env_step)
            policy.eval()
            tr = test_collector.collect_episodes(args.test_num)
            policy.train()
            if writer:
                writer.add_scalar("test/reward", tr["rew"], env_step)
                writer.add_scalar("test/length", tr["len"], env_step)
                writer.add_scalar("test/reward_std", tr["rew_std"], env_step)
                writer.add_scalar("test/length_std", tr["len_std"], env_step)
            if rank == 0:
                print(f"Epoch #{epoch}: test_reward: {tr['rew']:.6f} +/- {tr['rew_std']:.6f}, best_reward: {max(best, tr['rew']):.6f}")
                if tr["rew"] > best:
                    best = tr["rew"]
                    save_best_fn()
                if epoch % args.save_interval == 0:
                    save_checkpoint_fn(epoch)
            result = {"best_reward": best, "epoch": epoch, "env_step": env_step, "gradient_step": gradient_step}
        if rank == 0:
            pprint.pprint(result)

    policy.eval()
    tw["venv"].seed(args.seed)
    r = test_collector.collect_episodes(args.test_num)
    if rank == 0:
        print(f'Final reward: {r["rews"].mean() if len(r["rews"]) else 0.0}, length: {r["lens"].mean() if len(r["lens"]) else 0.0}')
    if world > 1:
        dist.destroy_process_group()
    return result


if __name__ == "__main__":
    main()
