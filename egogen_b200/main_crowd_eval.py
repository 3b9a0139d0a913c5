#!/usr/bin/env python3
"""Crowd evaluation entrypoint - mirror of the reference's motion/crowd_ppo/main_crowd_eval.py (BASELINE config 4):
a trained policy drives ``--n-agents`` (4) humans per scene that must avoid each other; every agent sees the others
as holes of the floor polygon (crowd_env_crowd_eval.py, dummy_vector_env.py). The reference runs ONE 4-agent scene
on one GPU; ``--n-scenes`` batches many independent scenes per GPU (agents of a scene stay on one GPU, scenes shard
over ranks with no collective - SURVEY.md 8e).

  python -m egogen_b200.main_crowd_eval --n-scenes 64 [--resume-path data/checkpoint_best.pth] [--deterministic-eval]

Start / goal pairs follow the reference's ``__main__`` block (agents start on the corners / edges of a square and
cross to the opposite side, main_crowd_eval.py:250-290); motion seeds are synthetic because the licensed SMPL-X
model and the SAMP locomotion clips are not redistributable."""
import argparse
import os
import time

import numpy as np
import torch

from . import assets
from .crowd_env import BoxSceneSampler, CrowdSceneVectorEnv, default_cfg_box
from .models_gamma_primitive import GAMMAPrimitiveComboGenOP, load_vposer
from .ppo_policy import Batch
from .runtime import build_policy
from .smplx_parser import get_lbs_model


def get_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--task", type=str, default="collision-avoidance")
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--lr", type=float, default=3e-4)
    p.add_argument("--gamma", type=float, default=0.99)
    p.add_argument("--gae-lambda", type=float, default=0.95)
    p.add_argument("--vf-coef", type=float, default=1.0)
    p.add_argument("--ent-coef", type=float, default=0.01)
    p.add_argument("--weight-kld", type=float, default=10.0)
    p.add_argument("--max-grad-norm", type=float, default=0.1)
    p.add_argument("--eps-clip", type=float, default=0.2)
    p.add_argument("--norm-adv", type=int, default=1)
    p.add_argument("--test-num", type=int, default=4)
    p.add_argument("--logdir", type=str, default="./log/log_eval_crowd_motion")
    p.add_argument("--device", type=str, default="cuda")
    p.add_argument("--resume-path", type=str, default="data/checkpoint_best.pth")
    p.add_argument("--deterministic-eval", default=False, action="store_true")
    # extensions of the batched build
    p.add_argument("--n-scenes", type=int, default=1, help="independent 4-agent scenes stepped together on this GPU")
    p.add_argument("--n-agents", type=int, default=4)
    p.add_argument("--jacobi", default=False, action="store_true",
                   help="step all agents against the previous step's boxes (one launch sequence) instead of the "
                        "reference's agent-by-agent update order")
    p.add_argument("--body-model-path", type=str, default=None)
    p.add_argument("--motion-results-root", type=str, default="results/crowd_ppo")
    p.add_argument("--predictor-dir", type=str, default=None)
    p.add_argument("--regressor-dir", type=str, default=None)
    p.add_argument("--vposer-dir", type=str, default=None)
    p.add_argument("--synthetic-assets", default=False, action="store_true")
    return p.parse_args(argv)


def crowd_start_data(sampler, n_scenes, n_agents, device, seed=0):
    """Per-agent start bodies and goals: agent a of every scene starts on a circle of radius 3 m facing the centre and
    walks to the diametrically opposite point (the crossing pattern of main_crowd_eval.py's 4 init_env blocks)."""
    E = n_scenes * n_agents
    s = sampler.next_body(E)
    wp, goals, betas = s["world_params"].clone(), s["goals"].clone(), s["betas"].clone()
    rng = np.random.default_rng(seed)
    for a in range(n_agents):
        for sc in range(n_scenes):
            e = a * n_scenes + sc
            ang = 2 * np.pi * a / n_agents + rng.uniform(-0.2, 0.2)
            pos = 3.0 * np.array([np.cos(ang), np.sin(ang)])
            d = wp[e, 1, :2] - wp[e, 0, :2]
            wp[e, 0, :2] = torch.as_tensor(pos, dtype=torch.float32)
            wp[e, 1, :2] = wp[e, 0, :2] + d
            goals[e, :2] = torch.as_tensor(-pos, dtype=torch.float32)
    return wp.to(device), goals.to(device), betas.to(device)


def build_crowd_world(device, n_scenes, n_agents=4, sequential=True, seed=0, body_model_path=None, args=None,
                      predictor_dir=None, regressor_dir=None, vposer_dir=None):
    dev = torch.device(device)
    cfg = default_cfg_box()
    markers = assets.marker_ids()
    lbs = get_lbs_model("male", dev, body_model_path=body_model_path, marker_vids=markers)
    genop = GAMMAPrimitiveComboGenOP(testconfig={"gpu_index": dev.index or 0})
    genop.build_model(load_pretrained_model=bool(predictor_dir), predictor_dir=predictor_dir, regressor_dir=regressor_dir, seed=0)
    vposer, _ = load_vposer(vposer_dir, seed=0, device=dev)
    scene = assets.make_box_scene(seed, n_boxes=0)
    sdf = {k: v.to(dev) for k, v in assets.rasterize_scene_sdf(scene, D=64, device=str(dev)).items()}
    venv = CrowdSceneVectorEnv(cfg, genop.model, lbs, vposer, sdf, n_scenes, dev, n_agents=n_agents, sequential=sequential)
    sampler = BoxSceneSampler(sdf, lbs, dev, seed=seed)
    policy, optim = build_policy(cfg, dev, args)
    return dict(cfg=cfg, venv=venv, sampler=sampler, policy=policy, optim=optim, lbs=lbs, genop=genop, vposer=vposer)


@torch.no_grad()
def run_episodes(w, wp, goals, betas, max_steps=None):
    """test_collector.collect(n_episode=...) over the crowd vector env: every agent runs until its episode terminates
    (goal or max_depth); finished agents keep their last box as an obstacle, like a stopped CrowdEnv worker."""
    venv, pol = w["venv"], w["policy"]
    E = venv.E
    venv.reset_from(torch.arange(E), wp, goals, betas)
    ret = torch.zeros(E, device=venv.dev)
    length = torch.zeros(E, dtype=torch.int32, device=venv.dev)
    alive = torch.ones(E, dtype=torch.bool, device=venv.dev)
    reached = torch.zeros(E, dtype=torch.bool, device=venv.dev)
    steps = 0
    max_steps = max_steps or int(w["cfg"].trainconfig.max_depth)
    while bool(alive.any()) and steps < max_steps:
        out = pol.forward(Batch(obs=venv.observation()))
        _, rew, term, _, _ = venv.step(out.act)
        ret += torch.where(alive, rew, torch.zeros_like(rew))
        length += alive.to(torch.int32)
        reached |= alive & venv.buf["goal_reached"].bool()
        alive &= ~term.bool()
        steps += 1
    return dict(rews=ret.cpu().numpy(), lens=length.cpu().numpy(), reached=reached.cpu().numpy(), vector_steps=steps)


def main(args=None):
    args = args or get_args()
    dev = torch.device(args.device if torch.cuda.is_available() else "cpu")
    if dev.type != "cuda":
        raise SystemExit("main_crowd_eval needs a CUDA device (no CPU path exists)")
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        torch.cuda.set_device(dev)
    np.random.seed(args.seed + rank)
    torch.manual_seed(args.seed + rank)
    from .main_ppo import resolve_asset_dirs
    pdir, rdir, vdir = resolve_asset_dirs(args)
    w = build_crowd_world(dev, args.n_scenes, args.n_agents, sequential=not args.jacobi, seed=args.seed + rank,
                          body_model_path=args.body_model_path, args=args, predictor_dir=pdir, regressor_dir=rdir, vposer_dir=vdir)
    if args.resume_path and os.path.exists(args.resume_path):
        ckpt = torch.load(args.resume_path, map_location=dev)
        w["policy"].load_state_dict(ckpt["model"])
        print("Loaded agent from: ", args.resume_path)
    w["policy"].eval()
    wp, goals, betas = crowd_start_data(w["sampler"], args.n_scenes, args.n_agents, dev, seed=args.seed + rank)
    t0 = time.perf_counter()
    res = run_episodes(w, wp, goals, betas)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    print(f'Final reward: {res["rews"].mean()}, length: {res["lens"].mean()}')
    print(f'[rank {rank}] {res["lens"].sum()} agent steps in {dt:.3f} s, goal reached by {int(res["reached"].sum())} / {len(res["rews"])} agents')
    return res


if __name__ == "__main__":
    main()
