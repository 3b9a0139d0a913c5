"""Wiring of the crowd_ppo hot path from (synthetic or real) assets: one call builds the operator handles,
the vectorised env, the PPO policy and the collector on one GPU. Mirrors the asset wiring block of the
reference's main_ppo.py:246-309 and the policy construction of :95-162."""
from __future__ import annotations

import numpy as np
import torch

from . import assets
from .collector import Collector
from .crowd_env import BoxSceneSampler, CrowdVectorEnv, default_cfg, default_cfg_box
from .models_gamma_primitive import GAMMAPrimitiveComboGenOP, load_vposer
from .models_policy_ppo import ActorCritic, GAMMAActor, GAMMACritic, GAMMAPolicyBase
from .ppo_policy import GAMMAPPOPolicy
from .smplx_parser import get_lbs_model


def build_policy(cfg, device, args=None, process_group=None):
    """main_ppo.py:104-162: nets, orthogonal init (gain sqrt 2, zero bias), actor.pnet Linears x0.01, AdamW."""
    a = args
    mc = vars(cfg.modelconfig) if not isinstance(cfg.modelconfig, dict) else cfg.modelconfig
    actor, critic, shared = GAMMAActor(mc).to(device), GAMMACritic(mc).to(device), GAMMAPolicyBase(mc).to(device)
    actor_critic = ActorCritic(actor, critic, shared)
    for m in actor_critic.modules():
        if isinstance(m, torch.nn.Linear):
            torch.nn.init.orthogonal_(m.weight, gain=np.sqrt(2))
            torch.nn.init.zeros_(m.bias)
    for m in actor_critic.actor.pnet.modules():
        if isinstance(m, torch.nn.Linear):
            torch.nn.init.zeros_(m.bias)
            m.weight.data.copy_(0.01 * m.weight.data)
    g = lambda k, d: getattr(a, k, d) if a is not None else d
    optim = torch.optim.AdamW(actor_critic.parameters(), lr=g("lr", 3e-4), weight_decay=0.01)
    policy = GAMMAPPOPolicy(actor, critic, shared, optim, None, discount_factor=g("gamma", 0.99),
                            gae_lambda=g("gae_lambda", 0.95), max_grad_norm=g("max_grad_norm", 0.1),
                            vf_coef=g("vf_coef", 1.0), ent_coef=g("ent_coef", 0.01), weight_kld=g("weight_kld", 0),
                            reward_normalization=g("rew_norm", False), action_space=None, action_scaling=False,
                            action_bound_method="", eps_clip=g("eps_clip", 0.1), value_clip=g("value_clip", 0),
                            dual_clip=g("dual_clip", None), advantage_normalization=g("norm_adv", 1),
                            recompute_advantage=g("recompute_adv", 0), deterministic_eval=g("deterministic_eval", False),
                            process_group=process_group)
    return policy, optim


def motion_checkpoint_dirs(results_root):
    """Where the reference's ConfigCreator puts the GAMMA checkpoints (primitive_model.py:9-37,56-72):
    <root>/MPVAE_samp20_2frame_rollout/checkpoints and <root>/MoshRegressor_v3_male/checkpoints."""
    import os
    return (os.path.join(results_root, "MPVAE_samp20_2frame_rollout", "checkpoints"),
            os.path.join(results_root, "MoshRegressor_v3_male", "checkpoints"))


def build_world(device, n_envs: int, seed: int = 0, sdf_res: int = 256, n_boxes: int = 1, finetuning: bool = False,
                body_model_path=None, scene_sdf=None, scene_rings=None, args=None, with_policy: bool = True,
                host_boundary: bool = False, process_group=None, cfg=None, box_mode: bool = False,
                predictor_dir=None, regressor_dir=None, vposer_dir=None, sampler=None, capture_rollout: bool = False):
    """predictor_dir / regressor_dir / vposer_dir: the pretrained GAMMA predictor, body regressor and VPoser v1.0 the
    reference always loads (primitive_model.py:70, main_ppo.py:259); a missing file raises. Left out, seeded synthetic
    weights are used (with a warning). sampler: any object with next_body(n) -> {world_params, goals, betas} (e.g. built
    from scene_sampler.CrowdMotionSampler dicts); default BoxSceneSampler over the scene's own extent."""
    dev = torch.device(device)
    cfg = cfg or (default_cfg_box() if box_mode else default_cfg())
    scene = None
    markers = assets.marker_ids()
    lbs = get_lbs_model("male", dev, body_model_path=body_model_path, marker_vids=markers)
    genop = GAMMAPrimitiveComboGenOP(testconfig={"gpu_index": dev.index or 0})
    if bool(predictor_dir) != bool(regressor_dir):
        raise ValueError("predictor_dir and regressor_dir must be given together")
    genop.build_model(load_pretrained_model=bool(predictor_dir), predictor_dir=predictor_dir, regressor_dir=regressor_dir, seed=0)
    vposer, _ = load_vposer(vposer_dir, seed=0, device=dev)
    if scene_sdf is None:
        scene = assets.make_box_scene(seed, n_boxes=n_boxes)
        scene_sdf = assets.rasterize_scene_sdf(scene, D=sdf_res, device=str(dev))
        scene_rings = assets.scene_polygon(scene)
    scene_sdf = {k: torch.as_tensor(v, dtype=torch.float32).to(dev) for k, v in scene_sdf.items()}
    if sampler is None:
        sampler = BoxSceneSampler(scene_sdf, lbs, dev, seed=seed)
    sampler.scene_rings = scene_rings
    tris = None
    if box_mode:
        if scene is None:
            raise ValueError("box_mode needs a synthetic box scene (or pass navmesh triangles through CrowdVectorEnv)")
        tris = assets.scene_navmesh_triangles(scene)
    venv = CrowdVectorEnv(cfg, genop.model, lbs, vposer, scene_sdf, scene_rings, sampler, n_envs, dev,
                          finetuning=finetuning, box_mode=box_mode, navmesh_tris=tris, capture_rollout=capture_rollout)
    out = dict(cfg=cfg, lbs=lbs, genop=genop, vposer=vposer, scene_sdf=scene_sdf, scene_rings=scene_rings,
               sampler=sampler, venv=venv)
    if with_policy:
        policy, optim = build_policy(cfg, dev, args, process_group)
        out.update(policy=policy, optim=optim, collector=Collector(policy, venv, host_boundary=host_boundary))
    return out
