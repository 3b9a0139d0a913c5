#!/usr/bin/env python3
"""Headline benchmark of the crowd_ppo hot path (BASELINE.json metric: env-steps/sec).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
  python bench.py --impl reference [...]                          # the reference-shaped CPU path (oracle)

A "step" is one PPO iteration of config[1] per GPU: collect 1024 transitions with 256 parallel envs
(4 vector env steps: policy forward -> C-VAE rollout -> SMPL-X LBS on 20 bodies/env -> SDF penetration ->
rewards -> re-canonicalisation -> ego-sensing) followed by GAE and one learn pass (4 minibatches of 256,
forward + backward + grad-clip + AdamW, one NCCL gradient allreduce per optimiser step when N > 1).
value = env transitions of all ranks / time, with every input resident in HBM; e2e = the same iteration with the
reference's host-side data path (obs / actions / rewards cross pinned host memory every vector step).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ENVS = 256            # per GPU (main_ppo.py --training-num)
STEP_PER_COLLECT = 1024  # per GPU (main_ppo.py --step-per-collect)
BATCH = 256             # per-GPU minibatch (main_ppo.py --batch-size)
B_BODY = 127636         # algorithmic bytes per materialised body
FLOP_BODY = 49.1e6      # dense-reference FLOPs per body


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.rows[0][1]) if self.rows and self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(sm)}


def run_reference(args):
    """Reference arm: the reference's execution shape on the host CPUs (per-env sequential loop, 4x duplicated
    batch, four SMPL-X passes per step, dense skinning, fp64 2-D rays) - the oracle restatement, because the
    reference's own env is hard-wired to CUDA + absent third-party packages."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import harness
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    world = harness.build_oracle_world(0, sdf_res=256)
    n_envs, n_steps = 4, 2          # bounded sample of the 256-env x 4-step collect
    for i in range(args.warmup):
        harness.run_iteration(world, n_envs, 1, True, seed=100 + i)
    tot_s, tot_n = 0.0, 0
    for i in range(args.steps):
        s, n = harness.run_iteration(world, n_envs, n_steps, True, seed=i)
        tot_s += s; tot_n += n
    v = tot_n / tot_s
    sample = f"{n_envs} envs x {n_steps} vector steps per step (of 256 x 4), sequential per-env loop with 4x duplicated batch"
    line = {"impl": "reference", "metric": "crowd_ppo env-steps/sec", "value": v, "unit": "env-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.gpus),
            "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def config_dict(n):
    return {"workload": "PPO phase-1 training, 256 parallel envs per GPU, synthetic random-box scene (256^3 SDF), "
                        "surrogate SMPL-X (V=10475), step = collect 1024 transitions + 4 minibatches of 256",
            "envs_per_gpu": N_ENVS, "step_per_collect_per_gpu": STEP_PER_COLLECT, "batch_per_gpu": BATCH,
            "global_batch": BATCH * n, "parallelism": f"env-sharded dp{n}",
            "cache": "L2 flushed (256 MiB write) between timed iterations; working set ~250 MB > 126 MB L2"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from egogen_b200 import _lib
    from egogen_b200.runtime import build_world
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback exists)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234 + rank)
    import numpy as np
    np.random.seed(1234 + rank)
    lib = _lib.lib()

    w = build_world(dev, N_ENVS, seed=rank, sdf_res=256)
    col, pol = w["collector"], w["policy"]
    pol.train()
    col.reset()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    last = {}

    def iteration(c):
        batch, _ = c.collect(STEP_PER_COLLECT)
        last["batch"] = batch
        return pol.learn(batch, BATCH, 1)

    def timed(c, k, profile):
        ms = []
        launches0 = lib.eg_launch_count()
        for _ in range(k):
            flush.zero_()
            if world_size > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if profile:
                lib.eg_profile_enable(1)
            launches_before = lib.eg_launch_count()
            e0.record()
            iteration(c)
            e1.record()
            torch.cuda.synchronize(dev)
            ms.append(e0.elapsed_time(e1))
        return ms, lib.eg_launch_count() - launches0 - 0

    for _ in range(max(args.warmup, 3)):
        iteration(col)
    torch.cuda.synchronize(dev)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    lib.eg_profile_enable(1)
    ncu_range = os.environ.get("EG_NCU_RANGE") == "1"      # ncu --profile-from-start off: capture the timed region only
    if ncu_range:
        torch.cuda.cudart().cudaProfilerStart()
    ms, launches = timed(col, args.steps, False)
    if ncu_range:
        torch.cuda.synchronize(dev)
        torch.cuda.cudart().cudaProfilerStop()
    tot_ms, n_l, n_units = C.c_double(), C.c_int64(), C.c_int64()
    _lib.check(lib.eg_profile_read(C.byref(tot_ms), C.byref(n_l), C.byref(n_units)))
    lib.eg_profile_enable(0)
    # the flush kernel is torch's, not ours; do not count it. launches counts only this library's kernels.
    total_ms = float(sum(ms))
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = STEP_PER_COLLECT * world_size * args.steps / (total_ms / 1e3)

    # ---- e2e: same iteration with the reference's host-side data path -------------------------
    from egogen_b200.collector import Collector
    col2 = Collector(pol, w["venv"], host_boundary=True)
    col2.reset()
    iteration(col2)
    col2.h2d_bytes = col2.d2h_bytes = 0
    ms2, _ = timed(col2, args.steps, False)
    t2 = torch.tensor([float(sum(ms2))], device=dev, dtype=torch.float64)
    if world_size > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = STEP_PER_COLLECT * world_size * args.steps / (float(t2.item()) / 1e3)
    clocks = sampler.stop() if sampler else None
    e2e = {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": col2.h2d_bytes // args.steps,
           "d2h_bytes_per_step": (col2.d2h_bytes + 5 * 8 * 4) // args.steps}

    # sanity (outside every timed region): the rollouts and the updated policy are finite
    assert bool(torch.isfinite(last["batch"].returns).all()) and bool(torch.isfinite(pol.flat_params).all()), \
        "non-finite values after the timed iterations"
    if rank != 0:
        if world_size > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel (fused LBS + SDF vertex kernel) -----------------------
    peak, how = peaks()
    k_ms = tot_ms.value / max(n_l.value, 1)
    bodies_per_launch = n_units.value / max(n_l.value, 1)
    hbm_achieved = bodies_per_launch * B_BODY / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    # the fused LBS kernel is an fp16-operand / fp32-accumulate tensor-core contraction [bodies,576] x [576, 3*V]
    # (V = 10475 real vertices; tile padding is not counted) + skinning/SDF epilogue
    tc_flops = bodies_per_launch * 2.0 * 576 * 3 * 10475
    tf_achieved = tc_flops / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    tf_peak = float(pk.get("bf16_tflops", 1590.0))
    # DRAM traffic of the same kernel from the committed ncu --set full capture (profiles/), scaled to this run's
    # bodies per launch (the kernel's HBM traffic is the basis + features + transforms, linear in bodies apart from
    # the 36 MB fp16 basis that is read once per launch)
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r1_lbs_tc_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) * bodies_per_launch / tj["bodies"]
        traffic_src = tj["source"]
    roofline = {"bound": "tensor", "achieved": tf_achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": tf_achieved / tf_peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": how + " dense bf16/fp16 cuBLAS burst (kind::f16 operands, fp32 accumulation)",
                "kernel": "lbs_verts_tc_kernel<FUSE_SDF> (tcgen05 kind::f16, fp32 accumulate in TMEM)",
                "avg_launch_ms": k_ms, "bodies_per_launch": bodies_per_launch, "launches_timed": n_l.value,
                "kernel_share_of_step": tot_ms.value / float(sum(ms)),
                "flops_per_body": 2.0 * 576 * 3 * 10475,
                "hbm_contract": {"achieved": hbm_achieved, "peak": peak, "unit": "GB/s", "frac": hbm_achieved / peak,
                                 "frac_of_nominal_8TBs": hbm_achieved / 8000.0,
                                 "note": "SURVEY 8(d) contract figure: bodies x 127636 B (what an unfused LBS must move); "
                                         "the fused kernel itself writes only the per-body counts"},
                "reference_dense_tflops": bodies_per_launch * FLOP_BODY / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0}
    # ---- CPU baseline: the oracle port on a bounded sample, host cores of this box --------------
    from oracle import harness
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ow = harness.build_oracle_world(0, sdf_res=256)
    harness.run_iteration(ow, 8, 1, False, seed=99)                       # warm-up
    s, n = harness.run_iteration(ow, 16, 2, False, seed=0)
    cpu = {"value": n / s, "unit": "env-steps/s", "cores": cores, "kind": "port",
           "sample": "oracle (batched, dup removed): 16 envs x 2 vector steps + GAE + 1 learn pass, torch-CPU all threads"}
    line = {"metric": "crowd_ppo env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world_size,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(world_size), "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line))
    if world_size > 1:
        dist.destroy_process_group()


def run_ego_depth(args):
    """BASELINE config 5: ego-depth ray-march sweep, 8192 agents x 64x64 rays per GPU against a resident 256^3 SDF
    (agents sharded over ranks, grid replicated, no collective). Secondary workload: prints its own JSON line."""
    import torch
    import torch.distributed as dist
    from egogen_b200 import _lib, assets, ego_depth
    world_size = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=dev)
    A, H, W = 8192, 64, 64
    scene = assets.make_box_scene(rank, n_boxes=4)
    sdf = {k: v.to(dev) for k, v in assets.rasterize_scene_sdf(scene, D=256, device=str(dev)).items()}
    g = torch.Generator(device=dev); g.manual_seed(rank)
    eye = torch.cat([(torch.rand(A, 2, device=dev, generator=g) * 2 - 1) * 3.0, torch.full((A, 1), 1.6, device=dev)], 1)
    yaw = torch.rand(A, device=dev, generator=g) * 6.2831853
    fwd = torch.stack([yaw.cos(), yaw.sin(), torch.full((A,), -0.05, device=dev)], 1)
    fwd = fwd / fwd.norm(dim=1, keepdim=True)
    right = torch.cross(fwd, torch.tensor([0.0, 0.0, 1.0], device=dev).expand(A, 3), dim=1)
    right = right / right.norm(dim=1, keepdim=True)
    cam = torch.cat([eye, right, torch.cross(right, fwd, dim=1), fwd], 1).contiguous()
    fx = fy = 64 * (200.0 / 320.0)
    for _ in range(max(args.warmup, 3)):
        ego_depth(sdf, cam, H, W, fx, fy)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ms = []
    for _ in range(args.steps):
        flush.zero_()
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); depth, steps = ego_depth(sdf, cam, H, W, fx, fy, return_steps=True); e1.record()
        torch.cuda.synchronize(dev)
        ms.append(e0.elapsed_time(e1))
    t = torch.tensor([sum(ms)], device=dev, dtype=torch.float64)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    mean_steps = float(steps.float().mean().item()) + 1.0       # samples per ray (the terminating sample included)
    if rank == 0:
        peak, how = peaks()
        rays = A * H * W * world_size * args.steps
        secs = float(t.item()) / 1e3
        gathered = A * H * W * mean_steps * 32.0 / (sum(ms) / args.steps / 1e3) / 1e9     # bytes gathered / s, this rank
        print(json.dumps({"metric": "ego-depth rays/sec", "value": rays / secs, "unit": "rays/s", "n_gpus": world_size,
                          "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": float(t.item()) / args.steps,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": "ego-depth sweep: 8192 agents x 64x64 rays per GPU, 256^3 SDF, <=64 sphere-trace steps, 7 m range"},
                          "roofline": {"bound": "hbm", "achieved": gathered, "peak": peak, "unit": "GB/s", "frac": gathered / peak,
                                       "traffic": None, "peak_source": how, "kernel": "ego_depth_kernel",
                                       "mean_samples_per_ray": mean_steps,
                                       "note": "achieved = rays x samples x 32 B of corner gathers (L2-served; grid 67 MB resident)"},
                          "gpu_launches": args.steps}))
    if world_size > 1:
        dist.destroy_process_group()


def run_crowd_eval(args):
    """BASELINE config 4: 4-human crowd evaluation, 256 agents (64 scenes x 4) per GPU, scenes sharded over ranks with
    no collective. One step = one vector step of every agent (policy forward + 4 agent-by-agent env sub-steps in the
    reference's update order). Secondary workload: prints its own JSON line."""
    import torch
    import torch.distributed as dist
    from egogen_b200.main_crowd_eval import build_crowd_world, crowd_start_data
    from egogen_b200.ppo_policy import Batch
    world_size = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=dev)
    S, A = 64, 4
    w = build_crowd_world(dev, S, A, sequential=True, seed=rank)
    w["policy"].eval()
    wp, goals, betas = crowd_start_data(w["sampler"], S, A, dev, seed=rank)
    venv, pol = w["venv"], w["policy"]

    def vector_step():
        with torch.no_grad():
            out = pol.forward(Batch(obs=venv.observation()))
            venv.step(out.act)

    venv.reset_from(torch.arange(S * A), wp, goals, betas)
    for _ in range(max(args.warmup, 3)):
        vector_step()
    venv.reset_from(torch.arange(S * A), wp, goals, betas)
    if world_size > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        vector_step()
    e1.record(); torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t.item())
        print(json.dumps({"metric": "crowd eval agent-steps/sec", "value": S * A * world_size * args.steps / (ms / 1e3),
                          "unit": "env-steps/s", "n_gpus": world_size, "steps": args.steps, "warmup": max(args.warmup, 3),
                          "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "dtype": "f32",
                          "data": "synthetic",
                          "config": {"workload": "4-human crowd eval (main_crowd_eval.py): 64 scenes x 4 agents per GPU, agents see "
                                                 "each other as holes of the floor polygon, agent-by-agent update order"}}))
    if world_size > 1:
        dist.destroy_process_group()


def run_cvae_train(args):
    """BASELINE config 3: C-VAE marker-predictor training on synthetic canonicalised primitives, batch 4096, 200-frame
    sequences, max_rollout 8 (8 chained primitives per optimiser step), Adam 5e-4. Secondary workload."""
    import torch
    from egogen_b200.train_gamma_predictor import GAMMAPrimitiveVAETrainOP, SyntheticPrimitiveBatchGen
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    B = 4096
    op = GAMMAPrimitiveVAETrainOP(trainconfig={"batch_size": B, "max_rollout": 8}, device=dev)
    op.build_model(seed=0)
    gen = SyntheticPrimitiveBatchGen(B, 200, dev, seed=0)
    data = gen.next_batch_with_jts(B)
    for _ in range(max(args.warmup, 3)):
        op.calc_loss_rollout(data, 0); op.optimizer_step(5e-4)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss, _ = op.calc_loss_rollout(data, 0); op.optimizer_step(5e-4)
    e1.record(); torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    print(json.dumps({"metric": "C-VAE training primitives/sec", "value": B * 8 * args.steps / (ms / 1e3), "unit": "primitives/s",
                      "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
                      "higher_is_better": True, "dtype": "f32", "data": "synthetic", "loss": loss,
                      "config": {"workload": "C-VAE predictor training, batch 4096 x 8-primitive rollout (200-frame sequences), Adam"}}))


def run_regressor_train(args):
    """Marker -> body regressor training (GAMMARegressorTrainOP, models_GAMMA_primitive.py:594-710): 512 sequences x 16
    frames = 8192 bodies per optimiser step (forward over 3 recurrences, SMPL-X markers, backward through SMPL-X, Adam).
    Secondary workload."""
    import torch
    from egogen_b200 import assets
    from egogen_b200.smplx_parser import get_lbs_model
    from egogen_b200.train_gamma_regressor import GAMMARegressorTrainOP, SyntheticBodyMarkerBatchGen
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    n_seq, n_frames = 512, 16
    op = GAMMARegressorTrainOP(trainconfig={"batch_size": n_seq}, device=dev)
    op.build_model(seed=0)
    gen = SyntheticBodyMarkerBatchGen(get_lbs_model("male", dev, marker_vids=assets.marker_ids()), n_seq, n_frames, dev, seed=0)
    betas, mk = gen.next_batch_genderselection(n_seq)
    mk = mk.reshape(-1, 201).contiguous(); betas = betas.reshape(-1, 10).contiguous()
    for _ in range(max(args.warmup, 3)):
        op.forward_loss_backward(mk, betas); op.optimizer_step(3e-4)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        _, loss, _ = op.forward_loss_backward(mk, betas); op.optimizer_step(3e-4)
    e1.record(); torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    print(json.dumps({"metric": "regressor training bodies/sec", "value": mk.shape[0] * args.steps / (ms / 1e3), "unit": "bodies/s",
                      "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
                      "higher_is_better": True, "dtype": "f32", "data": "synthetic", "loss": loss,
                      "config": {"workload": "body-regressor training, 512 sequences x 16 frames per step, SMPL-X marker loss, Adam"}}))


def run_combo_train(args):
    """Joint predictor + regressor objective (GAMMAPrimitiveComboTrainOP.calc_loss_one, :819-838): 1024 primitives per
    optimiser step (18 432 regressed bodies in the SMPL-X cycle loss). Secondary workload."""
    import torch
    from egogen_b200.train_gamma_combo import GAMMAPrimitiveComboTrainOP
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    B = 1024
    op = GAMMAPrimitiveComboTrainOP(trainconfig={"batch_size": B}, device=dev)
    op.build_model(seed=0)
    g = torch.Generator().manual_seed(0)
    ref = (torch.cumsum(torch.randn(20, B, 201, generator=g) * 0.02, dim=0) + torch.randn(1, B, 201, generator=g) * 0.3).to(dev)
    betas = (torch.randn(1, B, 10, generator=g) * 0.5).expand(20, B, 10).contiguous().to(dev)
    for _ in range(max(args.warmup, 3)):
        op.calc_loss_one([betas, ref], 0); op.optimizer_step(1e-4)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss, _, _ = op.calc_loss_one([betas, ref], 0); op.optimizer_step(1e-4)
    e1.record(); torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    print(json.dumps({"metric": "combo training primitives/sec", "value": B * args.steps / (ms / 1e3), "unit": "primitives/s",
                      "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
                      "higher_is_better": True, "dtype": "f32", "data": "synthetic", "loss": loss,
                      "config": {"workload": "predictor + regressor combo objective, 1024 primitives per step (18432 bodies), Adam on the predictor"}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", type=str, default="ppo", choices=["ppo", "ego_depth", "cvae_train", "crowd_eval", "regressor_train", "combo_train"],
                    help="ppo = headline (BASELINE config 2); ego_depth = secondary config-5 sweep")
    args = ap.parse_args()
    if args.workload == "ego_depth" and args.impl == "ours":
        return run_ego_depth(args)
    if args.workload == "cvae_train" and args.impl == "ours":
        return run_cvae_train(args)
    if args.workload == "crowd_eval" and args.impl == "ours":
        return run_crowd_eval(args)
    if args.workload == "regressor_train" and args.impl == "ours":
        return run_regressor_train(args)
    if args.workload == "combo_train" and args.impl == "ours":
        return run_combo_train(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
