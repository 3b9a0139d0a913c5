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
The default (ppo) line also carries `secondary`: BASELINE configs 1, 3, 4 and 5 measured on the same GPUs, each with its
own clocks / roofline / cpu_baseline (`--workload X` runs one of them alone; `--no-secondary` skips them).
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
B_LBS_ENV_STEP = 20 * B_BODY   # LBS bytes per env step (20 bodies)
FLOP_BODY = 49.1e6      # dense-reference FLOPs per body
FLOP_BLEND_BODY = 2.0 * (486 + 20) * 3 * 10475     # useful pose + shape blend contraction per body (31.8 MFLOP)
DTYPE = "f32 (tensor-core products on fp16 / tf32 split operands, f32 accumulate)"


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
    # bounded sample with the GPU arm's definition of a step: a collect of n_steps vector steps followed by GAE and one
    # learn pass of FOUR minibatches (one optimiser step each), scaled from 256 envs to 16
    n_envs, n_steps, n_mb = 16, 4, 4
    for i in range(args.warmup):
        harness.run_iteration(world, 2, 1, True, seed=100 + i, n_minibatch=1)
    tot_s, tot_n = 0.0, 0
    for i in range(args.steps):
        s, n = harness.run_iteration(world, n_envs, n_steps, True, seed=i, n_minibatch=n_mb)
        tot_s += s; tot_n += n
    v = tot_n / tot_s
    sample = (f"{n_envs} envs x {n_steps} vector steps + GAE + {n_mb} minibatch updates of {n_envs * n_steps // n_mb} rows per step "
              f"(the GPU arm: 256 envs x 4 vector steps + 4 x 256 rows), sequential per-env loop with the 4x duplicated batch")
    cfg = config_dict(args.gpus)
    cfg["reference_sample"] = {"envs": n_envs, "vector_steps": n_steps, "minibatches": n_mb,
                               "rows_per_minibatch": n_envs * n_steps // n_mb, "warmup_sample": "2 envs x 1 vector step"}
    line = {"impl": "reference", "metric": "crowd_ppo env-steps/sec", "value": v, "unit": "env-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def config_dict(n):
    return {"workload": "PPO phase-1 training, 256 parallel envs per GPU, synthetic random-box scene (256^3 SDF), "
                        "surrogate SMPL-X (V=10475), step = collect 1024 transitions + 4 minibatches of 256",
            "envs_per_gpu": N_ENVS, "step_per_collect_per_gpu": STEP_PER_COLLECT, "batch_per_gpu": BATCH,
            "global_batch": BATCH * n, "parallelism": f"env-sharded dp{n}",
            "cache": "L2 flushed (256 MiB write) between timed iterations; working set ~250 MB > 126 MB L2"}


class Ctx:
    """One process per GPU: rank / device / process group, set up once per bench.py run."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback exists)"
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world_size > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self._flush = None

    def flush_l2(self):
        if self._flush is None:
            self._flush = self.torch.empty(256 << 20, dtype=self.torch.uint8, device=self.dev)
        self._flush.zero_()

    def barrier(self):
        if self.world_size > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, ms):
        t = self.torch.tensor([float(ms)], device=self.dev, dtype=self.torch.float64)
        if self.world_size > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def time_steps(self, fn, steps, flush=True):
        """EXACTLY `steps` calls of fn, each bracketed by (L2 flush,) barrier + synchronize and CUDA events on the current
        stream; returns (sum of ms, max over ranks; this rank's per-step list)."""
        torch = self.torch
        ms = []
        for _ in range(steps):
            if flush:
                self.flush_l2()
            self.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize(self.dev)
            ms.append(e0.elapsed_time(e1))
        return self.max_over_ranks(sum(ms)), ms

    def close(self):
        if self.world_size > 1:
            self.dist.destroy_process_group()


class clocked:
    """`with clocked(ctx) as c: ...` -> c.result = nvidia-smi clock record of rank 0's GPU during the block."""

    def __init__(self, ctx):
        self.s = ClockSampler(ctx.local) if ctx.rank == 0 else None
        self.result = None

    def __enter__(self):
        if self.s:
            self.s.start()
        return self

    def __exit__(self, *a):
        if self.s:
            self.result = self.s.stop()
        return False


def tensor_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p)).get("bf16_tflops", 1590.0)), "measured"
    return 1590.0, "fallback"


def measure_ppo(ctx, args):
    """Headline workload (BASELINE config 2). Returns the JSON line (rank 0) or None."""
    import numpy as np
    torch, dist = ctx.torch, ctx.dist
    from egogen_b200 import _lib
    from egogen_b200.collector import Collector
    from egogen_b200.runtime import build_world
    world_size, rank, dev = ctx.world_size, ctx.rank, ctx.dev
    torch.manual_seed(1234 + rank)
    np.random.seed(1234 + rank)
    lib = _lib.lib()
    w = build_world(dev, N_ENVS, seed=rank, sdf_res=256)
    col, pol = w["collector"], w["policy"]
    pol.train()
    col.reset()
    last = {}

    def iteration(c):
        batch, _ = c.collect(STEP_PER_COLLECT)
        last["batch"] = batch
        return pol.learn(batch, BATCH, 1)

    for _ in range(max(args.warmup, 3)):
        iteration(col)
    torch.cuda.synchronize(dev)
    ncu_range = os.environ.get("EG_NCU_RANGE") == "1"      # ncu --profile-from-start off: capture the timed region only
    with clocked(ctx) as ck:
        lib.eg_profile_enable(1)
        launches0 = lib.eg_launch_count()
        if ncu_range:
            torch.cuda.cudart().cudaProfilerStart()
        total_ms, ms = ctx.time_steps(lambda: iteration(col), args.steps)
        if ncu_range:
            torch.cuda.synchronize(dev)
            torch.cuda.cudart().cudaProfilerStop()
        launches = lib.eg_launch_count() - launches0
        tot_ms, n_l, n_units = C.c_double(), C.c_int64(), C.c_int64()
        _lib.check(lib.eg_profile_read(C.byref(tot_ms), C.byref(n_l), C.byref(n_units)))
        lib.eg_profile_enable(0)
        value = STEP_PER_COLLECT * world_size * args.steps / (total_ms / 1e3)
        # ---- e2e: same iteration with the reference's host-side data path ---------------------
        col2 = Collector(pol, w["venv"], host_boundary=True)
        col2.reset()
        for _ in range(max(args.warmup, 3)):               # the same warm-up as the device-resident arm
            iteration(col2)
        torch.cuda.synchronize(dev)
        col2.h2d_bytes = col2.d2h_bytes = 0
        e2e_ms, e2e_list = ctx.time_steps(lambda: iteration(col2), args.steps)
    e2e = {"value": STEP_PER_COLLECT * world_size * args.steps / (e2e_ms / 1e3), "unit": "env-steps/s",
           "h2d_bytes_per_step": col2.h2d_bytes // args.steps, "d2h_bytes_per_step": (col2.d2h_bytes + 5 * 8 * 4) // args.steps,
           "ms_per_step_rank0": [round(x, 3) for x in e2e_list]}
    # sanity (outside every timed region): the rollouts and the updated policy are finite
    assert bool(torch.isfinite(last["batch"].returns).all()) and bool(torch.isfinite(pol.flat_params).all()), \
        "non-finite values after the timed iterations"
    # ---- where the step goes (untimed extra iterations with the env-stage profiler on) -----------
    # two untimed passes: collect / learn split without any instrumentation, then the env stages with the stage profiler on
    # (its per-stage events cost launch overlap, so its figures are shares of the env step, not of the timed iteration)
    stage = None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    K = 2
    tc = tl = 0.0
    for _ in range(K):
        torch.cuda.synchronize(dev)
        ev[0].record(); b_, _ = col.collect(STEP_PER_COLLECT); ev[1].record(); pol.learn(b_, BATCH, 1); ev[2].record()
        torch.cuda.synchronize(dev)
        tc += ev[0].elapsed_time(ev[1]); tl += ev[1].elapsed_time(ev[2])
    if rank == 0:
        lib.eg_stage_profile_enable(1)
    for _ in range(K):
        b_, _ = col.collect(STEP_PER_COLLECT); pol.learn(b_, BATCH, 1)
    torch.cuda.synchronize(dev)
    if rank == 0:
        sm = (C.c_double * 16)()
        _lib.check(lib.eg_stage_profile_read(sm, 16))
        lib.eg_stage_profile_enable(0)
        names = {1: "cvae_decode_regressor", 2: "param_blend", 3: "lbs_sdf", 4: "vposer", 5: "rewards_recanon",
                 6: "seed_joint_lbs", 7: "ego_sensing"}
        stage = {"collect_ms": tc / K, "learn_ms": tl / K, "env_stage_ms_instrumented": {n: sm[k] / K for k, n in names.items()}}
    if rank != 0:
        return None
    # ---- roofline of the path's headline kernel (fused LBS + SDF vertex kernel) ------------------
    peak, how = peaks()
    k_ms = tot_ms.value / max(n_l.value, 1)
    bodies_per_launch = n_units.value / max(n_l.value, 1)
    hbm_achieved = bodies_per_launch * B_BODY / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    # useful blend FLOPs: [bodies, 486 pose + 20 shape] x [506, 3 V]; the kernel issues K = 576 (hi/lo shape columns + padding)
    tf_achieved = bodies_per_launch * FLOP_BLEND_BODY / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
    tf_issued = bodies_per_launch * 2.0 * 576 * 3 * 10475 / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
    tf_peak, tf_how = tensor_peak()
    traffic, traffic_src = None, None
    for name in ("r2_lbs_tc_traffic.json", "r1_lbs_tc_traffic.json"):
        tp = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) * bodies_per_launch / tj["bodies"]
            traffic_src = tj["source"]
            break
    step_gbs = value / world_size * B_LBS_ENV_STEP / 1e9
    roofline = {"bound": "tensor", "achieved": tf_achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": tf_achieved / tf_peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": tf_how + " dense bf16/fp16 cuBLAS burst (kind::f16 operands, fp32 accumulation)",
                "kernel": "lbs_verts_tc_kernel<FUSE_SDF> (tcgen05 kind::f16, fp32 accumulate in TMEM)",
                "avg_launch_ms": k_ms, "bodies_per_launch": bodies_per_launch, "launches_timed": n_l.value,
                "kernel_share_of_step": tot_ms.value / float(sum(ms)),
                "flops_per_body": FLOP_BLEND_BODY, "issued_tflops_incl_k_padding": tf_issued,
                "hbm_contract": {"achieved": hbm_achieved, "peak": peak, "unit": "GB/s", "frac": hbm_achieved / peak,
                                 "note": "SURVEY 8(d) contract figure: bodies x 127636 B (what an unfused LBS must move); "
                                         "the fused kernel itself writes only the per-body counts"},
                "step": {"lbs_bytes_per_env_step": B_LBS_ENV_STEP, "achieved": step_gbs, "peak": peak, "unit": "GB/s",
                         "frac": step_gbs / peak, "frac_of_nominal_8TBs": step_gbs / 8000.0,
                         "note": "whole-iteration LBS-bytes roofline per GPU: env-steps/s x 2 552 720 B; the iteration is a chain of "
                                 "latency-bound dense layers and a tensor-bound LBS kernel, not an HBM stream", "where": stage}}
    # ---- CPU baseline: the oracle port on a bounded sample, host cores of this box --------------
    cpu = None
    if world_size == 1:
        from oracle import harness
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        ow = harness.build_oracle_world(0, sdf_res=256)
        harness.run_iteration(ow, 8, 1, False, seed=99)                       # warm-up
        s_, n_ = harness.run_iteration(ow, 16, 2, False, seed=0, n_minibatch=4)
        cpu = {"value": n_ / s_, "unit": "env-steps/s", "cores": cores, "kind": "port",
               "sample": "oracle (batched, dup removed): 16 envs x 2 vector steps + GAE + 4 minibatch updates, torch-CPU all threads"}
    return {"metric": "crowd_ppo env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world_size,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "config": config_dict(world_size), "e2e": e2e, "gpu_launches": int(launches), "clocks": ck.result,
            "roofline": roofline, "cpu_baseline": cpu}


# ---------------------------------------------------------------------------------------------------------------------------
# secondary workloads (BASELINE configs 1, 3, 4, 5): each returns one dict with its own clocks / roofline / cpu_baseline
# ---------------------------------------------------------------------------------------------------------------------------
def measure_single_env(ctx, steps, warmup):
    """BASELINE config 1: single-agent evaluation, 1 env, deterministic policy (main_ppo.py --watch): latency per env step.
    Every rank runs its own env (replicas, no collective); value = env-steps/s of all ranks."""
    torch = ctx.torch
    from egogen_b200.ppo_policy import Batch
    from egogen_b200.runtime import build_world
    w = build_world(ctx.dev, 1, seed=100 + ctx.rank, sdf_res=256)
    venv, pol = w["venv"], w["policy"]
    pol.eval()
    pol._deterministic_eval = True
    venv.reset()

    def step():
        with torch.no_grad():
            out = pol.forward(Batch(obs=venv.observation()))
            _, _, term, _, _ = venv.step(out.act)
            venv.reset_masked(term)
    for _ in range(max(warmup, 3)):
        step()
    n_inner = 16
    with clocked(ctx) as ck:
        launches0 = _launches()
        total_ms, _ = ctx.time_steps(lambda: [step() for _ in range(n_inner)], steps)
        launches = _launches() - launches0
    ms_step = total_ms / (steps * n_inner)
    peak, how = peaks()
    gbs = B_LBS_ENV_STEP / (ms_step * 1e-3) / 1e9
    out = {"config_id": 1, "metric": "single-env eval env-steps/sec", "value": ctx.world_size * 1e3 / ms_step, "unit": "env-steps/s",
           "latency_ms_per_env_step": ms_step, "n_gpus": ctx.world_size, "steps": steps * n_inner, "scaling": "replicas",
           "higher_is_better": True, "dtype": DTYPE, "data": "synthetic", "gpu_launches": int(launches), "clocks": ck.result,
           "config": {"workload": "single-agent evaluation (main_ppo.py --watch), 1 env per GPU, deterministic policy, 256^3 SDF; "
                                  "16 env steps per timed call, L2 flushed between calls"},
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None,
                        "peak_source": how, "kernel": "whole env step (LBS-bytes contract figure, 20 bodies x 127636 B)",
                        "note": "latency-bound by construction: one env is a chain of ~70 dependent launches on 20 bodies"}}
    if ctx.rank == 0 and ctx.world_size == 1:
        from oracle import harness
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        ow = harness.build_oracle_world(0, sdf_res=256)
        harness.run_eval_steps(ow, 1, 1, True, seed=9)
        s_, n_ = harness.run_eval_steps(ow, 1, 8, True, seed=0)
        out["cpu_baseline"] = {"value": n_ / s_, "unit": "env-steps/s", "cores": cores, "kind": "port",
                               "sample": "oracle in the reference's shape (1 env, 4x duplicated batch), 8 env steps"}
    return out if ctx.rank == 0 else None


def _launches():
    from egogen_b200 import _lib
    return _lib.lib().eg_launch_count()


def measure_cvae_train(ctx, steps, warmup):
    """BASELINE config 3: C-VAE marker-predictor training on synthetic canonicalised primitives, batch 4096, 200-frame
    sequences, max_rollout 8 (8 chained primitives per optimiser step), Adam 5e-4. Replicas (the reference trains on one GPU)."""
    torch = ctx.torch
    from egogen_b200.train_gamma_predictor import GAMMAPrimitiveVAETrainOP, SyntheticPrimitiveBatchGen
    dev = ctx.dev
    B = 4096
    op = GAMMAPrimitiveVAETrainOP(trainconfig={"batch_size": B, "max_rollout": 8}, device=dev)
    op.build_model(seed=0)
    gen = SyntheticPrimitiveBatchGen(B, 200, dev, seed=ctx.rank)
    data = gen.next_batch_with_jts(B)
    loss = [None]

    def step():
        loss[0], _ = op.calc_loss_rollout(data, 0); op.optimizer_step(5e-4)
    for _ in range(max(warmup, 3)):
        step()
    with clocked(ctx) as ck:
        launches0 = _launches()
        total_ms, _ = ctx.time_steps(step, steps)
        launches = _launches() - launches0
    prim = B * 8 * steps * ctx.world_size
    # algorithmic MACs per primitive, forward: x_enc 2 x 351k, e_rnn 18 x 351k, e_mlp 393k, mu/logvar 66k, drnn_mlp 328k,
    # decode 18 x (841 x 768 + 256 x 512 + 512 x 256 + 256 x 201) = 18 x 960k  -> 25.1 M; forward + backward = 3x
    flop_prim = 2.0 * 3.0 * 25.1e6
    tf = prim / ctx.world_size * flop_prim / (total_ms / 1e3) / 1e12
    tf_peak, tf_how = tensor_peak()
    out = {"config_id": 3, "metric": "C-VAE training primitives/sec", "value": prim / (total_ms / 1e3), "unit": "primitives/s",
           "n_gpus": ctx.world_size, "steps": steps, "ms_per_step": total_ms / steps, "scaling": "replicas", "higher_is_better": True,
           "dtype": DTYPE, "data": "synthetic", "loss": loss[0], "gpu_launches": int(launches), "clocks": ck.result,
           "config": {"workload": "C-VAE predictor training, batch 4096 x 8-primitive rollout (200-frame sequences), Adam; "
                                  "working set > L2, L2 flushed between steps"},
           "roofline": {"bound": "tensor", "achieved": tf, "peak": tf_peak, "unit": "TFLOP/s", "frac": tf / tf_peak, "traffic": None,
                        "peak_source": tf_how + " dense bf16 burst; the layers run 3xTF32 (3 tensor passes per useful FLOP, tf32 rate = 1/2)",
                        "kernel": "gemm_tc_kernel family (forward / dX / dW of the GRU + MLP layers)",
                        "flops_per_primitive": flop_prim}}
    if ctx.rank == 0 and ctx.world_size == 1:
        from oracle import harness
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        harness.run_cvae_train_steps(16, 200, 8, 1)
        s_, n_ = harness.run_cvae_train_steps(128, 200, 8, 2)
        out["cpu_baseline"] = {"value": n_ / s_, "unit": "primitives/s", "cores": cores, "kind": "port",
                               "sample": "oracle predictor + torch autograd + Adam: batch 128 x 8-primitive rollout, 2 steps"}
    return out if ctx.rank == 0 else None


def measure_crowd_eval(ctx, steps, warmup):
    """BASELINE config 4: 4-human crowd evaluation, 256 agents (64 scenes x 4) per GPU (2048 rollouts on 8 GPUs), scenes
    sharded over ranks with no collective. One step = one vector step of every agent (policy forward + 4 agent-by-agent
    env sub-steps in the reference's update order)."""
    torch = ctx.torch
    from egogen_b200.main_crowd_eval import build_crowd_world, crowd_start_data
    from egogen_b200.ppo_policy import Batch
    dev = ctx.dev
    S, A = 64, 4
    w = build_crowd_world(dev, S, A, sequential=True, seed=ctx.rank)
    w["policy"].eval()
    wp, goals, betas = crowd_start_data(w["sampler"], S, A, dev, seed=ctx.rank)
    venv, pol = w["venv"], w["policy"]

    def vector_step():
        with torch.no_grad():
            out = pol.forward(Batch(obs=venv.observation()))
            venv.step(out.act)
    venv.reset_from(torch.arange(S * A), wp, goals, betas)
    for _ in range(max(warmup, 3)):
        vector_step()
    venv.reset_from(torch.arange(S * A), wp, goals, betas)
    with clocked(ctx) as ck:
        launches0 = _launches()
        total_ms, _ = ctx.time_steps(vector_step, steps)
        launches = _launches() - launches0
    value = S * A * ctx.world_size * steps / (total_ms / 1e3)
    peak, how = peaks()
    gbs = value / ctx.world_size * B_LBS_ENV_STEP / 1e9
    out = {"config_id": 4, "metric": "crowd eval agent-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": ctx.world_size,
           "steps": steps, "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak", "dtype": DTYPE,
           "data": "synthetic", "gpu_launches": int(launches), "clocks": ck.result,
           "config": {"workload": "4-human crowd eval (main_crowd_eval.py): 64 scenes x 4 agents per GPU, agents see each other as "
                                  "holes of the floor polygon, agent-by-agent update order; L2 flushed between vector steps"},
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None,
                        "peak_source": how, "kernel": "whole vector step (LBS-bytes contract figure, 20 bodies x 127636 B per agent step)",
                        "note": "4 sequential sub-steps of 64 agents each: launch / latency bound"}}
    if ctx.rank == 0 and ctx.world_size == 1:
        from oracle import harness
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        harness.run_crowd_steps(1, 4, 1, seed=1)
        s_, n_ = harness.run_crowd_steps(2, 4, 3, seed=0)
        out["cpu_baseline"] = {"value": n_ / s_, "unit": "env-steps/s", "cores": cores, "kind": "port",
                               "sample": "oracle crowd env: 2 scenes x 4 agents x 3 vector steps, reference update order"}
    return out if ctx.rank == 0 else None


def measure_ego_depth(ctx, steps, warmup):
    """BASELINE config 5: ego-depth ray-march sweep, 8192 agents x 64x64 rays per GPU against a resident 256^3 SDF
    (agents sharded over ranks, grid replicated, no collective)."""
    torch = ctx.torch
    from egogen_b200 import assets, ego_depth
    dev, rank = ctx.dev, ctx.rank
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
    res = [None]

    def step():
        res[0] = ego_depth(sdf, cam, H, W, fx, fy, return_steps=True)
    for _ in range(max(warmup, 3)):
        step()
    with clocked(ctx) as ck:
        total_ms, ms = ctx.time_steps(step, steps)
    mean_steps = float(res[0][1].float().mean().item()) + 1.0       # samples per ray (the terminating sample included)
    peak, how = peaks()
    rays = A * H * W * ctx.world_size * steps
    gathered = A * H * W * mean_steps * 32.0 / (sum(ms) / steps / 1e3) / 1e9     # bytes gathered / s, this rank
    out = {"config_id": 5, "metric": "ego-depth rays/sec", "value": rays / (total_ms / 1e3), "unit": "rays/s", "n_gpus": ctx.world_size,
           "steps": steps, "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak", "dtype": "f32",
           "data": "synthetic", "gpu_launches": steps, "clocks": ck.result,
           "config": {"workload": "ego-depth sweep: 8192 agents x 64x64 rays per GPU, 256^3 SDF, <=64 sphere-trace steps, 7 m range; "
                                  "L2 flushed between sweeps"},
           "roofline": {"bound": "hbm", "achieved": gathered, "peak": peak, "unit": "GB/s", "frac": gathered / peak, "traffic": None,
                        "peak_source": how, "kernel": "ego_depth_kernel", "mean_samples_per_ray": mean_steps,
                        "note": "achieved = rays x samples x 32 B of corner gathers (L2-served; grid 67 MB resident)"}}
    if ctx.rank == 0 and ctx.world_size == 1:
        from oracle import harness
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sdf_cpu = {k: v.cpu() for k, v in sdf.items()}
        harness.run_ego_depth_cpu(sdf_cpu, cam[:4].cpu(), H, W, fx, fy)
        s_, n_ = harness.run_ego_depth_cpu(sdf_cpu, cam[:256].cpu(), H, W, fx, fy)
        out["cpu_baseline"] = {"value": n_ / s_, "unit": "rays/s", "cores": cores, "kind": "port",
                               "sample": "defining oracle (torch-CPU, all threads): 256 agents x 64x64 rays"}
    return out if ctx.rank == 0 else None


SECONDARY = {"single_env": measure_single_env, "cvae_train": measure_cvae_train, "crowd_eval": measure_crowd_eval,
             "ego_depth": measure_ego_depth}


def run_ours(args):
    ctx = Ctx()
    line = measure_ppo(ctx, args)
    if not args.no_secondary:
        sec = []
        for name, fn in SECONDARY.items():
            try:
                r = fn(ctx, max(2, min(args.steps, 5)), args.warmup)
            except Exception as e:                       # noqa: BLE001 - a secondary workload must not take the headline line down
                r = {"workload": name, "error": f"{type(e).__name__}: {e}"} if ctx.rank == 0 else None
            ctx.torch.cuda.empty_cache()
            if r is not None:
                sec.append(r)
        if line is not None:
            line["secondary"] = sec
    if line is not None:
        print(json.dumps(line))
    ctx.close()


def run_secondary(args):
    ctx = Ctx()
    r = SECONDARY[args.workload](ctx, args.steps, args.warmup)
    if r is not None:
        print(json.dumps(r))
    ctx.close()


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
    ap.add_argument("--workload", type=str, default="ppo",
                    choices=["ppo", "single_env", "ego_depth", "cvae_train", "crowd_eval", "regressor_train", "combo_train"],
                    help="ppo = headline (BASELINE config 2, with the secondary configs attached); the others run alone")
    ap.add_argument("--no-secondary", action="store_true", help="ppo workload only: skip BASELINE configs 1, 3, 4, 5")
    args = ap.parse_args()
    if args.workload in SECONDARY and args.impl == "ours":
        return run_secondary(args)
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
