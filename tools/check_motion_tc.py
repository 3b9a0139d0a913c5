"""GPU check of the tcgen05 decode / regressor (motion_tc.cu): parity against the layer-by-layer path and timing.
Run under gpurun:  timeout 300 python tools/check_motion_tc.py"""
import os, sys, time
import torch
sys.path.insert(0, ".")
from egogen_b200.models_gamma_primitive import GAMMAPrimitiveComboGenOP

dev = torch.device("cuda:0")
g = GAMMAPrimitiveComboGenOP(testconfig={"gpu_index": 0}); m = g.build_model(seed=0)
print("EG_MOTION_TC =", os.environ.get("EG_MOTION_TC", "1"), flush=True)
gen = torch.Generator().manual_seed(5)
for B in (5, 37, 128, 256):
    X = (torch.randn(B, 2, 201, generator=gen) * 0.3).to(dev)
    z = torch.randn(B, 128, generator=gen).to(dev)
    betas = (torch.randn(B, 10, generator=gen) * 0.5).to(dev)
    m.set_fused(True)
    Y1, Yb1 = m.sample_prior_env_major(X, 402, 201, z, betas, B)
    torch.cuda.synchronize()
    m.set_fused(False)
    Y0, Yb0 = m.sample_prior_env_major(X, 402, 201, z, betas, B)
    torch.cuda.synchronize()
    dy = (Y1 - Y0).abs().max().item()
    db = (Yb1[:, 2:] - Yb0[:, 2:]).abs().max().item()
    print(f"B={B}: max|dY|={dy:.3e} (|Y|max {Y0.abs().max().item():.2f})  max|dYb|={db:.3e}  finite={bool(torch.isfinite(Y1).all())}", flush=True)
B = 256
X = torch.randn(B, 2, 201, device=dev) * 0.3; z = torch.randn(B, 128, device=dev); betas = torch.zeros(B, 10, device=dev)
for fused in (True, False):
    m.set_fused(fused)
    for _ in range(3): m.sample_prior_env_major(X, 402, 201, z, betas, B)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): m.sample_prior_env_major(X, 402, 201, z, betas, B)
    e1.record(); torch.cuda.synchronize()
    print("fused" if fused else "layerwise", e0.elapsed_time(e1) / 20, "ms per sample_prior (B=256)", flush=True)
