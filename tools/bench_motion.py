import sys, torch
sys.path.insert(0, ".")
from egogen_b200.models_gamma_primitive import GAMMAPrimitiveComboGenOP
dev = torch.device("cuda:0")
g = GAMMAPrimitiveComboGenOP(testconfig={"gpu_index": 0}); m = g.build_model(seed=0)
B = 256
X = torch.randn(B, 2, 201, device=dev) * 0.3; z = torch.randn(B, 128, device=dev); betas = torch.zeros(B, 10, device=dev)
for fused in (True, False):
    m.set_fused(fused)
    for _ in range(3): m.sample_prior_env_major(X, 402, 201, z, betas, B)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): m.sample_prior_env_major(X, 402, 201, z, betas, B)
    e1.record(); torch.cuda.synchronize()
    print("fused" if fused else "layerwise", e0.elapsed_time(e1) / 10, "ms per sample_prior (B=256)")
