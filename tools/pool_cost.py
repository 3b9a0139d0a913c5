"""Cost and frequency of the restart-pool refill inside the 256-env rollout."""
import sys, time
import torch
sys.path.insert(0, ".")
from egogen_b200.runtime import build_world
dev = torch.device("cuda:0")
w = build_world(dev, 256, seed=0, sdf_res=256)
col, pol, venv = w["collector"], w["policy"], w["venv"]
pol.train(); col.reset()
for _ in range(3):
    b, _ = col.collect(1024); pol.learn(b, 256, 1)
torch.cuda.synchronize()
t0 = time.perf_counter(); s = venv.sampler.next_body(2048); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"sampler.next_body(2048): {(t1 - t0) * 1e3:.1f} ms")
t0 = time.perf_counter(); venv._refill_pool(); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"_refill_pool(): {(t1 - t0) * 1e3:.1f} ms, rows {venv._pool_rows}")
r0 = venv.pool_refills
K = 60
ms = []
for _ in range(K):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    b, _ = col.collect(1024); pol.learn(b, 256, 1)
    torch.cuda.synchronize(); ms.append((time.perf_counter() - t0) * 1e3)
ms_sorted = sorted(ms)
print(f"{K} iterations: mean {sum(ms) / K:.2f} ms, median {ms_sorted[K // 2]:.2f}, max {ms_sorted[-1]:.1f}, refills {venv.pool_refills - r0}")
print("slowest:", [round(x, 1) for x in ms_sorted[-8:]])
