"""Where does one PPO iteration spend its device time? CUDA-event sections around collect / learn and the
stage profiler inside eg_env_step (warm caches, real launch conditions - unlike ncu's cold serialised replay)."""
import ctypes as C
import sys
import time

import torch

sys.path.insert(0, ".")
from egogen_b200 import _lib
from egogen_b200.runtime import build_world

dev = torch.device("cuda:0")
w = build_world(dev, 256, seed=0, sdf_res=256)
col, pol = w["collector"], w["policy"]
pol.train(); col.reset()
lib = _lib.lib()
for _ in range(3):
    b, _ = col.collect(1024); pol.learn(b, 256, 1)
torch.cuda.synchronize()
ev = lambda: torch.cuda.Event(enable_timing=True)
tc = tl = 0.0
lib.eg_stage_profile_enable(1)
K = 5
t0 = time.perf_counter()
for _ in range(K):
    e0, e1, e2 = ev(), ev(), ev()
    e0.record(); b, _ = col.collect(1024); e1.record(); pol.learn(b, 256, 1); e2.record()
    torch.cuda.synchronize()
    tc += e0.elapsed_time(e1); tl += e1.elapsed_time(e2)
wall = (time.perf_counter() - t0) * 1e3 / K
ms = (C.c_double * 16)()
_lib.check(lib.eg_stage_profile_read(ms, 16))
lib.eg_stage_profile_enable(0)
names = {1: "cvae decode + regressor", 2: "param blend", 3: "fused LBS+SDF (prep+tc+compact+finish)", 4: "vposer",
         5: "rewards + recanon", 6: "seed-joint LBS", 7: "ego-sensing"}
print(f"per iteration: wall {wall:.2f} ms | collect {tc / K:.2f} ms | learn {tl / K:.2f} ms")
for k, n in names.items():
    print(f"  env stage {k} {n:42s} {ms[k] / K:8.3f} ms / iteration ({ms[k] / K / 4:.3f} per vector step)")
print(f"  env stages total {sum(ms[1:8]) / K:.3f} ms / iteration")
