"""Is the PPO iteration launch-bound? CPU enqueue time (no sync) vs device time (events) of collect and learn."""
import sys
import time

import torch

sys.path.insert(0, ".")
from egogen_b200.runtime import build_world

dev = torch.device("cuda:0")
w = build_world(dev, 256, seed=0, sdf_res=256)
col, pol = w["collector"], w["policy"]
pol.train(); col.reset()
for _ in range(3):
    b, _ = col.collect(1024); pol.learn(b, 256, 1)
torch.cuda.synchronize()
ev = lambda: torch.cuda.Event(enable_timing=True)
K = 6
cc = cl = gc = gl = 0.0
for _ in range(K):
    e0, e1, e2 = ev(), ev(), ev()
    torch.cuda.synchronize()
    e0.record(); t0 = time.perf_counter(); b, _ = col.collect(1024); t1 = time.perf_counter(); e1.record()
    torch.cuda.synchronize()
    t2 = time.perf_counter(); pol.learn(b, 256, 1); t3 = time.perf_counter(); e2.record()
    torch.cuda.synchronize()
    cc += (t1 - t0) * 1e3; cl += (t3 - t2) * 1e3; gc += e0.elapsed_time(e1); gl += e1.elapsed_time(e2)
# learn with the CPU far ahead of the device (a 10 ms spin kernel queued first): pure device time of the learn chain
gl2 = 0.0
for _ in range(K):
    b, _ = col.collect(1024)
    torch.cuda.synchronize()
    e1, e2 = ev(), ev()
    torch.cuda._sleep(20_000_000)
    e1.record(); pol.learn(b, 256, 1); e2.record()
    torch.cuda.synchronize()
    gl2 += e1.elapsed_time(e2)
print(f"learn with the launch queue pre-filled: device {gl2 / K:.2f} ms")
print(f"collect: cpu enqueue {cc / K:.2f} ms, device {gc / K:.2f} ms | learn: cpu enqueue {cl / K:.2f} ms, device {gl / K:.2f} ms")
