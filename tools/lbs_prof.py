"""Stall accounting of the tcgen05 LBS kernel inside the benchmark's 256-env rollout (debug build only:
EG_NVCC_EXTRA=-DEG_LBS_PROF=1 python -m egogen_b200.build --force). Prints, per CTA averages: epilogue warps' wait on the
table / accumulators / work clocks, the MMA warp's waits, the finish-time spread over CTAs and the per-vertex-tile cost."""
import ctypes as C
import sys

import numpy as np
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
    col.collect(1024)
cta = np.zeros((160, 48), dtype=np.uint64); vt = np.zeros(512, dtype=np.uint64)
lib.eg_lbs_prof_dump(None, None, 1)
K = 4
for _ in range(K):
    col.collect(1024)
lib.eg_lbs_prof_dump(cta.ctypes.data_as(C.c_void_p), vt.ctypes.data_as(C.c_void_p), 0)
cta = cta[:148].astype(np.float64)
L = K * 4                                   # launches
print("per launch and CTA, kclk (mean over CTAs | min | max):")
def row(name, x):
    x = x / L / 1e3
    print(f"  {name:34s} {x.mean():9.1f} | {x.min():9.1f} | {x.max():9.1f}")
row("epilogue wait table (mean of 8 warps)", cta[:, 0:8].mean(1))
row("epilogue wait accumulators", cta[:, 8:16].mean(1))
row("epilogue work", cta[:, 16:24].mean(1))
for wi in range(8):
    row(f"   warp {wi + 4} work", cta[:, 16 + wi])
row("mma wait accumulator-free", cta[:, 24])
row("mma wait operands", cta[:, 25])
row("mma issue block", cta[:, 32])
row("producer wait ring slot", cta[:, 26])
row("table producer wait buffer", cta[:, 27])
print("tiles per CTA per launch: mean %.1f min %.0f max %.0f" % ((cta[:, 30] / L).mean(), (cta[:, 30] / L).min(), (cta[:, 30] / L).max()))
fin = cta[:, 28]
print("last launch: epilogue finish spread over CTAs: max - min = %.1f us, max - mean = %.1f us" % ((fin.max() - fin.min()) / 1e3, (fin.max() - fin.mean()) / 1e3))
v = vt.astype(np.float64); v = v[v > 0] / (L * 40) / 1e3
print("per vertex tile epilogue work of warp 4 (kclk per tile visit): mean %.1f min %.1f max %.1f, n=%d" % (v.mean(), v.min(), v.max(), len(v)))
print("  by vertex tile:", " ".join(f"{x:.0f}" for x in v))
