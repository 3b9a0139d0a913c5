"""CPU emulation of the tensor-core regressor arithmetic (fp16 hi/lo split, per-MMA fp32 accumulate with a chosen
rounding of the accumulator) against float64, to see which error source dominates. Test infrastructure only."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import nets
from egogen_b200 import assets

torch.manual_seed(0)
reg = nets.RegressorOracle()
assets.fill_params_(reg, seed=12, w_gain=0.7)
with torch.no_grad():
    reg.pnet.out_fc.weight.mul_(0.3)
sd = {k: v.double().numpy() for k, v in reg.state_dict().items()}
nb, nrec = 10, 3
M = 512
rng = np.random.default_rng(1)
markers = rng.standard_normal((M, 201)) * 0.4
betas = rng.standard_normal((M, 10)) * 0.5

def rz32(x):
    y = x.astype(np.float32)
    over = np.abs(y.astype(np.float64)) > np.abs(x)
    y[over] = np.nextafter(y[over], np.float32(0))
    return y.astype(np.float64)

def rn32(x):
    return x.astype(np.float32).astype(np.float64)

def split16(x):
    hi = x.astype(np.float16).astype(np.float64)
    lo = (x - hi).astype(np.float16).astype(np.float64)
    return hi, lo

def scale_of(W):
    m = np.abs(W).max()
    e = int(np.floor(np.log2(m)))
    return 2.0 ** (13 - e)

def tc_matmul(A, W, mode, nacc=1):
    """A [M,K] fp32-valued, W [N,K]. mode: 'rz' / 'rn' accumulate rounding per 16-k MMA. nacc hi accumulators (k-steps dealt round-robin)."""
    s = scale_of(W)
    Wh, Wl = split16(W * s)
    Ah, Al = split16(A)
    K = A.shape[1]
    Kp = (K + 15) // 16 * 16
    pad = lambda X: np.pad(X, ((0, 0), (0, Kp - K)))
    Ah, Al, Wh, Wl = pad(Ah), pad(Al), pad(Wh), pad(Wl)
    rnd = rz32 if mode == "rz" else rn32
    hi = [np.zeros((A.shape[0], W.shape[0])) for _ in range(nacc)]
    corr = np.zeros((A.shape[0], W.shape[0]))
    for i, k in enumerate(range(0, Kp, 16)):
        sl = slice(k, k + 16)
        a = i % nacc
        hi[a] = rnd(hi[a] + Ah[:, sl] @ Wh[:, sl].T)
        corr = rnd(corr + Ah[:, sl] @ Wl[:, sl].T)
        corr = rnd(corr + Al[:, sl] @ Wh[:, sl].T)
    tot = corr
    for a in range(nacc):
        tot = rn32(tot + hi[a])
    return rn32(tot / s)

def fp32_matmul(A, W, mode=None, nacc=1):
    return (A.astype(np.float32) @ W.astype(np.float32).T).astype(np.float64)

def exact(A, W, mode=None, nacc=1):
    return A @ W.T

def forward(mm, mode=None, nacc=1, f32=True):
    r = rn32 if f32 else (lambda x: x)
    Win, bin_ = sd["pnet.in_fc.weight"], sd["pnet.in_fc.bias"]
    xb = np.zeros((M, 159))
    base = r(mm(np.concatenate([markers, betas], 1), np.concatenate([Win[:, :201], Win[:, 360:]], 1), mode, nacc) + bin_)
    for rec in range(nrec):
        h = base if rec == 0 else r(mm(xb, Win[:, 201:360], mode, nacc) + base)
        for b in range(nb):
            t = np.maximum(r(mm(h, sd[f"pnet.layers.{b}.layers.0.weight"], mode, nacc) + sd[f"pnet.layers.{b}.layers.0.bias"]), 0)
            h = r(np.maximum(r(mm(t, sd[f"pnet.layers.{b}.layers.1.weight"], mode, nacc) + sd[f"pnet.layers.{b}.layers.1.bias"]), 0) + h)
        xb = r(r(mm(h, sd["pnet.out_fc.weight"], mode, nacc) + sd["pnet.out_fc.bias"]) + xb)
    return xb

ref = forward(exact, f32=False)
def report(name, xb):
    aa_ref = nets.RegressorOracle.cont2aa(torch.from_numpy(ref)).numpy()
    aa = nets.RegressorOracle.cont2aa(torch.from_numpy(xb)).numpy()
    print(f"{name:28s} max|dxb|={np.abs(xb - ref).max():.2e}  max|dYb|={np.abs(aa - aa_ref).max():.2e}  |h|-scale xb max={np.abs(ref).max():.2f}")
report("fp32 sgemm", forward(fp32_matmul))
report("tc rn, 1 acc", forward(tc_matmul, "rn", 1))
report("tc rz, 1 acc", forward(tc_matmul, "rz", 1))
report("tc rz, 2 acc", forward(tc_matmul, "rz", 2))
report("tc rz, 4 acc", forward(tc_matmul, "rz", 4))

# ---- harsher hardware model: every addend (16 products + accumulator) is aligned to the largest exponent of the
# instruction and truncated toward zero with G guard bits below the fp32 ulp of that exponent, then summed exactly
def make_tc_addend(G, debias=0.0):
    def mm(A, W, mode, nacc=1):
        s = scale_of(W)
        Wh, Wl = split16(W * s)
        Ah, Al = split16(A)
        K = A.shape[1]
        Kp = (K + 15) // 16 * 16
        pad = lambda X: np.pad(X, ((0, 0), (0, Kp - K)))
        Ah, Al, Wh, Wl = pad(Ah), pad(Al), pad(Wh), pad(Wl)
        def mma(acc, a, w):
            prod = a[:, None, :] * w[None, :, :]                      # [M,N,16] exact in float64
            allv = np.concatenate([prod, acc[:, :, None]], axis=2)
            mx = np.abs(allv).max(axis=2, keepdims=True)
            e = np.floor(np.log2(np.maximum(mx, 1e-300)))
            q = 2.0 ** (e - 23 - G)
            tr = np.trunc(allv / q) * q
            ssum = tr.sum(axis=2)
            return rz32(ssum)
        hi = np.zeros((A.shape[0], W.shape[0])); corr = np.zeros_like(hi)
        for k in range(0, Kp, 16):
            sl = slice(k, k + 16)
            hi = mma(hi, Ah[:, sl], Wh[:, sl])
            corr = mma(corr, Ah[:, sl], Wl[:, sl])
            corr = mma(corr, Al[:, sl], Wh[:, sl])
        return rn32(rn32(hi * (1.0 + debias) + corr) / s)
    return mm
M = 256
markers = markers[:M]; betas = betas[:M]
ref = forward(exact, f32=False)
report("fp32 sgemm (M=256)", forward(fp32_matmul))
report("tc rz 1 acc (M=256)", forward(tc_matmul, "rz", 1))
for G in (0, 1, 2, 3):
    report(f"addend-trunc G={G}", forward(make_tc_addend(G)))
report("addend-trunc G=0 debias", forward(make_tc_addend(0, 8 * 0.5 * 0.72 * 2.0 ** -23)))

# ---- hypothesis: the tensor core flushes fp16 subnormal operands to zero
def ftz16(x16):
    y = x16.copy()
    y[np.abs(y) < 2.0 ** -14] = 0.0
    return y
def make_tc_ftz(act_scale=1.0):
    def mm(A, W, mode, nacc=1):
        s = scale_of(W)
        Wh, Wl = split16(W * s)
        Ah, Al = split16(A * act_scale)
        Ah, Al, Wh, Wl = ftz16(Ah), ftz16(Al), ftz16(Wh), ftz16(Wl)
        return rn32((Ah @ Wh.T + Ah @ Wl.T + Al @ Wh.T) / (s * act_scale))
    return mm
report("ftz, no act scale", forward(make_tc_ftz(1.0)))
report("ftz, act scale 2^8", forward(make_tc_ftz(256.0)))
