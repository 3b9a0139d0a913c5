"""Time eg_linear_forward (tcgen05 3xTF32 path vs the fp32 SIMT tiles, EG_GEMM_TC=0) on the path's layer shapes."""
import os, sys, torch
sys.path.insert(0, ".")
from egogen_b200 import _lib
dev = torch.device("cuda:0")
shapes = [(256, 1152, 1152), (256, 256, 1152), (256, 1536, 512), (5120, 512, 512), (4096, 768, 256), (4096, 256, 512), (1024, 1152, 1152)]
lib = _lib.lib()
for M, N, K in shapes:
    x = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    y = torch.empty(M, N, device=dev)
    def run():
        _lib.check(lib.eg_linear_forward(_lib.ptr(x), K, M, _lib.ptr(W), _lib.ptr(b), K, N, 3, 0.01, None, 0, _lib.ptr(y), N, _lib.stream_ptr(dev)))
    for _ in range(5): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    ref = torch.nn.functional.leaky_relu(x.double() @ W.double().t() + b.double(), 0.01)
    err = (y.double() - ref).abs().max().item()
    print(f"EG_GEMM_TC={os.environ.get('EG_GEMM_TC','1')} M={M:5d} N={N:5d} K={K:5d}: {us:8.1f} us  {2.0*M*N*K/us/1e6:7.2f} TFLOP/s  max err {err:.2e}")
