"""GPU parity of the tcgen05 3xTF32 dense layer (egogen_b200/csrc/gemm_tc.cu) through the C ABI (eg_linear_forward):
fp32-level accuracy against a float64 torch reference on the layer shapes of the path (policy MLP 256x1152x1152 with
split-k clusters, ragged tails, large-M C-VAE / VPoser shapes, fused activation + residual)."""
import ctypes as C

import pytest
import torch

from egogen_b200 import _lib

pytestmark = pytest.mark.gpu


def _linear(x, W, b, act=0, slope=0.01, residual=None):
    dev = x.device
    M, K = x.shape
    N = W.shape[0]
    y = torch.empty(M, N, device=dev, dtype=torch.float32)
    _lib.check(_lib.lib().eg_linear_forward(_lib.ptr(x), K, M, _lib.ptr(W), _lib.ptr(b), K, N, act, slope,
                                            _lib.ptr(residual), N if residual is not None else 0, _lib.ptr(y), N,
                                            _lib.stream_ptr(dev)))
    return y


@pytest.mark.parametrize("M,N,K,act,res", [
    (256, 1152, 1152, 3, True),      # policy MLP block layer: split-k over a 4-CTA cluster
    (256, 256, 1152, 0, False),      # actor head: 8-CTA cluster
    (256, 1536, 512, 0, False),      # GRU hidden projection
    (300, 200, 136, 1, False),       # ragged M / N / K tails (TMA zero fill, masked epilogue)
    (5120, 512, 512, 3, False),      # VPoser-sized rows: no split
    (4096, 768, 256, 1, True),       # C-VAE training shape
    (1, 64, 64, 0, False),           # single row
])
def test_gemm_tc_matches_float64(M, N, K, act, res):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    x = (torch.randn(M, K, generator=g) * 1.5).to(dev)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    r = torch.randn(M, N, generator=g).to(dev) if res else None
    y = _linear(x, W, b, act, 0.01, r)
    ref = x.double() @ W.double().t() + b.double()
    ref = {0: ref, 1: torch.tanh(ref), 2: torch.relu(ref), 3: torch.nn.functional.leaky_relu(ref, 0.01)}[act]
    if res:
        ref = ref + r.double()
    err = (y.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    # 3xTF32 drops the lo*lo term (2^-22 per product): an order of magnitude inside the path's 1e-4 relative bar
    assert err <= 1e-5 * max(scale, 1.0), (err, scale)
    # and no worse than the fp32 SIMT tiles would be: compare with torch's fp32 matmul error
    y32 = x @ W.t() + b
    y32 = {0: y32, 1: torch.tanh(y32), 2: torch.relu(y32), 3: torch.nn.functional.leaky_relu(y32, 0.01)}[act]
    if res:
        y32 = y32 + r
    err32 = (y32.double() - ref).abs().max().item()
    assert err <= 16 * err32 + 1e-6, (err, err32)
