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


@pytest.mark.parametrize("M,N,K,ta,tb", [
    (256, 1152, 1152, 0, 0),      # dX = dY W: A K-major, B (the weight [n_out, k_in] read along k_in) MN-major
    (1152, 1152, 256, 1, 0),      # dW = dY^T X: both operands MN-major, contraction over the 256-row batch
    (256, 1152, 256, 1, 0),       # dW of the actor head
    (1536, 402 + 2, 256, 1, 0),   # GRU input weights (ragged N)
    (300, 200, 136, 0, 0),        # ragged tails, B MN-major
    (200, 136, 300, 1, 0),        # ragged tails, both MN-major
    (4096, 512, 256, 0, 0),
])
def test_gemm_tc_backward_layouts(M, N, K, ta, tb):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M + 3 * N + 5 * K + ta)
    A = torch.randn(M, K, generator=g).to(dev)
    B = (torch.randn(K, N, generator=g) / K ** 0.5).to(dev)
    ref = A.double() @ B.double()
    A_st = A.t().contiguous() if ta else A.contiguous()             # storage the layout flag describes
    B_st = B.t().contiguous() if tb else B.contiguous()
    C0 = torch.randn(M, N, generator=g).to(dev)
    for acc in (0, 1):
        Cm = C0.clone()
        _lib.check(_lib.lib().eg_matmul(_lib.ptr(A_st), A_st.shape[1], ta, _lib.ptr(B_st), B_st.shape[1], tb, M, N, K,
                                        _lib.ptr(Cm), N, acc, _lib.stream_ptr(dev)))
        want = ref + (C0.double() if acc else 0.0)
        err = (Cm.double() - want).abs().max().item()
        assert err <= 1e-5 * max(want.abs().max().item(), 1.0), (acc, err)


def test_matmul_edge_cases_and_fallback_layouts():
    """Empty products are no-ops, null pointers are rejected with EG_ERR_INVALID_ARG, and shapes the tensor-core path
    does not take (unaligned pitch, tiny K) still give the right answer through the SIMT tiles."""
    dev = torch.device("cuda:0")
    lib = _lib.lib()
    A = torch.randn(8, 70, device=dev); B = torch.randn(70, 40, device=dev); Cm = torch.full((8, 40), 7.0, device=dev)
    assert lib.eg_matmul(_lib.ptr(A), 70, 0, _lib.ptr(B), 40, 0, 0, 40, 70, _lib.ptr(Cm), 40, 0, _lib.stream_ptr(dev)) == 0
    assert bool((Cm == 7.0).all())                                    # M = 0: untouched
    assert lib.eg_matmul(None, 70, 0, _lib.ptr(B), 40, 0, 8, 40, 70, _lib.ptr(Cm), 40, 0, _lib.stream_ptr(dev)) < 0
    assert b"invalid argument" in lib.eg_last_error()
    # pitch 70 is not 16-byte aligned and K = 6 is below the tensor-core threshold: SIMT fallback
    for (M, N, K) in [(8, 40, 70), (33, 65, 6), (256, 130, 402)]:
        A = torch.randn(M, K, device=dev); B = torch.randn(K, N, device=dev); Cm = torch.empty(M, N, device=dev)
        _lib.check(lib.eg_matmul(_lib.ptr(A), K, 0, _lib.ptr(B), N, 0, M, N, K, _lib.ptr(Cm), N, 0, _lib.stream_ptr(dev)))
        ref = A.double() @ B.double()
        assert (Cm.double() - ref).abs().max().item() <= 1e-5 * max(ref.abs().max().item(), 1.0)
