// Dense-layer building blocks shared by the C-VAE / regressor / VPoser / policy operators.
// The parity bar is 1e-4 relative through 60-layer chains, so single-pass TF32 is not used: nn.Linear-forward layers
// run the 3xTF32 split on tcgen05 (gemm_tc.cu, ~2^-21 relative), every other layout the fp32 SIMT tiles below.
#pragma once
#include "common.cuh"

namespace eg {

enum Act { ACT_NONE = 0, ACT_TANH = 1, ACT_RELU = 2, ACT_LRELU = 3 };

// C[M,N] (ldc) = act( A[M,K] * B[K,N] + bias[N] + (beta ? C : 0) ) + residual[M,N]
//   TA=false: A element (m,k) at A[(m / a_div) * lda + k]      (row-major activations)
//   TA=true : A element (m,k) at A[k * lda + m]                (transposed, for dW = dY^T X)
//   TB=true : B element (k,n) at B[n * ldb + k]                (nn.Linear weight [out,in])
//   TB=false: B element (k,n) at B[k * ldb + n]
struct GemmArgs {
  const float* A; int lda; int a_div;
  const float* B; int ldb;
  float* C; int ldc;
  const float* bias;
  const float* residual; int ldr;
  int M, N, K;
  int act; float slope;
  int beta;           // 1: accumulate into existing C (before activation)
  float alpha;        // scale on the product
};

int launch_gemm(const GemmArgs& g, bool TA, bool TB, cudaStream_t st);
// tensor-core path (gemm_tc.cu) for every layout but TA && TB: EG_OK = launched, 1 = not eligible, < 0 = error
int launch_gemm_tc(const GemmArgs& g, bool TA, bool TB, cudaStream_t st);
void gemm_tc_set_enabled(int on);

// y = act(x W^T + b) (+ residual) for an nn.Linear weight W[out,in]
inline int linear(cudaStream_t st, const float* x, int ldx, int M, const float* W, int ldw,
                  const float* b, int in_dim, int out_dim, float* y, int ldy, int act = ACT_NONE,
                  float slope = 0.01f, const float* residual = nullptr, int ldr = 0, int beta = 0,
                  int a_div = 1) {
  GemmArgs g{x, ldx, a_div, W, ldw, y, ldy, b, residual, ldr, M, out_dim, in_dim, act, slope, beta, 1.0f};
  return launch_gemm(g, false, true, st);
}

// PyTorch GRU gate math on pre-computed gi = x W_ih^T + b_ih, gh = h W_hh^T + b_hh ([M,3H], order r,z,n).
// gh == nullptr means h == 0 (gh = b_hh broadcast).  h_out may alias h_in.
int launch_gru_gate(cudaStream_t st, const float* gi, const float* gh, const float* b_hh,
                    const float* h_in, float* h_out, int M, int H, int ld_h_out,
                    float* r_save = nullptr, float* z_save = nullptr, float* n_save = nullptr,
                    float* ghn_save = nullptr);

// dst[r, 0..ld_dst) = src[r, 0..cols) zero-padded: re-pitches rows to a multiple of 16 bytes so TMA can stage them
int launch_pad_rows(cudaStream_t st, const float* src, int ld_src, int rows, int cols, float* dst, int ld_dst);

}  // namespace eg
