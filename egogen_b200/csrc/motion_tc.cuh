// tcgen05 forms of the two long layer chains of GAMMAPrimitiveCombo.sample_prior (motion_tc.cu):
//   * the 18-step GRUCell + MLP decode of the C-VAE predictor (reference models_GAMMA_primitive.py:83-101)
//   * the 3 x 22-layer marker -> body regressor                (reference models_GAMMA_primitive.py:160-175, 222-259)
// Both run every product as FP16 hi/lo-split tensor-core passes (x = hi + lo, hi*hi + hi*lo + lo*hi, fp32 accumulation in
// TMEM: ~2^-22 relative, the accuracy class of the 3xTF32 dense layers at half the operand bytes and twice the rate).
#pragma once
#include "common.cuh"

namespace eg {
namespace mtc {

struct MotionTc;   // prepared (scaled, hi/lo-split, slice-ordered) weights + activation exchange buffers

// w: the EgMotion weight table (device pointers, order documented in include/egogen_b200.h)
int create(MotionTc** out, const EgMotionDims& d, const float* const* w, cudaStream_t st);
int refresh(MotionTc* m, const float* const* w, cudaStream_t st);      // after the weights changed in place
void destroy(MotionTc* m);

bool decode_available(const MotionTc* m);       // dims match and a 16-CTA cluster can be scheduled
bool regress_available(const MotionTc* m);
// 18 decode steps: gi1 [B][3H] = first-step GRU input term (b_ih + [hx,z,y0] W_ih^T), h0 [B][H] = drnn_mlp(hx),
// Y [B][20][D] holds the 2 history frames on entry and receives frames 2..19
int decode(MotionTc* m, const float* gi1, const float* h0, float* Y, int B, cudaStream_t st);
// regressor over frames 2..19 of every env: xbc [B*20][159] rows of frames 2..19 are written (cont. 6-D body params)
int regress(MotionTc* m, const float* Y, const float* betas, int B, float* xbc, cudaStream_t st);

}  // namespace mtc
}  // namespace eg
