// Shared helpers for the egogen_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <string>

#include "../../include/egogen_b200.h"

namespace eg {

extern thread_local char g_last_error[512];
extern std::atomic<int64_t> g_launch_count;

inline int set_error(int code, const char* fmt, const char* a = "", const char* b = "") {
  snprintf(g_last_error, sizeof(g_last_error), fmt, a, b);
  return code;
}

#define EG_CUDA_CHECK(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      snprintf(eg::g_last_error, sizeof(eg::g_last_error), "%s:%d %s -> %s", __FILE__,   \
               __LINE__, #expr, cudaGetErrorString(_e));                                 \
      return EG_ERR_CUDA;                                                                \
    }                                                                                    \
  } while (0)

#define EG_REQUIRE(cond, msg)                                                            \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      snprintf(eg::g_last_error, sizeof(eg::g_last_error), "%s:%d invalid argument: %s", \
               __FILE__, __LINE__, msg);                                                 \
      return EG_ERR_INVALID_ARG;                                                         \
    }                                                                                    \
  } while (0)

// every kernel launch goes through this so eg_launch_count() is the library's own claim
#define EG_LAUNCH(kernel, grid, block, smem, stream, ...)                                \
  do {                                                                                   \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                          \
    eg::g_launch_count.fetch_add(1, std::memory_order_relaxed);                          \
    EG_CUDA_CHECK(cudaGetLastError());                                                   \
  } while (0)

// Programmatic dependent launch for the short kernels that sit between the dense layers: the kernel may be scheduled
// while its predecessor drains; eg_pdl_enter() (FIRST statement of such a kernel) releases this kernel's own dependents and
// then blocks until the predecessor grid has completed and its writes are visible. A kernel launched this way without
// eg_pdl_enter() would race with its producer - only kernels that call it may use EG_LAUNCH_PDL.
__device__ __forceinline__ void eg_pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
#define EG_LAUNCH_PDL(kernel, grid_, block_, smem_, stream_, ...)                          \
  do {                                                                                   \
    cudaLaunchConfig_t _cfg{};                                                           \
    _cfg.gridDim = dim3(grid_); _cfg.blockDim = dim3(block_);                            \
    _cfg.dynamicSmemBytes = (smem_); _cfg.stream = (stream_);                            \
    cudaLaunchAttribute _at[1];                                                          \
    _at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                      \
    _at[0].val.programmaticStreamSerializationAllowed = 1;                               \
    _cfg.attrs = _at; _cfg.numAttrs = 1;                                                 \
    cudaError_t _le = cudaLaunchKernelEx(&_cfg, kernel, __VA_ARGS__);                    \
    eg::g_launch_count.fetch_add(1, std::memory_order_relaxed);                          \
    EG_CUDA_CHECK(_le);                                                                  \
  } while (0)

// Optional device-side timing of one named kernel class (bench.py roofline): CUDA events recorded on the
// launching stream around the kernel; eg_profile_read() synchronises and sums them.
struct ProfSlot { cudaEvent_t a, b; int64_t units; };
void prof_begin(cudaStream_t st, int64_t units);
void prof_end(cudaStream_t st);
void stage_mark(cudaStream_t st, int id);   // env-step stage boundaries (eg_stage_profile_*)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
inline int dev_alloc_copy(T** dst, const T* src_host, size_t n) {
  EG_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(dst), n * sizeof(T)));
  EG_CUDA_CHECK(cudaMemcpy(*dst, src_host, n * sizeof(T), cudaMemcpyHostToDevice));
  return EG_OK;
}

constexpr int kNumSMs = 148;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- SDF sampling shared by sdf.cu and lbs.cu (ATen grid_sampler_3d semantics) ----------------
struct SdfGrid {
  const float* grid;
  int D0, D1, D2;
  const float* center;  // device [3]
  const float* scale;   // device [1]
  // optional conservative coarse grid (eg_sdf_prepare): coarse[c] = min of (-grid) over every fine corner a
  // trilinear sample inside the 8^3-cell c can touch; > 0 means no sample in the cell can be negative
  const float* coarse = nullptr;
  int C0 = 0, C1 = 0, C2 = 0;
  const uint32_t* coarse_bits = nullptr;   // 1 bit per coarse cell (coarse <= 0), n_bit_words words
  int n_bit_words = 0;
  const uint32_t* fine_bits = nullptr;     // same for 2^3 cells ([ceil(D/2)]^3 bits, dilated), global memory
  // exact per-cell class of the trilinear sample's sign, one bit pair per fine cell (x0, y0, z0) = floor of the sample
  // index, 32 cells per uint2: .x bit = every in-range corner of the cell is > 0 (the sample IS a penetration),
  // .y bit = corners of both signs / zero / tiny (the sample must be evaluated); neither = no corner is > 0 (never)
  const uint2* cell_class = nullptr;
};
constexpr int kCoarseShift = 3;

// host-side lookup of the coarse grid registered for `grid` by eg_sdf_prepare (fills g.coarse / C*)
void sdf_attach_coarse(SdfGrid& g);

// un-normalise with align_corners=False then clamp to [0, D-1] (padding_mode='border').
// Written with explicit round-to-nearest intrinsics so ptxas cannot contract the sequence into
// FMAs: the floor of this value is a bit-exact parity target.
__device__ __forceinline__ float sdf_unnormalize(float p, int D) {
  float i = __fmul_rn(__fadd_rn(p, 1.0f), (float)D);
  i = __fmul_rn(__fsub_rn(i, 1.0f), 0.5f);
  return fminf(fmaxf(i, 0.0f), (float)(D - 1));
}

// returns -trilinear(grid, p) with p the world-space point; writes the base corner indices.
__device__ __forceinline__ float sdf_sample_point(const SdfGrid& g, float cx, float cy, float cz,
                                                  float s, float x, float y, float z, int& ox,
                                                  int& oy, int& oz) {
  const float px = __fmul_rn(__fsub_rn(x, cx), s);
  const float py = __fmul_rn(__fsub_rn(y, cy), s);
  const float pz = __fmul_rn(__fsub_rn(z, cz), s);
  const float ix = sdf_unnormalize(px, g.D0);
  const float iy = sdf_unnormalize(py, g.D1);
  const float iz = sdf_unnormalize(pz, g.D2);
  const float fx0 = floorf(ix), fy0 = floorf(iy), fz0 = floorf(iz);
  const int x0 = (int)fx0, y0 = (int)fy0, z0 = (int)fz0;
  ox = x0; oy = y0; oz = z0;
  // ATen weights: (corner+1 - i) and (i - corner); accumulate in ATen's corner order
  // (x here is ATen's z/D axis, z is ATen's x/W axis): tnw,tne,tsw,tse,bnw,bne,bsw,bse.
  const float wx1 = __fsub_rn(ix, fx0), wx0 = __fsub_rn(__fadd_rn(fx0, 1.0f), ix);
  const float wy1 = __fsub_rn(iy, fy0), wy0 = __fsub_rn(__fadd_rn(fy0, 1.0f), iy);
  const float wz1 = __fsub_rn(iz, fz0), wz0 = __fsub_rn(__fadd_rn(fz0, 1.0f), iz);
  const bool x1ok = x0 + 1 <= g.D0 - 1, y1ok = y0 + 1 <= g.D1 - 1, z1ok = z0 + 1 <= g.D2 - 1;
  const int64_t sx = (int64_t)g.D1 * g.D2, sy = g.D2;
  const float* b = g.grid + x0 * sx + y0 * sy + z0;
  float acc = 0.0f;
  // ATen: weight = (wW * wH) * wD with W = our z, H = our y, D = our x
  acc = __fadd_rn(acc, __fmul_rn(__ldg(b), __fmul_rn(__fmul_rn(wz0, wy0), wx0)));
  if (z1ok) acc = __fadd_rn(acc, __fmul_rn(__ldg(b + 1), __fmul_rn(__fmul_rn(wz1, wy0), wx0)));
  if (y1ok) acc = __fadd_rn(acc, __fmul_rn(__ldg(b + sy), __fmul_rn(__fmul_rn(wz0, wy1), wx0)));
  if (y1ok && z1ok) acc = __fadd_rn(acc, __fmul_rn(__ldg(b + sy + 1), __fmul_rn(__fmul_rn(wz1, wy1), wx0)));
  if (x1ok) {
    acc = __fadd_rn(acc, __fmul_rn(__ldg(b + sx), __fmul_rn(__fmul_rn(wz0, wy0), wx1)));
    if (z1ok) acc = __fadd_rn(acc, __fmul_rn(__ldg(b + sx + 1), __fmul_rn(__fmul_rn(wz1, wy0), wx1)));
    if (y1ok) acc = __fadd_rn(acc, __fmul_rn(__ldg(b + sx + sy), __fmul_rn(__fmul_rn(wz0, wy1), wx1)));
    if (y1ok && z1ok) acc = __fadd_rn(acc, __fmul_rn(__ldg(b + sx + sy + 1), __fmul_rn(__fmul_rn(wz1, wy1), wx1)));
  }
  return -acc;
}

// out-of-line exact sign query for rare paths (one copy of the 8-corner sample instead of one per call site)
static __device__ __noinline__ bool sdf_exact_negative(const float* grid, int D0, int D1, int D2, float cx, float cy, float cz,
                                                float s, float x, float y, float z) {
  SdfGrid g{grid, D0, D1, D2, nullptr, nullptr};
  int i0, i1, i2;
  return sdf_sample_point(g, cx, cy, cz, s, x, y, z, i0, i1, i2) < 0.0f;
}

// value of the conservative coarse cell containing the world point (<= 0: the cell may hold a negative sample);
// -1 when no coarse grid is attached, so callers fall through to the exact sample.
__device__ __forceinline__ float sdf_coarse_value(const SdfGrid& g, float cx, float cy, float cz, float s, float x,
                                                  float y, float z) {
  if (g.coarse == nullptr) return -1.0f;
  const float ix = sdf_unnormalize(__fmul_rn(__fsub_rn(x, cx), s), g.D0);
  const float iy = sdf_unnormalize(__fmul_rn(__fsub_rn(y, cy), s), g.D1);
  const float iz = sdf_unnormalize(__fmul_rn(__fsub_rn(z, cz), s), g.D2);
  const int c0 = (int)ix >> kCoarseShift, c1 = (int)iy >> kCoarseShift, c2 = (int)iz >> kCoarseShift;
  return __ldg(g.coarse + (c0 * g.C1 + c1) * g.C2 + c2);
}

// linear index of the conservative coarse cell containing the world point
__device__ __forceinline__ int sdf_coarse_index(const SdfGrid& g, float cx, float cy, float cz, float s, float x,
                                                float y, float z) {
  const float ix = sdf_unnormalize(__fmul_rn(__fsub_rn(x, cx), s), g.D0);
  const float iy = sdf_unnormalize(__fmul_rn(__fsub_rn(y, cy), s), g.D1);
  const float iz = sdf_unnormalize(__fmul_rn(__fsub_rn(z, cz), s), g.D2);
  const int c0 = (int)ix >> kCoarseShift, c1 = (int)iy >> kCoarseShift, c2 = (int)iz >> kCoarseShift;
  return (c0 * g.C1 + c1) * g.C2 + c2;
}

// sign-only query used by the fused penetration count: identical to sdf_sample_point(...) < 0 (the trilinear
// sample is a non-negative combination of the cell's corners, so a positive coarse minimum proves "not negative"
// without touching the fine grid); falls back to the exact sample otherwise.
__device__ __forceinline__ bool sdf_is_negative(const SdfGrid& g, float cx, float cy, float cz, float s, float x,
                                                float y, float z) {
  if (g.coarse != nullptr) {
    const float ix = sdf_unnormalize(__fmul_rn(__fsub_rn(x, cx), s), g.D0);
    const float iy = sdf_unnormalize(__fmul_rn(__fsub_rn(y, cy), s), g.D1);
    const float iz = sdf_unnormalize(__fmul_rn(__fsub_rn(z, cz), s), g.D2);
    const int c0 = (int)ix >> kCoarseShift, c1 = (int)iy >> kCoarseShift, c2 = (int)iz >> kCoarseShift;
    if (__ldg(g.coarse + ((int64_t)c0 * g.C1 + c1) * g.C2 + c2) > 0.0f) return false;
  }
  int ox, oy, oz;
  return sdf_sample_point(g, cx, cy, cz, s, x, y, z, ox, oy, oz) < 0.0f;
}

}  // namespace eg
