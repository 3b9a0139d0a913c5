// tcgen05 / TMEM / TMA mainloop of the full-mesh LBS vertex kernel (sm_100a).
//
//   D_c[v, n] = sum_k basisT_c[v, k] * F[n, k]      c in {x,y,z}, v = 128 vertices, n = 128 bodies, k = 576
//
// A operand: basisT [3][n_pad][KT] fp32 (planar x/y/z rows, K-major), pre-rounded to TF32 on the host.
// B operand: F [N_pad][KT] fp32 written by the prep kernel (TF32-rounded pose features; the shape
//            coefficients are split hi/lo against hi/lo shape rows of the basis so the shape blend keeps
//            ~fp32 accuracy - see build in lbs.cu).
// Per CTA (persistent, 1 per SM): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 4..19 =
// epilogue. smem: ring of 2 stages x (3 A tiles 128x32 + 1 B tile 128x32, SWIZZLE_128B) = 128 KB, plus
// 2 x 48 KB staging of per-body joint transforms for the skinning epilogue.
// Accumulators: 3 x 128 fp32 columns of TMEM; the epilogue reads its lane (= vertex) with tcgen05.ld,
// applies skinning (+transl, optional store, optional world transform + SDF sample + penetration count).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace eg {
namespace tc {

constexpr int KT = 576;            // padded contraction: 486 pose + 3 x 20 shape (hi*hi, lo*hi, hi*lo) + 30 zero
constexpr int BKT = 32;            // k-chunk per stage: 32 tf32 = 128 B = one swizzle atom row
constexpr int NCHUNK = KT / BKT;   // 18
constexpr int TV = 128;            // vertices per tile (UMMA M)
constexpr int TB = 128;            // bodies per tile (UMMA N)
constexpr int STAGES = 2;
constexpr int A_TILE_BYTES = TV * BKT * 4;   // 16 KB
constexpr int B_TILE_BYTES = TB * BKT * 4;   // 16 KB
constexpr int STAGE_BYTES = 3 * A_TILE_BYTES + B_TILE_BYTES;   // 64 KB
constexpr int EPI_WARPS = 16;
constexpr int THREADS = 128 + EPI_WARPS * 32;   // 640
constexpr int TMEM_COLS = 512;
constexpr int CHUNK_B = 16;                           // bodies per staged joint-transform chunk
constexpr int ASTAGE_BYTES = CHUNK_B * 64 * 48;        // up to 64 joints x 12 floats per body = 48 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers, tmem ptr*/ + 2 * ASTAGE_BYTES;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// K-major SWIZZLE_128B shared-memory descriptor (cute::UMMA::SmemDescriptor): LBO = 1, SBO = 1024 B, version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                 // leading byte offset (16 B units)
  d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}
// cute::UMMA::InstrDescriptor for kind::tf32, fp32 accumulate, K-major A and B, M=128, N=TB
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TB >> 3) << 17) | ((uint32_t)(TV >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(kIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
  uint32_t r0, r1, r2, r3;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7) : "r"(taddr));
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
  v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ void cp_async16_u32(uint32_t smem_addr, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_addr), "l"(gmem));
}

// Tile order: vertex tiles are walked in bands of BAND_V tiles (BAND_V x 884 KB of basis ~ 25 MB) with every body
// tile visited inside a band before the next band starts, so the basis band stays L2-resident while it is reused
// by all body tiles (the full 72.5 MB basis plus features/transforms does not survive a full sweep in L2).
constexpr int BAND_V = 28;
__device__ __forceinline__ void tile_coords(int tile, int n_vt, int n_bt, int& vt, int& bt) {
  const int full = BAND_V * n_bt;
  const int band = tile / full;
  const int rem = tile - band * full;
  const int w = min(BAND_V, n_vt - band * BAND_V);
  bt = rem / w;
  vt = band * BAND_V + (rem - bt * w);
}

}  // namespace tc
}  // namespace eg
