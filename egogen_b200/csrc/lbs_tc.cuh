// tcgen05 / TMEM / TMA mainloop of the full-mesh LBS vertex kernel (sm_100a).
//
//   D_c[n, v] = sum_k F[n, k] * basisT_c[v, k]      c in {x,y,z}, n = 128 bodies (UMMA M, TMEM lanes),
//                                                   v = 80 vertices (UMMA N, TMEM columns), k = 576
//
// Operands are FP16 (kind::f16, fp32 accumulation in TMEM): fp16 carries the same 10-bit mantissa as TF32, so for
// every value in fp16's normal range the products are the ones the TF32 form produced, at half the operand bytes
// (the mainloop is bound by the L2->SM operand feed) and twice the tensor rate.
// A operand: F [N_pad][KT] fp16 written by the prep kernel (pose features; the shape coefficients are split hi/lo
//            against hi/lo shape rows of the basis so the shape blend keeps ~fp32 accuracy - see build in lbs.cu).
// B operand: basisT [3][n_pad_tc][KT] fp16 (planar x/y/z rows, K-major), rounded on the host, rows in
//            the JOINT-COHERENT vertex order built by eg_lbs_create (vertices sorted by their skinning-joint
//            tuple, tiles closed at 80 vertices or NJ_MAX distinct joints).
// Per CTA (persistent, 1 per SM): warp 0 = operand producer (TMA ring), warp 1 = MMA issuer (+TMEM alloc), warp 2 =
// table producer (bulk copies of the tile's joint-transform table and vertex records), warp 3 = tile scheduler (first
// tile = blockIdx, then tickets from a device counter, published to the other roles through a 4-slot ring - vertex
// tiles differ 4x in epilogue cost, so a static round-robin left a 50 us tail), warps 4..11 = epilogue.
// CTAs run as CLUSTERS OF 2 on the same vertex tile and adjacent body tiles: each CTA loads half of every basis tile
// and multicasts it to both, so the basis (the bulk of the L2->SM traffic) crosses the fabric once per pair.
// TMEM: TWO accumulator sets of 3 x 80 fp32 columns, so the MMA of tile i+1 runs under the epilogue of tile i.
// smem: ring of 2 stages x (3 basis tiles 80x64 + 1 feature tile 128x64 fp16, SWIZZLE_128B) = 92 KB, two 54 KB tables
// with the transforms of the tile's <= 9 joints for its 128 bodies (joint-major in HBM, so one 6 KB bulk copy per
// joint), two 3.75 KB record buffers, 4 KB of SDF coarse-cell sign bits, 12 KB of per-warp SDF work queues.
// Epilogue thread = one BODY (its TMEM lane): it walks the tile's vertices, keeps the four skinning-slot transforms
// of ITS body in registers and reloads a slot from the table only when the (warp-uniform) joint of that slot
// changes between consecutive vertices - runs of vertices that share joints cost no shared-memory traffic at all,
// where the vertex-per-lane form needed 12 ld.shared.v4 per (vertex, body).
// Fused SDF: the 8^3-cell sign bit (smem) is tested inline; a (vertex, body) whose cell may hold a negative sample is
// COMPACTED into the warp's shared-memory queue {x, y, z, body} instead of being resolved on the spot (a lane = a
// body, so on the spot one flagged body dragged the other 31 through the 2^3-cell lookup and the 8-corner sample:
// 11 and 9 of 32 lanes were active there and it was half of the epilogue's time). Whenever 32 entries are queued the
// warp resolves them with every lane busy: stage 1 = 2^3-cell bit (global), survivors re-queued; stage 2 = the exact
// trilinear sample, one atomic per negative sample. Queues persist across tiles and are drained at the end.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace eg {
namespace tc {

constexpr int KT = 576;            // padded contraction: 486 pose + 3 x 20 shape (hi*hi, lo*hi, hi*lo) + 30 zero
constexpr int BKT = 64;            // k-chunk per stage: 64 fp16 = 128 B = one swizzle atom row
constexpr int NCHUNK = KT / BKT;   // 9
constexpr int UMMA_K = 16;         // kind::f16
constexpr int TV = 80;             // vertices per tile (UMMA N)
constexpr int TB = 128;            // bodies per tile (UMMA M)
constexpr int STAGES = 2;
constexpr int CLUSTER = 1;          // CTAs per cluster: same vertex tile, adjacent body tiles; basis tiles are multicast
constexpr int HALF_ROWS = 80 / CLUSTER;   // basis rows each CTA of the pair loads (and multicasts) per component
constexpr int V_TILE_BYTES = TV * BKT * 2;   // 10 KB (basis, one component)
constexpr int HALF_BYTES = HALF_ROWS * BKT * 2;
constexpr int F_TILE_BYTES = TB * BKT * 2;   // 16 KB (features)
constexpr int STAGE_BYTES = 3 * V_TILE_BYTES + F_TILE_BYTES;   // 46 KB
constexpr int ACC_COLS = 3 * TV;                 // one accumulator set
constexpr int TMEM_COLS = 512;
constexpr int EPI_SUB = 2;                       // epilogue warps per TMEM lane quarter (each takes TV / EPI_SUB vertices)
constexpr int EPI_WARPS = 4 * EPI_SUB;
constexpr int THREADS = 128 + EPI_WARPS * 32;   // 384
constexpr int VPW = TV / EPI_SUB;                // vertices per epilogue warp and tile
constexpr int NJ_MAX = 9;                        // distinct skinning joints per vertex tile (tiles are closed earlier otherwise)
constexpr int SLOT_BYTES = TB * 48;              // one joint's transforms for the tile's 128 bodies
constexpr int TAB_BYTES = NJ_MAX * SLOT_BYTES;   // 60 KB
constexpr int REC_BYTES = TV * 48;               // the tile's per-vertex records
constexpr int NTAB = 2;                          // table / record buffers
constexpr int MASK_WORDS = 1024;                 // coarse-cell sign bits of the SDF grid (32^3 cells for a 256^3 grid)
constexpr int QCAP = 96;                         // entries of one epilogue warp's SDF queue pair (stage 1 grows up, stage 2 grows down)
constexpr int Q_BYTES = QCAP * 16;
constexpr int NSCHED = 2;                        // tile-scheduler ring slots
constexpr int BAR_BYTES = 256;
constexpr int OFF_BARS = STAGES * STAGE_BYTES;
constexpr int OFF_TAB = OFF_BARS + BAR_BYTES;
constexpr int OFF_REC = OFF_TAB + NTAB * TAB_BYTES;
constexpr int OFF_MASK = OFF_REC + NTAB * REC_BYTES;
constexpr int OFF_Q = OFF_MASK + MASK_WORDS * 4;
constexpr int SMEM_BYTES = OFF_Q + EPI_WARPS * Q_BYTES + 1024 /*align slack*/;
static_assert((2 * STAGES + 8 + 2 * NSCHED) * 8 + 4 + NSCHED * 4 <= BAR_BYTES, "barrier block");
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
static_assert(HALF_ROWS * CLUSTER == TV && HALF_BYTES % 1024 == 0, "multicast split");
static_assert(VPW % 4 == 0 && STAGE_BYTES % 1024 == 0 && V_TILE_BYTES % 1024 == 0, "tile geometry");

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
// for the single-lane producer roles, which run far ahead of their consumers: do not fight the epilogue warps of the
// same scheduler for issue slots while waiting
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(64);
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// multicast variant: the box lands at the same CTA-relative offset (and signals the same-offset mbarrier) in every
// CTA of `mask`
__device__ __forceinline__ void tma_load_2d_mc(uint32_t smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 1-D bulk copy global -> shared, completion counted on an mbarrier
__device__ __forceinline__ void bulk_load(uint32_t smem_dst, const void* gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
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
// cute::UMMA::InstrDescriptor for kind::f16 with fp16 A and B (a_format = b_format = 0), fp32 accumulate (c_format = 1),
// K-major A and B, M = TB, N = TV
// K-major A and B, M = TB, N = 3 * TV: the three component tiles of a stage are contiguous 80-row groups of one
// 240-row K-major tile, so ONE UMMA per k-step feeds all three accumulators and the feature tile (A) is fetched from
// shared memory once instead of three times (at N = 80 the operand fetch, 6.5 KB per 40-clock MMA, outran the 128 B/clk
// shared-memory port; at N = 240 it is 11.5 KB per 120 clocks)
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)((3 * TV) >> 3) << 17) | ((uint32_t)(TB >> 4) << 24);
static_assert((3 * TV) % 16 == 0 && 3 * TV <= 256, "UMMA N");

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
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
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t r;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
  return r;
}
__device__ __forceinline__ void cp_async16_u32(uint32_t smem_addr, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_addr), "l"(gmem));
}

// Tile order: vertex tiles are walked in bands of BAND_V tiles (BAND_V x 553 KB of basis ~ 24 MB) with every body
// tile visited inside a band before the next band starts, so the basis band stays L2-resident while it is reused
// by all body tiles (the full 72.5 MB basis plus features/transforms does not survive a full sweep in L2).
constexpr int BAND_V = 44;
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
