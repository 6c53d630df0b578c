// tcgen05 / TMEM / TMA dense contraction kernel for the relation head (sm_100a).
//
//   out[m,n] = epilogue( sum_k A[m,k] * B[n,k] ),  bf16 operands, fp32 accumulation in tensor memory.
//
// One persistent, warp-specialised kernel serves every dense stage of model.py:138-150,175:
//   plain GEMM      (1x1 convolutions as pixel GEMMs, fc1, fc2; SGB post_cat)
//   implicit conv   (3x3/pad 1 convolutions: per (64-channel block, kx) ONE 4-D TMA box of the NHWC activation tensor,
//                    8*MS+2 pixel rows tall and shifted by kx-1, is fetched; the three ky taps read it through
//                    descriptor start addresses 2048 B (= one 16-pixel row = two swizzle atoms) apart, so A is
//                    fetched 3x instead of 9x per channel block; TMA zero-fills the padding halo, so no im2col
//                    buffer ever exists)
//   block-sparse conv (the same 3x3 convolution restricted to a WORK LIST of 8 x {8,4}-pixel blocks: a 128-row sub-tile is
//                    assembled from 2 or 4 blocks of possibly different images, one 4-D TMA box per block and tap; used for
//                    conv3_1, whose output equals a weights-only background outside the dilated footprint of the two boxes)
//   K-cell-sparse GEMM (plain GEMM whose K axis is cut into <= 64 cells; every CTA M tile carries a 64-bit mask of the cells in which
//                    ANY of its rows is non-zero and visits only those K blocks - exact, because the skipped operand is zero; used
//                    for fc1 over the cells both boxes of a pair reach, with per-box fc1 rows gathered in the epilogue)
// CTA = 6 warps: warp 0 TMA producer, warp 1 tcgen05.mma issuer (+TMEM owner), warps 2-5 epilogue
// (TMEM -> registers -> fused bias/activation/2x2-max-pool -> global).  Pipelines: smem full/empty ring
// (TMA <-> MMA) and TMEM full/empty (MMA <-> epilogue) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// CTA tile = (MS*128) x BN: MS 128-row sub-tiles share every B stage (halves L2->smem weight traffic for MS=2).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "hc_common.cuh"

namespace hc {
namespace tc {

constexpr int BM = 128;          // UMMA M
constexpr int BK = 64;           // K per pipeline stage: 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;      // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
// CTA-pair kernels: 256 x 256 (or 2 x 256 x 256) fp32 accumulators per CTA fill TMEM, so the epilogue of a tile is NOT hidden behind the
// next tile's MMAs - it runs on EIGHT warps (two per TMEM lane quarter, alternating 32-column chunks) to halve its exposed time
constexpr int NUM_THREADS_CG2 = 320;
template <bool CG2> constexpr int cta_threads() { return CG2 ? NUM_THREADS_CG2 : NUM_THREADS; }
constexpr int TMEM_COLS = 512;
constexpr int A_SUB_BYTES = BM * BK * 2;   // 16 KB

template <int BN, int MS>
struct Cfg {
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = MS * A_SUB_BYTES + B_BYTES;
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 6 ? 6 : (200 * 1024) / STAGE_BYTES;
  // cta_group::2 pairs (block mode): each CTA of a pair stages its own A rows and HALF of every weight tile; the tensor cores of
  // both SMs read the other half from the peer's shared memory, so a stage costs MS*16 KB + B_BYTES/2 of L2->SM delivery per SM
  static constexpr int CG2_STAGE_BYTES = MS * A_SUB_BYTES + B_BYTES / 2;
  static constexpr int CG2_STAGES = (200 * 1024) / CG2_STAGE_BYTES > 6 ? 6 : (200 * 1024) / CG2_STAGE_BYTES;
  static constexpr int MAX_STAGES = 6;                 // barrier slots (the ring depth is a run-time choice: STAGES or CG2_STAGES)
  static constexpr int ACC_COLS = MS * BN;
  static constexpr int ACC_STAGES = TMEM_COLS / ACC_COLS;
  // implicit-conv "patch" pipeline: one A patch (8*MS+2 pixel rows x 16 pixels x 64 channels) per (channel block, kx) serves
  // the three ky taps through descriptor row offsets, so A is fetched 3 times per channel block instead of 9
  static constexpr int PATCH_ROWS = 8 * MS + 2;
  static constexpr int A_PATCH_BYTES = PATCH_ROWS * 16 * 128;        // multiple of 1024 (16 pixels x 128 B = 2 swizzle atoms per row)
  static constexpr int PA_SLOTS = 3;
  static constexpr int PB_SLOTS = 3;
  static constexpr int PATCH_DATA_BYTES = PA_SLOTS * A_PATCH_BYTES + PB_SLOTS * B_BYTES;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES > CG2_STAGES * CG2_STAGE_BYTES ? STAGES * STAGE_BYTES : CG2_STAGES * CG2_STAGE_BYTES;
  static constexpr int DATA_BYTES = RING_BYTES > PATCH_DATA_BYTES ? RING_BYTES : PATCH_DATA_BYTES;
  static constexpr int BAR_BYTES = (2 * MAX_STAGES + 2 * ACC_STAGES + 2 * PA_SLOTS + 2 * PB_SLOTS) * 8 + 16;
  static constexpr int SMEM_BYTES = DATA_BYTES + BAR_BYTES + 1024;  // +1024 alignment slack
  static_assert(A_PATCH_BYTES % 1024 == 0 && SMEM_BYTES <= 227 * 1024, "patch pipeline does not fit");
  static_assert(STAGES >= 2, "pipeline too shallow");
  static_assert(ACC_STAGES >= 1, "accumulators do not fit TMEM");
};

struct Params {
  int M, N, K;
  int mode;                      // HC_GEMM_PLAIN / HC_GEMM_CONV3
  int H, W, c_in, c_base;        // conv
  int tiles_x, tiles_y;          // conv: spatial tiles per image (16 wide, 8*MS tall)
  int tiles_m, tiles_n, group_m;
  int epi, act;
  long long ldc, c_off;
  const float* bias;
  const float* mul;              // optional elementwise multiplier [M, ld_mul] applied after bias/activation
  long long ld_mul;
  void* out;
  int patch;                     // 1 = implicit-conv patch pipeline (A fetched once per (channel block, kx)), 0 = plain GEMM
  const int* blocks;             // HC_GEMM_CONV3_BLOCKS: work list, entry = img << 8 | (y0/2) << 4 | (x0/2)
  const int* n_blocks;           // device scalar: entries in the work list
  int blk_h;                     // pixel rows per block (8 or 4)
  int blk_w;                     // pixel columns per block (8 or 4)
  int cl2;                       // block mode: tcgen05 cta_group::2 - CTA pairs (cluster of 2) on two M tiles of one N tile, UMMA M = 256,
                                 // each CTA stages half of every weight tile
  // K-cell-sparse plain GEMM: bit c of k_masks[CTA m tile] set = K blocks [c*k_cell_kb, (c+1)*k_cell_kb) are visited
  const unsigned long long* k_masks;
  int k_cell_kb;
  // EPI_BF16 row gathers: act(acc + bias + add_a[add_a_rows[row]] + add_b[add_b_rows[row]])
  const float* add_a; const int* add_a_rows;
  const float* add_b; const int* add_b_rows;
  long long ld_add;
  const int* out_rows;           // plain GEMM: output row of GEMM row r (NULL = r)
  // HC_EPI_POOL_DIFF_BF16 (block mode): out[pair_row[pair]] = (x - diff_sub[pair_sub[pair]]) - (diff_obj[pair_obj[pair]] - diff_bg)
  const __nv_bfloat16* diff_sub; const __nv_bfloat16* diff_obj; const __nv_bfloat16* diff_bg;
  const int* pair_sub; const int* pair_obj; const int* pair_row;
  int f16;                       // operand format: 0 = bf16 (default), 1 = IEEE fp16 (A, B, 16-bit outputs and the difference maps)
  const int* m_order;            // plain GEMM: visiting order of the M tiles (m_order[i] = M tile visited i-th; NULL = ascending)
  int dbg;                       // TIMING EXPERIMENTS ONLY (HC_TC_DEBUG, block mode; results are garbage): 1 = skip the A boxes, 2 = skip
                                 // the weight tile, 4 = skip the MMAs - splits a stage's time into its TMA and tensor-core parts
};

// ------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// cta_group::2 loads: the data lands in THIS CTA's shared memory, the byte count is credited to the barrier `bar`, which may live in
// the peer CTA (shared::cluster address) - both CTAs of a pair report to the leader's full barrier
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;      // shared::cluster address of the same offset in the EVEN CTA of the pair
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_cg2(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// arrive on a barrier that may live in the peer CTA of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// K-major, 128-byte-swizzled operand tile: rows of 128 bytes, 8-row swizzle atoms 1024 bytes apart.
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout [61,64))
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                 // leading byte offset (unused for swizzled K-major), canonical value
  d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}

// cute::UMMA::InstrDescriptor for kind::f16: D=f32, A=B=bf16, both K-major, M=128 (256 across a cta_group::2 pair), N=BN
template <int BN>
__device__ __forceinline__ uint32_t umma_idesc(int m = BM, int f16 = 0) {
  const uint32_t fmt = f16 ? 0u : 1u;        // A / B format field: 0 = F16, 1 = BF16
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// cta_group::2: issued by the leader CTA only; A rows 0-127 / 128-255 and the two halves of B come from the shared memory of the
// even / odd CTA at the SAME offsets, the accumulator rows go to the same TMEM address in each CTA
__device__ __forceinline__ void umma_bf16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// arrives on the barrier at this offset in every CTA of `mask` once the pair's MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_cg2(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
// arrives on the mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// 16-bit operand format of the launch: bf16, or fp16 with saturation (fp16 has 3 more mantissa bits - 8x smaller rounding error -
// but overflows at 65504: values beyond it are clamped to the largest finite number instead of becoming inf)
constexpr float F16_MAX = 65504.0f;
__device__ __forceinline__ uint32_t pack16(float lo, float hi, int f16) {
  if (f16) {
    __half2 v = __floats2half2_rn(fminf(fmaxf(lo, -F16_MAX), F16_MAX), fminf(fmaxf(hi, -F16_MAX), F16_MAX));
    return *reinterpret_cast<uint32_t*>(&v);
  }
  return pack_bf16(lo, hi);
}
__device__ __forceinline__ float2 unpack16(uint32_t w, int f16) {
  if (f16) return __half22float2(*reinterpret_cast<const __half2*>(&w));
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}
__device__ __forceinline__ unsigned short cvt16(float x, int f16) {
  if (f16) { __half h = __float2half_rn(fminf(fmaxf(x, -F16_MAX), F16_MAX)); return *reinterpret_cast<unsigned short*>(&h); }
  __nv_bfloat16 b = __float2bfloat16_rn(x);
  return *reinterpret_cast<unsigned short*>(&b);
}

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == HC_ACT_RELU) return fmaxf(x, 0.0f);
  if (act == HC_ACT_TANH) return tanhf(x);
  return x;
}

// bias + activation (+ elementwise multiplier) on one lane's 32 accumulator columns -> 64 bytes of bf16
template <int ACT>
__device__ __forceinline__ void store_bf16_row(const uint32_t (&r)[32], const float* __restrict__ bias, const float* __restrict__ mul,
                                               __nv_bfloat16* __restrict__ dst, int f16) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) v[t] = __uint_as_float(r[8 * i + t]);
    if (bias) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias) + 2 * i), b1 = __ldg(reinterpret_cast<const float4*>(bias) + 2 * i + 1);
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      if (ACT == HC_ACT_RELU) v[t] = fmaxf(v[t], 0.0f);
      if (ACT == HC_ACT_TANH) v[t] = tanhf(v[t]);
    }
    if (mul) {
      const float4 m0 = __ldg(reinterpret_cast<const float4*>(mul) + 2 * i), m1 = __ldg(reinterpret_cast<const float4*>(mul) + 2 * i + 1);
      v[0] *= m0.x; v[1] *= m0.y; v[2] *= m0.z; v[3] *= m0.w; v[4] *= m1.x; v[5] *= m1.y; v[6] *= m1.z; v[7] *= m1.w;
    }
    *reinterpret_cast<uint4*>(dst + 8 * i) = make_uint4(pack16(v[0], v[1], f16), pack16(v[2], v[3], f16), pack16(v[4], v[5], f16),
                                                        pack16(v[6], v[7], f16));
  }
}

// tile id -> (m block, n block): bands of `group_m` m-blocks; inside a band the n index is the slow one, so a
// wave of consecutive tile ids shares few B column-panels and a bounded set of A row-panels through L2.
__device__ __forceinline__ void tile_coords(const Params& p, int tiles_m, int tile, int& m_blk, int& n_blk, int = 0) {
  int per_band = p.group_m * p.tiles_n;
  int band = tile / per_band;
  int in = tile - band * per_band;
  int gm = min(p.group_m, tiles_m - band * p.group_m);
  m_blk = band * p.group_m + in % gm;
  n_blk = in / gm;
  // K-cell-sparse launches: tiles differ widely in length, and the static round-robin over them leaves the slowest CTA well above the
  // mean; the caller passes the tiles longest-first (tools/fc1_order_sim.py: max / mean 1.07 -> 1.03 for 74 CTA pairs)
  if (p.m_order) m_blk = __ldg(p.m_order + m_blk);
}

// =============================================================================================== kernel
// CG2 = tcgen05 cta_group::2 CTA pairs.  A separate instantiation: a kernel that contains 2-CTA instructions can only be launched
// with an even cluster dimension, so the single-CTA variants must not contain any.
template <int BN, int MS, bool CG2>
// (pair kernels: min-blocks 2 caps them at ~100 registers, so that the 320-thread CTA leaves half of the register file to the pooling kernel)
__global__ void __launch_bounds__(cta_threads<CG2>(), CG2 ? 2 : 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_bh, const Params p) {
  using C = Cfg<BN, MS>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B atoms need 1024-byte alignment
  const uint32_t bar_base = smem_base + C::DATA_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::MAX_STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * C::MAX_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * C::MAX_STAGES + C::ACC_STAGES + s); };
  constexpr int PBAR0 = 2 * C::MAX_STAGES + 2 * C::ACC_STAGES;
  auto pa_full = [&](int s) { return bar_base + 8u * (PBAR0 + s); };
  auto pa_empty = [&](int s) { return bar_base + 8u * (PBAR0 + C::PA_SLOTS + s); };
  auto pb_full = [&](int s) { return bar_base + 8u * (PBAR0 + 2 * C::PA_SLOTS + s); };
  auto pb_empty = [&](int s) { return bar_base + 8u * (PBAR0 + 2 * C::PA_SLOTS + C::PB_SLOTS + s); };
  const uint32_t pa_base = smem_base, pb_base = smem_base + C::PA_SLOTS * C::A_PATCH_BYTES;
  const uint32_t tmem_slot = bar_base + 8u * (PBAR0 + 2 * C::PA_SLOTS + 2 * C::PB_SLOTS);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // block-sparse conv: the number of M tiles comes from the device-side work-list length (no host round trip)
  const bool blk_mode = p.mode == HC_GEMM_CONV3_BLOCKS;
  const int blk_per_sub = blk_mode ? BM / (p.blk_w * p.blk_h) : 1;  // blocks per 128-row sub-tile (2, 4 or 8)
  const int n_blocks = blk_mode ? __ldg(p.n_blocks) : 0;
  const int tiles_m = blk_mode ? (n_blocks + MS * blk_per_sub - 1) / (MS * blk_per_sub) : p.tiles_m;
  // CTA pairs (CG2, tcgen05 cta_group::2): both CTAs walk the same sequence of pair tiles; the even CTA (leader) issues the MMAs
  const int rank = CG2 ? (int)cluster_ctarank() : 0;
  const int n_stages = CG2 ? C::CG2_STAGES : C::STAGES;
  const uint32_t stage_bytes = CG2 ? (uint32_t)C::CG2_STAGE_BYTES : (uint32_t)C::STAGE_BYTES;
  // pair tile (CG2): 16 listed blocks (256 pixels, the N side, half of them staged by each CTA) x 512 output channels (the M side:
  // 2 sub-tiles x 256 rows across the pair); a CTA keeps its 256 channels x 256 pixels in its 512 TMEM columns
  // (4x4-pixel blocks: 16 per tile, 8 staged by each CTA; 4x2-pixel blocks - one pooled cell tall: 32 per tile, 16 per CTA)
  constexpr int PAIR_COUT = 2 * 2 * BM;
  const int PAIR_BLOCKS = blk_mode ? 256 / (p.blk_w * p.blk_h) : 16;
  const int pair_ct = p.N / PAIR_COUT;                                 // output-channel tiles per pixel tile
  // (plain GEMM on pairs: a pair owns one 256-row M tile - 128 rows per CTA - of one N tile; p.tiles_m counts 256-row tiles)
  const int num_tiles = CG2 ? (blk_mode ? ((n_blocks + PAIR_BLOCKS - 1) / PAIR_BLOCKS) * pair_ct : tiles_m * p.tiles_n) : tiles_m * p.tiles_n;
  const int tile0 = CG2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = CG2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int num_kb = p.K / BK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C::MAX_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    // pair: the leader's accumulator-empty barrier collects the epilogue warps of BOTH CTAs
    for (int s = 0; s < C::ACC_STAGES; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), CG2 ? 16 : 4); }
    for (int s = 0; s < C::PA_SLOTS; ++s) { mbar_init(pa_full(s), 1); mbar_init(pa_empty(s), 1); }
    for (int s = 0; s < C::PB_SLOTS; ++s) { mbar_init(pb_full(s), 1); mbar_init(pb_empty(s), 1); }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CG2) {      // the same warp of both CTAs of the pair
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG2) cluster_sync_all();          // the peer's barriers and TMEM allocation exist before any remote arrive / pair MMA reaches them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (!CG2 && lane == 0 && p.patch) {      // (pair kernels are block-mode only: every tcgen05 op of a kernel uses one cta_group)
      int a_slot = 0, b_slot = 0;
      uint32_t a_phase = 0, b_phase = 0;
      const int cblks = p.c_in / BK;
      const int per_img = p.tiles_x * p.tiles_y;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        int m_blk, n_blk;
        tile_coords(p, tiles_m, tile, m_blk, n_blk, rank);
        const int img = m_blk / per_img;
        const int r = m_blk - img * per_img;
        const int y0 = (r / p.tiles_x) * (8 * MS), x0 = (r % p.tiles_x) * 16;
        for (int cb = 0; cb < cblks; ++cb) {
          for (int kx = 0; kx < 3; ++kx) {
            mbar_wait(pa_empty(a_slot), a_phase ^ 1u);
            mbar_expect_tx(pa_full(a_slot), C::A_PATCH_BYTES);
            tma_load_4d(pa_base + a_slot * C::A_PATCH_BYTES, &tmap_a, pa_full(a_slot), p.c_base + cb * BK, x0 + kx - 1, y0 - 1, img);
            if (++a_slot == C::PA_SLOTS) { a_slot = 0; a_phase ^= 1u; }
            for (int ky = 0; ky < 3; ++ky) {
              mbar_wait(pb_empty(b_slot), b_phase ^ 1u);
              mbar_expect_tx(pb_full(b_slot), C::B_BYTES);
              tma_load_2d(pb_base + b_slot * C::B_BYTES, &tmap_b, pb_full(b_slot), ((ky * 3 + kx) * cblks + cb) * BK, n_blk * BN);
              if (++b_slot == C::PB_SLOTS) { b_slot = 0; b_phase ^= 1u; }
            }
          }
        }
      }
    } else if (blk_mode) {
      // one stage per (channel block, kx, ky) - the K order of the patch pipeline, so results are bit-identical to the dense
      // kernel; every block of the tile contributes one {64 ch, blk_w x, blk_h y} box shifted by the tap (TMA zero-fills the halo).
      // The whole warp runs the loop: lane 0 waits for the slot and posts the byte count, then lane i issues the box of block i
      // and the last lane the weight tile, so the up to 17 TMA instructions of a stage issue side by side instead of in a chain.
      int stage = 0;
      uint32_t phase = 0;
      const int cblks = p.c_in / BK;
      const int nblk = MS * blk_per_sub;
      const uint32_t blk_bytes = (uint32_t)(p.blk_w * p.blk_h) * 128u;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        int m_blk, n_blk;
        if constexpr (CG2) { m_blk = tile / pair_ct; n_blk = tile - m_blk * pair_ct; }
        else tile_coords(p, tiles_m, tile, m_blk, n_blk, rank);
        const int e = CG2 ? (lane < PAIR_BLOCKS / 2 ? __ldg(p.blocks + min(m_blk * PAIR_BLOCKS + rank * (PAIR_BLOCKS / 2) + lane, n_blocks - 1)) : 0)
                          : (lane < nblk ? __ldg(p.blocks + min(m_blk * nblk + lane, n_blocks - 1)) : 0);
        const int ex = 2 * (e & 15) - 1, ey = 2 * ((e >> 4) & 15) - 1, eimg = e >> 8;
        for (int cb = 0; cb < cblks; ++cb) {
          for (int kx = 0; kx < 3; ++kx) {
            for (int ky = 0; ky < 3; ++ky) {
              const uint32_t a_dst = smem_base + stage * stage_bytes;
              if (lane == 0) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                // pair: the LEADER's full barrier counts the bytes of both CTAs (the peer only sends bytes, it never arrives)
                // bytes of the pixel boxes / of the weight tile(s) this CTA fetches per stage
                const uint32_t a_bytes = (p.dbg & 1) ? 0u : (uint32_t)(CG2 ? C::B_BYTES / 2 : MS * A_SUB_BYTES);
                const uint32_t b_bytes = (p.dbg & 2) ? 0u : (uint32_t)(CG2 ? MS * A_SUB_BYTES : C::B_BYTES);
                if (!CG2) mbar_expect_tx(full_bar(stage), a_bytes + b_bytes);
                else if (rank == 0) mbar_expect_tx(full_bar(stage), 2u * (a_bytes + b_bytes));
              }
              __syncwarp();
              if constexpr (CG2) {
                // pair: lanes 0-7 fetch this CTA's 8 of the tile's 16 pixel blocks (the N-side operand: the tensor cores of BOTH SMs
                // read all 16), lanes 30/31 this CTA's 128 output channels of weight sub-tile 0/1 (the M-side operand)
                const uint32_t lead_full = full_bar(stage) & PEER_BIT_MASK;
                if (lane < PAIR_BLOCKS / 2) {
                  if (!(p.dbg & 1))
                    tma_load_4d_cg2(a_dst + MS * A_SUB_BYTES + lane * blk_bytes, &tmap_a, lead_full, p.c_base + cb * BK, ex + kx, ey + ky, eimg);
                } else if (lane >= 30 && !(p.dbg & 2)) {
                  const int j = lane - 30;
                  tma_load_2d_cg2(a_dst + j * A_SUB_BYTES, &tmap_bh, lead_full, ((ky * 3 + kx) * cblks + cb) * BK,
                                  n_blk * PAIR_COUT + j * 2 * BM + rank * BM);
                }
              } else if (lane < nblk) {
                if (!(p.dbg & 1)) tma_load_4d(a_dst + lane * blk_bytes, &tmap_a, full_bar(stage), p.c_base + cb * BK, ex + kx, ey + ky, eimg);
              } else if (lane == 31 && !(p.dbg & 2)) {
                tma_load_2d(a_dst + MS * A_SUB_BYTES, &tmap_b, full_bar(stage), ((ky * 3 + kx) * cblks + cb) * BK, n_blk * BN);
              }
              if (++stage == n_stages) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    } else if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        int m_blk, n_blk;
        tile_coords(p, tiles_m, tile, m_blk, n_blk, rank);
        auto load_kb = [&](int kb, int jmask = (1 << MS) - 1) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t a_dst = smem_base + stage * stage_bytes;
          const uint32_t b_dst = a_dst + MS * A_SUB_BYTES;
          if constexpr (CG2) {
            // pair: this CTA stages ITS 128 rows of every 256-row unit of the tile (sub-tile j = unit m_blk*MS + j; `jmask`: the units
            // that visit this K block) and ITS half of the weight tile's columns; the leader's full barrier counts the bytes of both
            // CTAs (the peer only sends bytes, it never arrives)
            const uint32_t lead_full = full_bar(stage) & PEER_BIT_MASK;
            if (rank == 0) mbar_expect_tx(full_bar(stage), 2u * (uint32_t)(__popc(jmask) * A_SUB_BYTES + C::B_BYTES / 2));
#pragma unroll
            for (int j = 0; j < MS; ++j)
              if ((jmask >> j) & 1)
                tma_load_2d_cg2(a_dst + j * A_SUB_BYTES, &tmap_a, lead_full, kb * BK, ((m_blk * MS + j) * 2 + rank) * BM);
            tma_load_2d_cg2(b_dst, &tmap_bh, lead_full, kb * BK, n_blk * BN + rank * (BN / 2));
          } else {
            mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
#pragma unroll
            for (int j = 0; j < MS; ++j)
              tma_load_2d(a_dst + j * A_SUB_BYTES, &tmap_a, full_bar(stage), kb * BK, (m_blk * MS + j) * BM);
            tma_load_2d(b_dst, &tmap_b, full_bar(stage), kb * BK, n_blk * BN);
          }
          if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        };
        if (p.k_masks) {
          // K-cell-sparse: only the K cells some row of this M tile is non-zero in (ascending cell order)
          if constexpr (CG2) {
            // pair tile of MS 256-row units, each with its OWN mask: the pair walks the UNION of the cells, a weight K block is staged
            // once for all units, a unit's rows only for the cells in its own mask (no extra MMAs, the weight slab is fetched once)
            unsigned long long kj[MS], ku = 0ull;
#pragma unroll
            for (int j = 0; j < MS; ++j) { kj[j] = (m_blk * MS + j) * 2 * BM < p.M ? __ldg(p.k_masks + m_blk * MS + j) : 0ull; ku |= kj[j]; }
            for (; ku; ku &= ku - 1) {
              const int c = __ffsll((long long)ku) - 1;
              int jm = 0;
#pragma unroll
              for (int j = 0; j < MS; ++j) jm |= (int)((kj[j] >> c) & 1ull) << j;
              for (int i = 0; i < p.k_cell_kb; ++i) load_kb(c * p.k_cell_kb + i, jm);
            }
          } else {
            for (unsigned long long km = __ldg(p.k_masks + m_blk); km; km &= km - 1) {
              const int kb0 = (__ffsll((long long)km) - 1) * p.k_cell_kb;
              for (int i = 0; i < p.k_cell_kb; ++i) load_kb(kb0 + i);
            }
          }
        } else {
          for (int kb = 0; kb < num_kb; ++kb) load_kb(kb);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (!CG2 && lane == 0 && p.patch) {      // (pair kernels are block-mode only: every tcgen05 op of a kernel uses one cta_group)
      const uint32_t idesc = umma_idesc<BN>(BM, p.f16);
      int a_slot = 0, b_slot = 0, acc = 0;
      uint32_t a_phase = 0, b_phase = 0, acc_phase = 0;
      const int n_ax = (p.c_in / BK) * 3;                 // (channel block, kx) steps per tile
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        for (int ax = 0; ax < n_ax; ++ax) {
          mbar_wait(pa_full(a_slot), a_phase);            // the A patch of this (channel block, kx) has landed
          tc_fence_after();
          const uint32_t a_src = pa_base + a_slot * C::A_PATCH_BYTES;
          for (int ky = 0; ky < 3; ++ky) {
            mbar_wait(pb_full(b_slot), b_phase);
            tc_fence_after();
            const uint64_t bdesc = umma_desc_sw128(pb_base + b_slot * C::B_BYTES);
#pragma unroll
            for (int j = 0; j < MS; ++j) {
              // tap ky of sub-tile j starts (8j + ky) pixel rows into the patch: 16 pixels x 128 B = 2048 B per row,
              // a whole number of 1024-byte swizzle atoms, so only the descriptor start address moves
              const uint64_t adesc = umma_desc_sw128(a_src + (uint32_t)(8 * j + ky) * 2048u);
              const uint32_t d = tmem_base + (uint32_t)(acc * C::ACC_COLS + j * BN);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)
                umma_bf16(d, adesc + 2u * k, bdesc + 2u * k, idesc, (ax | ky | k) ? 1u : 0u);
            }
            umma_commit(pb_empty(b_slot));
            if (++b_slot == C::PB_SLOTS) { b_slot = 0; b_phase ^= 1u; }
          }
          umma_commit(pa_empty(a_slot));                  // patch free once the MMAs of its three taps retire
          if (ax == n_ax - 1) umma_commit(tfull_bar(acc));
          if (++a_slot == C::PA_SLOTS) { a_slot = 0; a_phase ^= 1u; }
        }
        if (++acc == C::ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
      }
    } else if (lane == 0 && rank == 0) {               // pair: only the leader CTA issues (its MMAs run on both SMs)
      const uint32_t idesc = umma_idesc<BN>(CG2 ? 2 * BM : BM, p.f16);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);      // epilogue (of both CTAs of a pair) has drained this accumulator stage
        tc_fence_after();
        uint32_t started = 0;                             // bit j: sub-tile j has issued its first MMA (which overwrites the accumulator)
        auto mma_kb = [&](int jmask = (1 << MS) - 1) {
          mbar_wait(full_bar(stage), phase);              // TMA bytes of this stage have landed
          tc_fence_after();
          const uint32_t a_src = smem_base + stage * stage_bytes;
          const uint64_t bdesc = umma_desc_sw128(a_src + MS * A_SUB_BYTES);
#pragma unroll
          for (int j = 0; j < MS; ++j) {
            if (p.dbg & 4) break;
            if (!((jmask >> j) & 1)) continue;            // pair tile, K-cell-sparse: this unit does not visit the cell
            const uint64_t adesc = umma_desc_sw128(a_src + j * A_SUB_BYTES);
            const uint32_t d = tmem_base + (uint32_t)(acc * C::ACC_COLS + j * BN);
            const uint32_t acc_on = (started >> j) & 1u;
            if constexpr (CG2) {
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)
                umma_bf16_cg2(d, adesc + 2u * k, bdesc + 2u * k, idesc, (acc_on | k) ? 1u : 0u);
            } else {
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)       // +32 bytes (>>4 = 2) per 16-element K step inside the swizzle row
                umma_bf16(d, adesc + 2u * k, bdesc + 2u * k, idesc, (acc_on | k) ? 1u : 0u);
            }
          }
          started |= (uint32_t)jmask;
          if constexpr (CG2) umma_commit_cg2(empty_bar(stage), (uint16_t)3);  // the slot of BOTH CTAs is free once the pair's MMAs retire
          else umma_commit(empty_bar(stage));             // smem slot free once these MMAs retire
          if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        };
        if (p.k_masks) {
          int m_blk, n_blk;
          tile_coords(p, tiles_m, tile, m_blk, n_blk, rank);
          if constexpr (CG2) {                            // the producer's walk: union of the units' masks, per-unit participation
            unsigned long long kj[MS], ku = 0ull;
#pragma unroll
            for (int j = 0; j < MS; ++j) { kj[j] = (m_blk * MS + j) * 2 * BM < p.M ? __ldg(p.k_masks + m_blk * MS + j) : 0ull; ku |= kj[j]; }
            for (; ku; ku &= ku - 1) {
              const int c = __ffsll((long long)ku) - 1;
              int jm = 0;
#pragma unroll
              for (int j = 0; j < MS; ++j) jm |= (int)((kj[j] >> c) & 1ull) << j;
              for (int i = 0; i < p.k_cell_kb; ++i) mma_kb(jm);
            }
          } else {
            const int n_cells = __popcll(__ldg(p.k_masks + m_blk));
            for (int i = 0; i < n_cells * p.k_cell_kb; ++i) mma_kb();
          }
        } else {
          for (int kb = 0; kb < num_kb; ++kb) mma_kb();
        }
        // accumulator complete once the MMAs above retire; an empty cell mask issued none: plain arrive, the epilogue substitutes zeros
        if constexpr (CG2) {
          if (started) umma_commit_cg2(tfull_bar(acc), (uint16_t)3);         // both CTAs' epilogues
          else { mbar_arrive(tfull_bar(acc)); mbar_arrive_cluster((tfull_bar(acc) & PEER_BIT_MASK) | ~PEER_BIT_MASK); }   // empty cell mask: no MMA to wait for
        } else if (started) umma_commit(tfull_bar(acc));
        else mbar_arrive(tfull_bar(acc));
        if (++acc == C::ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps 2..5
    const int q = warp & 3;                               // TMEM lane quarter this warp may access
    const int egrp = CG2 ? (warp - 2) >> 2 : 0;           // pair kernels: which of the two warps of the quarter (alternate chunks)
    constexpr int ESTEP = CG2 ? 2 : 1;
    const int row_in_tile = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      if (CG2 && blk_mode) {
        // Pair tile, transposed roles: TMEM lanes are OUTPUT CHANNELS (this thread owns one channel of sub-tile j), columns are the
        // 256 pixels of the tile's 16 blocks - a 4x4-pixel block is 16 consecutive columns, so the 2x2 max-pool is register-local.
        const int m_blk = tile / pair_ct, n_blk = tile - m_blk * pair_ct;
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const long long map_elems = (long long)(p.H / 2) * (p.W / 2) * p.ldc;
#pragma unroll 1
        for (int j = 0; j < MS; ++j) {
          const int cout = n_blk * PAIR_COUT + j * 2 * BM + rank * BM + q * 32 + lane;
          const float bias = __ldg(p.bias + cout);
#pragma unroll 1
          for (int ch = egrp; ch < BN / 32; ch += ESTEP) {
            uint32_t r[32];
            __syncwarp();
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C::ACC_COLS + j * BN + ch * 32), r);
#pragma unroll
            for (int hb = 0; hb < 4; ++hb) {                 // 4x2-pixel blocks: four per 32-column chunk, pixel (y, x) = column 4y + x,
              if (p.blk_h != 2) break;                       // i.e. two pooled cells side by side
              const int e = __ldg(p.blocks + min(m_blk * PAIR_BLOCKS + 4 * ch + hb, n_blocks - 1));
              const int o_img = e >> 8, cy0 = (e >> 4) & 15, cx0 = e & 15;
              const long long out_base = (long long)o_img * map_elems;
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                const int i0 = 8 * hb + 2 * c;
                const float m = fmaxf(fmaxf(__uint_as_float(r[i0]), __uint_as_float(r[i0 + 1])),
                                      fmaxf(__uint_as_float(r[i0 + 4]), __uint_as_float(r[i0 + 5])));
                const unsigned short x = cvt16(fmaxf(m + bias, 0.0f), p.f16);
                const long long off = ((long long)cy0 * (p.W / 2) + (cx0 + c)) * p.ldc + p.c_off + cout;
                reinterpret_cast<unsigned short*>(p.out)[out_base + off] = x;
              }
            }
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {                 // 4x4-pixel blocks: the two blocks of this 32-column chunk (warp-uniform entries)
              if (p.blk_h == 2) break;
              const int e = __ldg(p.blocks + min(m_blk * PAIR_BLOCKS + 2 * ch + hb, n_blocks - 1));
              const int o_img = e >> 8, cy0 = (e >> 4) & 15, cx0 = e & 15;
              // (the pooled-difference epilogue is split for pairs: this kernel writes the pooled value x by LOCAL pair into the
              // caller's scratch map, pair_diff_kernel then forms d from coalesced 16-byte vectors - per-channel 2-byte gathers of
              // the three maps from here cost more than the whole main loop saved)
              const long long out_base = (long long)o_img * map_elems;
#pragma unroll
              for (int c = 0; c < 4; ++c) {                  // pooled cell (c >> 1, c & 1) of the block: pixels (2cy + dy, 2cx + dx)
                const int i0 = 16 * hb + 8 * (c >> 1) + 2 * (c & 1);
                const float m = fmaxf(fmaxf(__uint_as_float(r[i0]), __uint_as_float(r[i0 + 1])),
                                      fmaxf(__uint_as_float(r[i0 + 4]), __uint_as_float(r[i0 + 5])));
                const unsigned short x = cvt16(fmaxf(m + bias, 0.0f), p.f16);
                const long long off = ((long long)(cy0 + (c >> 1)) * (p.W / 2) + (cx0 + (c & 1))) * p.ldc + p.c_off + cout;
                reinterpret_cast<unsigned short*>(p.out)[out_base + off] = x;
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_bar(acc) & PEER_BIT_MASK);    // the leader's barrier (a remote arrive for the odd CTA)
        if (++acc == C::ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
        continue;
      }
      int m_blk, n_blk;
      tile_coords(p, tiles_m, tile, m_blk, n_blk, rank);
      int t_img = 0, t_y0 = 0, t_x0 = 0;                  // conv: tile origin (image, first pixel row / column)
      if (p.mode == HC_GEMM_CONV3) {   // dense conv only; block mode decodes its origin per warp below
        const int per_img = p.tiles_x * p.tiles_y;
        t_img = m_blk / per_img;
        const int rr = m_blk - t_img * per_img;
        t_y0 = (rr / p.tiles_x) * (8 * MS);
        t_x0 = (rr % p.tiles_x) * 16;
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int n0 = n_blk * BN;
      // K-cell-sparse tile with an empty cell mask: no MMA ran, the accumulator is all zeros by definition
      // (pair tiles carry one mask per 256-row unit = sub-tile)
      const bool no_acc_tile = !CG2 && p.k_masks != nullptr && __ldg(p.k_masks + m_blk) == 0ull;
#pragma unroll 1
      for (int j = 0; j < MS; ++j) {
        const bool no_acc = no_acc_tile || (CG2 && p.k_masks != nullptr &&
                                            ((m_blk * MS + j) * 2 * BM >= p.M || __ldg(p.k_masks + m_blk * MS + j) == 0ull));
#pragma unroll 1
        for (int ch = egrp; ch < BN / 32; ch += ESTEP) {
          uint32_t r[32];
          __syncwarp();                                    // tcgen05.ld is .sync.aligned: reconverge after the predicated stores
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C::ACC_COLS + j * BN + ch * 32), r);
          if (no_acc) {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = 0u;
          }
          const int col0 = n0 + ch * 32;
          if (!CG2 && (p.epi == HC_EPI_POOL_BF16 || p.epi == HC_EPI_POOL_DIFF_BF16)) {      // (pair kernels pool in their own epilogue above)
            // rows of a sub-tile are pixels (yl, xl) = (row/16, row%16); this warp holds yl in {2q, 2q+1}.
            // 2x2 max-pool partners are lane^1 (x) and lane^16 (y): butterfly reduce-scatter, after which the
            // lane with bits (ybit, xbit) owns the pooled maximum of columns [ybit*16 + xbit*8, +8).
            // (block mode: rows of a block are pixels (yl, xl) = (row / blk_w, row % blk_w), so the y partner is lane ^ blk_w)
            const int ysh = blk_mode ? (p.blk_w == 8 ? 3 : 2) : 4;
            const int ybit = (lane >> ysh) & 1, xbit = lane & 1;
            float h[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float mine = __uint_as_float(ybit ? r[16 + i] : r[i]);
              float send = __uint_as_float(ybit ? r[i] : r[16 + i]);
              h[i] = fmaxf(mine, __shfl_xor_sync(0xffffffffu, send, 1 << ysh));
            }
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float mine = xbit ? h[8 + i] : h[i];
              float send = xbit ? h[i] : h[8 + i];
              o[i] = fmaxf(mine, __shfl_xor_sync(0xffffffffu, send, 1));
            }
            const int cbase = col0 + ybit * 16 + xbit * 8;
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float a = fmaxf(o[2 * i] + __ldg(p.bias + cbase + 2 * i), 0.0f);
              float b = fmaxf(o[2 * i + 1] + __ldg(p.bias + cbase + 2 * i + 1), 0.0f);
              w[i] = pack16(a, b, p.f16);
            }
            int py = (t_y0 + 8 * j) / 2 + q;                            // pooled row
            int px = t_x0 / 2 + ((lane & 15) >> 1);                     // pooled col
            int o_img = t_img;
            if (blk_mode) {
              // this lane's row sits in block (row / rows_per_block) of sub-tile j (a warp spans 1 or 2 blocks), at pixel
              // (yl, xl) = (r / blk_w, r % blk_w) of it, r = row % rows_per_block
              const int rows_pb = p.blk_w * p.blk_h;
              const int row_s = q * 32 + lane;
              const int bi = j * blk_per_sub + row_s / rows_pb;
              const int rb = row_s % rows_pb;
              const int e = __ldg(p.blocks + min((m_blk * MS * blk_per_sub) + bi, n_blocks - 1));
              o_img = e >> 8;
              py = ((e >> 4) & 15) + ((rb / p.blk_w) >> 1);
              px = (e & 15) + ((rb % p.blk_w) >> 1);
            }
            if (p.epi == HC_EPI_POOL_DIFF_BF16) {
              // shared-footprint fc1 operand: what this pair's cell adds to the sum of its two per-box maps,
              //   d = (x - sub_map[s]) - (obj_map[o] - background), x already rounded to bf16 like the maps;
              // exactly 0 wherever only one box (or none) reaches the cell, because x then equals that map bit for bit
              const long long map_elems = (long long)(p.H / 2) * (p.W / 2) * p.ldc;
              const long long cell_off = ((long long)py * (p.W / 2) + px) * p.ldc + p.c_off + cbase;
              const uint4 S = __ldg(reinterpret_cast<const uint4*>(p.diff_sub + (long long)__ldg(p.pair_sub + o_img) * map_elems + cell_off));
              const uint4 O = __ldg(reinterpret_cast<const uint4*>(p.diff_obj + (long long)__ldg(p.pair_obj + o_img) * map_elems + cell_off));
              const uint4 G = __ldg(reinterpret_cast<const uint4*>(p.diff_bg + cell_off));
              const uint32_t sv[4] = {S.x, S.y, S.z, S.w}, ov[4] = {O.x, O.y, O.z, O.w}, gv[4] = {G.x, G.y, G.z, G.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 xf = unpack16(w[i], p.f16), sf = unpack16(sv[i], p.f16), of = unpack16(ov[i], p.f16), gf = unpack16(gv[i], p.f16);
                w[i] = pack16(__fsub_rn(__fsub_rn(xf.x, sf.x), __fsub_rn(of.x, gf.x)), __fsub_rn(__fsub_rn(xf.y, sf.y), __fsub_rn(of.y, gf.y)), p.f16);
              }
              __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)__ldg(p.pair_row + o_img) * map_elems + cell_off;
              *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
            } else {
              __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) +
                                   (((long long)o_img * (p.H / 2) + py) * (p.W / 2) + px) * p.ldc + p.c_off + cbase;
              *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          } else {
            // output row of tile row `tr` (plain: GEMM row; conv: NHWC pixel index), -1 when outside M
            auto out_row = [&](int tr) -> long long {
              if constexpr (CG2) {   // plain GEMM on a pair: this CTA holds rows [rank*128, +128) of sub-tile j of the 256-row tile
                long long row = ((long long)(m_blk * MS + j) * 2 + rank) * BM + tr;
                return row < p.M ? row : -1;
              }
              if (p.mode == HC_GEMM_CONV3)
                return ((long long)t_img * p.H + (t_y0 + 8 * j + (tr >> 4))) * p.W + (t_x0 + (tr & 15));
              if (blk_mode) {   // un-pooled block mode: tile row -> pixel (y0 + r / blk_w, x0 + r % blk_w) of its block's image
                const int rows_pb = p.blk_w * p.blk_h;
                const int e = __ldg(p.blocks + min((m_blk * MS + j) * blk_per_sub + tr / rows_pb, n_blocks - 1));
                const int rb = tr % rows_pb;
                return ((long long)(e >> 8) * p.H + (2 * ((e >> 4) & 15) + rb / p.blk_w)) * p.W + (2 * (e & 15) + rb % p.blk_w);
              }
              long long row = (long long)(m_blk * MS + j) * BM + tr;
              return row < p.M ? row : -1;
            };
            const long long grow = out_row(row_in_tile);          // GEMM row (indexes mul / the row gathers)
            const long long row = (grow >= 0 && p.out_rows) ? (long long)__ldg(p.out_rows + grow) : grow;   // output row
            if (p.epi == HC_EPI_F32) {
              if (row >= 0) {
                float* dst = reinterpret_cast<float*>(p.out) + row * p.ldc + p.c_off + col0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  float4 v;
                  v.x = __uint_as_float(r[4 * i]); v.y = __uint_as_float(r[4 * i + 1]);
                  v.z = __uint_as_float(r[4 * i + 2]); v.w = __uint_as_float(r[4 * i + 3]);
                  if (p.bias) {
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + i);
                    v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
                  }
                  if (p.mul) {
                    float4 mm = __ldg(reinterpret_cast<const float4*>(p.mul + grow * p.ld_mul + col0 + 4 * i));
                    v.x *= mm.x; v.y *= mm.y; v.z *= mm.z; v.w *= mm.w;
                  }
                  *reinterpret_cast<float4*>(dst + 4 * i) = v;
                }
              }
            } else if (!CG2 && p.epi == HC_EPI_SPLIT3_BF16) {
              // bf16x3 A-operand layout for a following GEMM: out[row] = [hi | lo | hi] over 3*N columns, hi = bf16(x),
              // lo = bf16(x - hi), x = act(acc + bias) * mul - the f32 intermediate never goes to HBM
              if (row >= 0) {
                __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldc + p.c_off + col0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  float v[8];
#pragma unroll
                  for (int t = 0; t < 8; ++t) v[t] = __uint_as_float(r[8 * i + t]);
                  if (p.bias) {
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + 2 * i);
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + 2 * i + 1);
                    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
                  }
                  if (p.act == HC_ACT_RELU) {
#pragma unroll
                    for (int t = 0; t < 8; ++t) v[t] = fmaxf(v[t], 0.0f);
                  }
                  if (p.mul) {
                    const float4 m0 = __ldg(reinterpret_cast<const float4*>(p.mul + grow * p.ld_mul + col0) + 2 * i);
                    const float4 m1 = __ldg(reinterpret_cast<const float4*>(p.mul + grow * p.ld_mul + col0) + 2 * i + 1);
                    v[0] *= m0.x; v[1] *= m0.y; v[2] *= m0.z; v[3] *= m0.w; v[4] *= m1.x; v[5] *= m1.y; v[6] *= m1.z; v[7] *= m1.w;
                  }
                  uint32_t hi[4], lo[4];
#pragma unroll
                  for (int t = 0; t < 4; ++t) {
                    const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * t]), h1 = __float2bfloat16_rn(v[2 * t + 1]);
                    hi[t] = pack_bf16(__bfloat162float(h0), __bfloat162float(h1));
                    lo[t] = pack_bf16(v[2 * t] - __bfloat162float(h0), v[2 * t + 1] - __bfloat162float(h1));
                  }
                  const uint4 H = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                  *reinterpret_cast<uint4*>(dst + 8 * i) = H;
                  *reinterpret_cast<uint4*>(dst + p.N + 8 * i) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                  *reinterpret_cast<uint4*>(dst + 2 * p.N + 8 * i) = H;
                }
              }
            } else {
              if (row >= 0) {
                __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldc + p.c_off + col0;
                const float* mul_row = p.mul ? p.mul + grow * p.ld_mul + col0 : nullptr;
                if (p.add_a) {
                  // shared-footprint fc1: the per-box fc1 rows of this pair's subject and object join the accumulator in fp32
                  const float* ra = p.add_a + (long long)__ldg(p.add_a_rows + grow) * p.ld_add + col0;
                  const float* rb = p.add_b + (long long)__ldg(p.add_b_rows + grow) * p.ld_add + col0;
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const float4 a4 = __ldg(reinterpret_cast<const float4*>(ra) + i), b4 = __ldg(reinterpret_cast<const float4*>(rb) + i);
                    r[4 * i] = __float_as_uint(__uint_as_float(r[4 * i]) + (a4.x + b4.x));
                    r[4 * i + 1] = __float_as_uint(__uint_as_float(r[4 * i + 1]) + (a4.y + b4.y));
                    r[4 * i + 2] = __float_as_uint(__uint_as_float(r[4 * i + 2]) + (a4.z + b4.z));
                    r[4 * i + 3] = __float_as_uint(__uint_as_float(r[4 * i + 3]) + (a4.w + b4.w));
                  }
                }
                // the activation switch is hoisted out of the element loops (a branch per element serialises the
                // single epilogue warp of each scheduler)
                if (p.act == HC_ACT_RELU) store_bf16_row<HC_ACT_RELU>(r, p.bias ? p.bias + col0 : nullptr, mul_row, dst, p.f16);
                else if (p.act == HC_ACT_TANH) store_bf16_row<HC_ACT_TANH>(r, p.bias ? p.bias + col0 : nullptr, mul_row, dst, p.f16);
                else store_bf16_row<HC_ACT_NONE>(r, p.bias ? p.bias + col0 : nullptr, mul_row, dst, p.f16);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG2) mbar_arrive_cluster(tempty_bar(acc) & PEER_BIT_MASK);   // the leader's barrier collects both CTAs' epilogue warps
        else mbar_arrive(tempty_bar(acc));
      }
      if (++acc == C::ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CG2) cluster_sync_all();          // no CTA leaves (or frees TMEM) while the pair's MMAs / remote arrives can still reach it
  if (warp == 1) {
    if constexpr (CG2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// Second half of HC_EPI_POOL_DIFF_BF16 for the pair kernel: for every listed block, d = (x - sub_map) - (obj_map - background) on the
// cells of the block, x = the pooled bf16 value the pair kernel left in `scratch` (by local pair), d -> out[pair_row[pair]].  Same
// operations in the same order as the fused single-CTA epilogue, so d is bit-identical; a cell listed twice (clamped blocks may
// overlap) is simply written twice with the same value because x is read from the scratch map, never from `out`.
__global__ void __launch_bounds__(256)
pair_diff_kernel(const int* __restrict__ blocks, const int* __restrict__ n_blocks_p, const uint4* __restrict__ scratch,
                 const uint4* __restrict__ diff_sub, const uint4* __restrict__ diff_obj, const uint4* __restrict__ diff_bg,
                 const int* __restrict__ pair_sub, const int* __restrict__ pair_obj, const int* __restrict__ pair_row, int cells_w,
                 int cells_h, int map_w, long long cell_vec, long long map_vec, uint4* __restrict__ out, int f16) {
  const int n_blocks = __ldg(n_blocks_p);
  const int per_block = cells_w * cells_h * (int)cell_vec;          // 16-byte vectors per block
  const long long total = (long long)n_blocks * per_block;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_block), r = (int)(i - (long long)b * per_block);
    const int cell = r / (int)cell_vec, v = r - cell * (int)cell_vec;
    const int e = __ldg(blocks + b);
    const int pair = e >> 8, cy = ((e >> 4) & 15) + cell / cells_w, cx = (e & 15) + cell % cells_w;
    const long long off = ((long long)cy * map_w + cx) * cell_vec + v;
    const uint4 X = __ldg(scratch + (long long)pair * map_vec + off);
    const uint4 S = __ldg(diff_sub + (long long)__ldg(pair_sub + pair) * map_vec + off);
    const uint4 O = __ldg(diff_obj + (long long)__ldg(pair_obj + pair) * map_vec + off);
    const uint4 G = __ldg(diff_bg + off);
    const uint32_t xv[4] = {X.x, X.y, X.z, X.w}, sv[4] = {S.x, S.y, S.z, S.w}, ov[4] = {O.x, O.y, O.z, O.w}, gv[4] = {G.x, G.y, G.z, G.w};
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 xf = unpack16(xv[k], f16), sf = unpack16(sv[k], f16), of = unpack16(ov[k], f16), gf = unpack16(gv[k], f16);
      w[k] = pack16(__fsub_rn(__fsub_rn(xf.x, sf.x), __fsub_rn(of.x, gf.x)), __fsub_rn(__fsub_rn(xf.y, sf.y), __fsub_rn(of.y, gf.y)), f16);
    }
    out[(long long)__ldg(pair_row + pair) * map_vec + off] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// =============================================================================================== host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                    const cuuint32_t* box, int f16 = 0) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(HC_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_last_error, sizeof(g_last_error), "cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
    return HC_E_CUDA;
  }
  return HC_OK;
}

template <int BN, int MS, bool CG2>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tbh, const Params& p, cudaStream_t stream) {
  using C = Cfg<BN, MS>;
  // the opt-in is per device (context): keyed on the current device, so a process that touches a second GPU configures it too
  static bool configured[HC_MAX_DEVICES] = {};
  const int dev = current_device();
  if (dev < 0 || dev >= HC_MAX_DEVICES) return fail(HC_E_CUDA, "tc_gemm: device index out of range");
  if (!configured[dev]) {
    if (cudaFuncSetAttribute(tc_gemm_kernel<BN, MS, CG2>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) != cudaSuccess)
      return cuda_status("cudaFuncSetAttribute(tc_gemm_kernel)");
    configured[dev] = true;
  }
  int tiles = p.tiles_m * p.tiles_n;
  if (CG2 && p.mode == HC_GEMM_PLAIN) tiles *= 2;                                          // a pair of CTAs per tile
  int grid = (p.mode == HC_GEMM_CONV3_BLOCKS || tiles >= num_sms()) ? num_sms() : tiles;   // block mode: tile count lives on the device
  if (CG2) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(grid & ~1));
    cfg.blockDim = dim3(cta_threads<CG2>());
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, tc_gemm_kernel<BN, MS, CG2>, ta, tb, tbh, p) != cudaSuccess) return cuda_status("tc_gemm_kernel cluster launch");
    return cuda_status("tc_gemm_kernel cluster launch");
  }
  tc_gemm_kernel<BN, MS, CG2><<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, tb, p);
  return cuda_status("tc_gemm_kernel launch");
}

}  // namespace tc
}  // namespace hc

using namespace hc;

extern "C" int hc_tc_gemm(const hc_gemm_desc* d, hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(d != nullptr, HC_E_NULL, "hc_tc_gemm: desc is NULL");
  HC_REQUIRE(d->a && d->b && d->out, HC_E_NULL, "hc_tc_gemm: a/b/out must be non-NULL");
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  HC_REQUIRE(d->m > 0 && d->n > 0 && d->k > 0, HC_E_SHAPE, "hc_tc_gemm: m,n,k must be positive");
  HC_REQUIRE(d->k % tc::BK == 0, HC_E_SHAPE, "hc_tc_gemm: K must be a multiple of 64");
  HC_REQUIRE(d->n % 128 == 0, HC_E_SHAPE, "hc_tc_gemm: N must be a multiple of 128");
  HC_REQUIRE(aligned16(d->a) && aligned16(d->b) && aligned16(d->out), HC_E_ALIGN, "hc_tc_gemm: a/b/out must be 16-byte aligned");
  HC_REQUIRE(d->ldc % 8 == 0 && d->c_off % 8 == 0, HC_E_ALIGN, "hc_tc_gemm: ldc and c_off must be multiples of 8");
  HC_REQUIRE(!d->bias || aligned16(d->bias), HC_E_ALIGN, "hc_tc_gemm: bias must be 16-byte aligned");
  HC_REQUIRE(d->epilogue >= 0 && d->epilogue <= 4, HC_E_SHAPE, "hc_tc_gemm: unknown epilogue");
  HC_REQUIRE(d->epilogue != HC_EPI_SPLIT3_BF16 || (d->mode == HC_GEMM_PLAIN && d->act != HC_ACT_TANH && d->ldc >= 3 * d->n), HC_E_SHAPE,
             "hc_tc_gemm: the bf16x3 split epilogue needs a plain GEMM, no tanh and ldc >= 3*N");
  const bool pooled = d->epilogue == HC_EPI_POOL_BF16 || d->epilogue == HC_EPI_POOL_DIFF_BF16;
  HC_REQUIRE(!pooled || ((d->mode == HC_GEMM_CONV3 || d->mode == HC_GEMM_CONV3_BLOCKS) && d->bias), HC_E_SHAPE,
             "hc_tc_gemm: pooled epilogue needs conv mode and a bias");
  const int blk_w = d->block_cols ? d->block_cols : 8;
  HC_REQUIRE(d->mode != HC_GEMM_CONV3_BLOCKS || ((pooled || d->epilogue == HC_EPI_BF16) && d->blocks && d->n_blocks &&
                                                 (((d->block_rows == 8 || d->block_rows == 4) && (blk_w == 8 || (blk_w == 4 && d->block_rows == 4))) ||
                                                  (blk_w == 4 && d->block_rows == 2 && d->cta_pairs))),
             HC_E_SHAPE, "hc_tc_gemm: block-sparse conv needs a pooled or the bf16 epilogue, a work list and 8x8, 8x4, 4x4 or (CTA pairs only) 4x2-pixel blocks");
  HC_REQUIRE(d->epilogue != HC_EPI_POOL_DIFF_BF16 ||
                 (d->mode == HC_GEMM_CONV3_BLOCKS && d->diff_sub && d->diff_obj && d->diff_bg && d->pair_sub && d->pair_obj && d->pair_row &&
                  aligned16(d->diff_sub) && aligned16(d->diff_obj) && aligned16(d->diff_bg)),
             HC_E_SHAPE, "hc_tc_gemm: the pooled-difference epilogue needs block mode, three 16-byte aligned maps and the three pair index arrays");
  HC_REQUIRE(!d->k_masks || (d->mode == HC_GEMM_PLAIN && d->k_cell > 0 && d->k_cell % tc::BK == 0 && d->k % d->k_cell == 0 && d->k / d->k_cell <= 64),
             HC_E_SHAPE, "hc_tc_gemm: k_masks needs a plain GEMM and k_cell a multiple of 64 with K / k_cell <= 64");
  HC_REQUIRE((!d->add_a && !d->add_b) || (d->add_a && d->add_b && d->add_a_rows && d->add_b_rows && d->mode == HC_GEMM_PLAIN &&
                                          d->epilogue == HC_EPI_BF16 && d->ld_add % 4 == 0 && d->ld_add >= d->n && aligned16(d->add_a) &&
                                          aligned16(d->add_b)),
             HC_E_SHAPE, "hc_tc_gemm: row gathers need a plain bf16-epilogue GEMM, both f32 tables (16-byte aligned, ld_add % 4 == 0) and both index arrays");
  HC_REQUIRE(!d->out_rows || (d->mode == HC_GEMM_PLAIN && !d->mul && d->epilogue != HC_EPI_SPLIT3_BF16), HC_E_SHAPE,
             "hc_tc_gemm: out_rows needs a plain GEMM without mul and a bf16 / f32 epilogue");
  HC_REQUIRE(d->m < (1ll << 31) && d->n < (1ll << 31) && d->k < (1ll << 31), HC_E_SHAPE, "hc_tc_gemm: dims exceed int32");

  const int BN = (d->n % 256 == 0) ? 256 : 128;
  int MS = d->m_sub ? d->m_sub : 1;
  HC_REQUIRE(MS == 1 || MS == 2, HC_E_SHAPE, "hc_tc_gemm: m_sub must be 1 or 2");

  tc::Params p;
  memset(&p, 0, sizeof(p));
  p.M = (int)d->m; p.N = (int)d->n; p.K = (int)d->k;
  p.mode = d->mode; p.epi = d->epilogue; p.act = d->act;
  p.ldc = d->ldc; p.c_off = d->c_off; p.bias = d->bias; p.out = d->out;
  p.mul = d->mul; p.ld_mul = d->ld_mul;
  p.f16 = d->operand_f16 ? 1 : 0;
  HC_REQUIRE(!p.f16 || d->epilogue != HC_EPI_SPLIT3_BF16, HC_E_SHAPE, "hc_tc_gemm: the bf16x3 split epilogue is bf16-only");
  p.patch = d->mode == HC_GEMM_CONV3 ? 1 : 0;
  p.blocks = d->blocks; p.n_blocks = d->n_blocks; p.blk_h = d->block_rows; p.blk_w = blk_w;
  p.k_masks = reinterpret_cast<const unsigned long long*>(d->k_masks); p.k_cell_kb = d->k_masks ? (int)(d->k_cell / tc::BK) : 0;
  p.add_a = d->add_a; p.add_a_rows = d->add_a_rows; p.add_b = d->add_b; p.add_b_rows = d->add_b_rows; p.ld_add = d->ld_add;
  p.out_rows = d->out_rows;
  p.m_order = d->m_order;
  HC_REQUIRE(!d->m_order || d->mode == HC_GEMM_PLAIN, HC_E_SHAPE, "hc_tc_gemm: m_order needs a plain GEMM");
  p.diff_sub = reinterpret_cast<const __nv_bfloat16*>(d->diff_sub); p.diff_obj = reinterpret_cast<const __nv_bfloat16*>(d->diff_obj);
  p.diff_bg = reinterpret_cast<const __nv_bfloat16*>(d->diff_bg);
  p.pair_sub = d->pair_sub; p.pair_obj = d->pair_obj; p.pair_row = d->pair_row;
  HC_REQUIRE(!d->mul || (!pooled && d->mode == HC_GEMM_PLAIN && d->ld_mul % 4 == 0 && aligned16(d->mul)), HC_E_SHAPE,
             "hc_tc_gemm: mul needs a plain GEMM, a non-pooled epilogue and a 16-byte aligned [M, ld_mul] f32 operand");
  p.tiles_n = p.N / BN;

  CUtensorMap ta, tb, tbh;
  {
    cuuint64_t dims[2] = {(cuuint64_t)d->k, (cuuint64_t)d->n};
    cuuint64_t str[1] = {(cuuint64_t)d->k * 2};
    cuuint32_t box[2] = {(cuuint32_t)tc::BK, (cuuint32_t)BN};
    rc = tc::make_map(&tb, d->b, 2, dims, str, box, d->operand_f16 ? 1 : 0);
    if (rc != HC_OK) return rc;
    box[1] = (cuuint32_t)(BN / 2);            // half tile: what one CTA of a pair multicasts
    rc = tc::make_map(&tbh, d->b, 2, dims, str, box, d->operand_f16 ? 1 : 0);
    if (rc != HC_OK) return rc;
  }
  const bool blk = d->mode == HC_GEMM_CONV3_BLOCKS;
  if (d->mode == HC_GEMM_CONV3 || blk) {
    HC_REQUIRE(d->h > 0 && d->w > 0 && d->n_img > 0, HC_E_SHAPE, "hc_tc_gemm: conv needs n_img,h,w");
    HC_REQUIRE(blk ? (d->w >= 8 && d->w <= 32 && d->h >= d->block_rows && d->h <= 32 && d->n_img < (1 << 23))
                   : (d->w % 16 == 0 && d->h % (8 * MS) == 0),
               HC_E_SHAPE, "hc_tc_gemm: conv H,W must be multiples of the 16x8 tile (block mode: 8 <= W,H <= 32, n_img < 2^23)");
    HC_REQUIRE(d->c_in % tc::BK == 0 && d->c_total % 8 == 0 && d->c_base % 8 == 0 && d->c_base + d->c_in <= d->c_total, HC_E_SHAPE,
               "hc_tc_gemm: conv channel slice must be 64-aligned inside c_total");
    HC_REQUIRE(d->k == 9ll * d->c_in, HC_E_SHAPE, "hc_tc_gemm: conv needs K == 9*c_in");
    HC_REQUIRE(d->m == (int64_t)d->n_img * d->h * d->w, HC_E_SHAPE, "hc_tc_gemm: conv needs M == n_img*H*W");
    HC_REQUIRE(!pooled || (d->h % 2 == 0 && d->w % 2 == 0), HC_E_SHAPE, "hc_tc_gemm: pooling needs even H,W");
    p.H = d->h; p.W = d->w; p.c_in = d->c_in; p.c_base = d->c_base;
    p.tiles_x = d->w / 16; p.tiles_y = d->h / (8 * MS);
    p.tiles_m = d->n_img * p.tiles_x * p.tiles_y;
    cuuint64_t dims[4] = {(cuuint64_t)d->c_total, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->n_img};
    cuuint64_t str[3] = {(cuuint64_t)d->c_total * 2, (cuuint64_t)d->w * d->c_total * 2, (cuuint64_t)d->h * d->w * d->c_total * 2};
    // dense: one patch serves the three ky taps; block mode: one {64 ch, block_cols x, block_rows y} box per block and tap
    cuuint32_t box[4] = {(cuuint32_t)tc::BK, blk ? (cuuint32_t)blk_w : 16u, (cuuint32_t)(blk ? d->block_rows : 8 * MS + 2), 1};
    rc = tc::make_map(&ta, d->a, 4, dims, str, box, d->operand_f16 ? 1 : 0);
    if (rc != HC_OK) return rc;
  } else {
    HC_REQUIRE(d->mode == HC_GEMM_PLAIN, HC_E_SHAPE, "hc_tc_gemm: unknown mode");
    HC_REQUIRE(d->lda >= d->k && d->lda % 8 == 0, HC_E_ALIGN, "hc_tc_gemm: lda must be >= K and a multiple of 8");
    p.tiles_m = (int)((d->m + tc::BM * MS - 1) / (tc::BM * MS));
    cuuint64_t dims[2] = {(cuuint64_t)d->k, (cuuint64_t)d->m};
    cuuint64_t str[1] = {(cuuint64_t)d->lda * 2};
    cuuint32_t box[2] = {(cuuint32_t)tc::BK, (cuuint32_t)tc::BM};
    rc = tc::make_map(&ta, d->a, 2, dims, str, box, d->operand_f16 ? 1 : 0);
    if (rc != HC_OK) return rc;
  }
  p.group_m = d->group_m > 0 ? d->group_m : 1;
  if (blk) p.group_m = 1;                 // the 4 N tiles of an M tile run side by side and share its blocks through L2
  else if (p.group_m > p.tiles_m) p.group_m = p.tiles_m;

  // block mode: tcgen05 cta_group::2 CTA pairs (d->cta_pairs); needs m_sub * N tile == 512
  // TMEM columns per CTA or fewer, which every configuration satisfies
  if (blk) {
    // the pair kernel is built for conv3_1's shape: 4x4-pixel blocks, pooled epilogues, 512-channel tiles, 256-row CTA tiles
    const bool diff = d->epilogue == HC_EPI_POOL_DIFF_BF16;
    const bool pair_ok = pooled && blk_w == 4 && (d->block_rows == 4 || d->block_rows == 2) && d->n % 512 == 0 && MS == 2 && d->c_off == 0 &&
                         d->ldc == d->n && (!diff || (d->scratch && aligned16(d->scratch)));
    p.cl2 = (pair_ok && d->cta_pairs) ? 1 : 0;
    HC_REQUIRE(d->block_rows != 2 || p.cl2, HC_E_SHAPE,
               "hc_tc_gemm: 4x2-pixel blocks run on the CTA-pair kernel only (pooled epilogue, m_sub 2, N % 512 == 0, c_off 0, ldc == N, scratch "
               "for the difference epilogue)");
    static int env_dbg = -1;
    if (env_dbg == -1) { const char* e = getenv("HC_TC_DEBUG"); env_dbg = e ? atoi(e) : 0; }
    p.dbg = env_dbg;
  }
  if (p.cl2) {
    const bool diff = d->epilogue == HC_EPI_POOL_DIFF_BF16;
    if (diff) p.out = d->scratch;               // pooled values by local pair; pair_diff_kernel writes the differences to d->out
    rc = tc::launch<256, 2, true>(ta, tb, tbh, p, stream);
    if (rc != HC_OK || !diff) return rc;
    const long long cell_vec = d->ldc / 8, map_vec = (long long)(d->h / 2) * (d->w / 2) * cell_vec;
    tc::pair_diff_kernel<<<num_sms() * 8, 256, 0, stream>>>(
        d->blocks, d->n_blocks, reinterpret_cast<const uint4*>(d->scratch), reinterpret_cast<const uint4*>(d->diff_sub),
        reinterpret_cast<const uint4*>(d->diff_obj), reinterpret_cast<const uint4*>(d->diff_bg), d->pair_sub, d->pair_obj, d->pair_row,
        blk_w / 2, d->block_rows / 2, d->w / 2, cell_vec, map_vec, reinterpret_cast<uint4*>(d->out), p.f16);
    return cuda_status("pair_diff_kernel launch");
  }
  // plain GEMM on tcgen05 cta_group::2 pairs (d->cta_pairs, N % 256 == 0): a pair owns m_sub 256-row units of one N tile, each CTA staging
  // its 128 rows of every unit and its 128 columns of B per K step.  m_sub 1: two accumulator stages, the epilogue overlaps the next
  // tile.  m_sub 2: both units' accumulators live in TMEM and share every staged weight block; with k_masks (one mask per 256-row
  // unit either way) the pair walks the union of the two units' cells and a unit skips the cells it does not have
  if (d->mode == HC_GEMM_PLAIN && d->cta_pairs) {
    HC_REQUIRE(BN == 256 && d->epilogue != HC_EPI_SPLIT3_BF16, HC_E_SHAPE,
               "hc_tc_gemm: plain-GEMM CTA pairs need N % 256 == 0 and a bf16 / f32 epilogue");
    p.cl2 = 1;
    p.tiles_m = (int)((d->m + 2 * tc::BM * MS - 1) / (2 * tc::BM * MS));       // pair tiles of MS 256-row units
    if (p.group_m > p.tiles_m) p.group_m = p.tiles_m;
    if (MS == 2) return tc::launch<256, 2, true>(ta, tb, tbh, p, stream);
    return tc::launch<256, 1, true>(ta, tb, tbh, p, stream);
  }
  if (BN == 256 && MS == 1) return tc::launch<256, 1, false>(ta, tb, tbh, p, stream);
  if (BN == 256 && MS == 2) return tc::launch<256, 2, false>(ta, tb, tbh, p, stream);
  if (BN == 128 && MS == 1) return tc::launch<128, 1, false>(ta, tb, tbh, p, stream);
  return tc::launch<128, 2, false>(ta, tb, tbh, p, stream);
}
