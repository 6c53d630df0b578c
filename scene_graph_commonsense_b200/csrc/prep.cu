// HBM-bound layout / gather stages around the dense kernels (R3 and the conv2 split):
//   pack_pixels     NCHW f32 maps -> pixel-major bf16 rows (A operand of the 1x1 convolutions)
//   box_select      per-box masked conv1 activations without ever materialising feature*mask
//   pair_relu_pool  relu(U[sub] + V[obj] + b) + 2x2 max-pool per directed pair
// All three are pure streaming kernels: 16-byte vector accesses, channel index fastest so a warp touches
// 512 contiguous bytes.
#include <cuda_fp16.h>

#include "hc_common.cuh"

namespace hc {

__global__ void pack_pixels_kernel(const float* __restrict__ src0, int c0, const float* __restrict__ src1, int c1, int hw, int k_pad,
                                   unsigned short* __restrict__ out, int f16) {
  __shared__ float tile[32][33];
  const int img = blockIdx.z;
  const int p0 = blockIdx.x * 32, ch0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = ch0 + i, p = p0 + threadIdx.x;
    float v = 0.0f;
    if (p < hw) {
      if (c < c0) v = src0[((size_t)img * c0 + c) * hw + p];
      else if (c < c0 + c1) v = src1[((size_t)img * c1 + (c - c0)) * hw + p];
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int p = p0 + i, c = ch0 + threadIdx.x;
    if (p < hw && c < k_pad) {
      const float x = tile[threadIdx.x][i];
      unsigned short bits;
      if (f16) { __half h = __float2half_rn(fminf(fmaxf(x, -65504.0f), 65504.0f)); bits = *reinterpret_cast<unsigned short*>(&h); }
      else { __nv_bfloat16 b = __float2bfloat16_rn(x); bits = *reinterpret_cast<unsigned short*>(&b); }
      out[((size_t)img * hw + p) * k_pad + c] = bits;
    }
  }
}

// One CTA per (box, band of BS_ROWS pixel rows): the rectangle test is block-uniform per pixel row, every thread moves
// 16-byte vectors with no integer division, a warp writes 512 contiguous bytes.
constexpr int BS_ROWS = 4;

__global__ void __launch_bounds__(256)
box_select_kernel(const uint4* __restrict__ t_img, const int4* __restrict__ boxes, const int* __restrict__ box_img, int fs, int cvec,
                  const uint4* __restrict__ fill, uint4* __restrict__ out) {
  const int box = blockIdx.y;
  const int y0 = blockIdx.x * BS_ROWS;
  const Rect r = rect_of(boxes[box], fs);
  const long long src_base = (long long)box_img[box] * fs * fs * cvec;
  const long long dst_base = (long long)box * fs * fs * cvec;
  const int row_vec = fs * cvec;                          // vectors per pixel row
  for (int dy = 0; dy < BS_ROWS && y0 + dy < fs; ++dy) {
    const int y = y0 + dy;
    const bool row_in = y >= r.y0 && y < r.y1;
    const int v0 = r.x0 * cvec, v1 = r.x1 * cvec;         // inside <=> v0 <= vector index < v1 on this row
    const long long off = (long long)y * row_vec;
    for (int i = threadIdx.x; i < row_vec; i += blockDim.x) {
      const bool inside = row_in && i >= v0 && i < v1;
      int cv = i % cvec;                                  // cvec is a power of two on this path (256 channels -> 32)
      out[dst_base + off + i] = inside ? __ldg(t_img + src_base + off + i) : __ldg(fill + cv);
    }
  }
}

__device__ __forceinline__ void add8(float (&acc)[8], uint4 a, uint4 b, const float (&bias)[8]) {
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float2 fa = __bfloat1622float2(pa[k]), fb = __bfloat1622float2(pb[k]);
    acc[2 * k] = fmaxf(acc[2 * k], fa.x + (fb.x + bias[2 * k]));          // same association as the tiled kernel
    acc[2 * k + 1] = fmaxf(acc[2 * k + 1], fa.y + (fb.y + bias[2 * k + 1]));
  }
}

__global__ void pair_relu_pool_kernel(const uint4* __restrict__ u, const uint4* __restrict__ v, const float* __restrict__ bias,
                                      const int* __restrict__ pair_sub, const int* __restrict__ pair_obj, long long total_vec, int fs,
                                      int cvec, uint4* __restrict__ out) {
  const int hp = fs / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
    int cv = (int)(i % cvec);
    long long pix = i / cvec;
    int px = (int)(pix % hp);
    int py = (int)((pix / hp) % hp);
    int pr = (int)(pix / ((long long)hp * hp));
    const long long su = (long long)pair_sub[pr] * fs * fs, so = (long long)pair_obj[pr] * fs * fs;
    float b[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) b[k] = __ldg(bias + cv * 8 + k);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.0f;            // relu folded in: max(0, ...)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        long long off = ((long long)(2 * py + dy) * fs + (2 * px + dx)) * cvec + cv;
        add8(acc, __ldg(u + su * cvec + off), __ldg(v + so * cvec + off), b);
      }
    uint4 o;
    __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) po[k] = __floats2bfloat162_rn(acc[2 * k], acc[2 * k + 1]);
    out[i] = o;
  }
}

// ---- tiled variant: out[(a,b)] = pool(relu(U[a] + V[b] + bias)) is an OUTER SUM over the boxes of one image, so a thread
// keeps the U tiles of TA subject boxes in registers (packed bf16) and streams V[b] of every object box once:
// L2/HBM reads per pair drop from 2 MiB to ~(1/TA + 1/N) MiB.  Output rows are found through a (subject box, local
// object index) -> directed-pair-index table built from the enumerated pair list.
constexpr int PP_TA = 4;

__global__ void pair_lut_kernel(const int* __restrict__ pair_sub, const int* __restrict__ pair_obj, const int* __restrict__ pair_img,
                                const int* __restrict__ box_off, int n_pairs, int n_max, int* __restrict__ lut) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += gridDim.x * blockDim.x)
    lut[(long long)pair_sub[p] * n_max + (pair_obj[p] - box_off[pair_img[p]])] = p;
}

__device__ __forceinline__ void unpack8(uint4 a, float (&f)[8]) {
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    f[2 * k] = __uint_as_float(w[k] << 16);            // bf16 -> f32 is a 16-bit shift
    f[2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u);
  }
}

// same conversion, but opaque to loop-invariant code motion: the U tiles must stay PACKED in registers across the
// object loop (hoisting the unpack would need 128 live floats per thread and spill)
__device__ __forceinline__ void unpack8_pinned(uint4 a, float (&f)[8]) {
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint32_t lo, hi;
    asm volatile("shl.b32 %0, %1, 16;" : "=r"(lo) : "r"(w[k]));
    asm volatile("and.b32 %0, %1, 0xFFFF0000;" : "=r"(hi) : "r"(w[k]));
    f[2 * k] = __uint_as_float(lo);
    f[2 * k + 1] = __uint_as_float(hi);
  }
}

__global__ void __launch_bounds__(128, 4)
pair_relu_pool_tiled_kernel(const uint4* __restrict__ u, const uint4* __restrict__ v, const float* __restrict__ bias,
                            const int* __restrict__ box_off, const int* __restrict__ lut, int n_max, int img0, int pair_base,
                            int chunk_pairs, int fs, int cvec, int tiles_per_img, int slabs, uint4* __restrict__ out) {
  // grid = (slabs, tiles_per_img, images of the chunk); block = 128 consecutive (pooled pixel, channel vector) slots
  const int img = img0 + blockIdx.z;
  const int b0 = box_off[img], n = box_off[img + 1] - b0;
  const int a0 = blockIdx.y * PP_TA;
  if (a0 >= n) return;
  const int hp = fs / 2;
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  const int cv = slot % cvec;
  const int pix = slot / cvec;
  const int px = pix % hp, py = pix / hp;
  if (py >= hp) return;
  float bia[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) bia[k] = __ldg(bias + cv * 8 + k);
  uint4 ua[PP_TA][4];
#pragma unroll
  for (int t = 0; t < PP_TA; ++t) {
    const int a = min(a0 + t, n - 1);
    const long long base = (long long)(b0 + a) * fs * fs;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      ua[t][q] = __ldg(u + (base + (long long)(2 * py + (q >> 1)) * fs + (2 * px + (q & 1))) * cvec + cv);
  }
  for (int b = 0; b < n; ++b) {
    int prow[PP_TA];
    bool any = false;
#pragma unroll
    for (int t = 0; t < PP_TA; ++t) {
      int p = (a0 + t < n) ? __ldg(lut + (long long)(b0 + a0 + t) * n_max + b) : -1;
      p = (p >= 0) ? p - pair_base : -1;
      if (p >= chunk_pairs) p = -1;
      prow[t] = p;
      any |= p >= 0;
    }
    if (!any) continue;                                  // block-uniform (lut entries do not depend on the thread)
    const long long vb = (long long)(b0 + b) * fs * fs;
    float acc[PP_TA][8];
#pragma unroll
    for (int t = 0; t < PP_TA; ++t)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[t][k] = 0.0f;       // relu folded in: max(0, ...)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float vf[8];
      unpack8(__ldg(v + (vb + (long long)(2 * py + (q >> 1)) * fs + (2 * px + (q & 1))) * cvec + cv), vf);
#pragma unroll
      for (int k = 0; k < 8; ++k) vf[k] += bia[k];
#pragma unroll
      for (int t = 0; t < PP_TA; ++t) {
        float uf[8];
        unpack8_pinned(ua[t][q], uf);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[t][k] = fmaxf(acc[t][k], uf[k] + vf[k]);
      }
    }
#pragma unroll
    for (int t = 0; t < PP_TA; ++t) {
      if (prow[t] < 0) continue;
      uint4 o;
      __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int k = 0; k < 4; ++k) po[k] = __floats2bfloat162_rn(acc[t][2 * k], acc[t][2 * k + 1]);
      out[((long long)prow[t] * hp * hp + pix) * cvec + cv] = o;
    }
  }
}

// ---- packed-bf16 variants (bias == NULL): V already carries the conv2 bias (it is added in fp32 inside the object-half
// GEMM epilogue before the single rounding to bf16), so a pair is  relu(max_2x2(bf16(U + V))).  add.rn.bf16x2 rounds the
// exact sum once, which is what rounding the fp32 sum of two bf16 values gives, and rounding commutes with max and relu:
// the result equals the fp32 formulation bit for bit while issuing ~1/3 of the instructions (2 per two elements instead of
// unpack + add + max + repack), which is what moves the kernel from issue-bound to HBM-bound.
// F16 = the fp16 operand format (add.rn.f16x2 / max.f16x2): the same two packed instructions per two elements
template <bool F16>
__device__ __forceinline__ uint32_t bf2_add(uint32_t a, uint32_t b) {
  if (F16) {
    __half2 r = __hadd2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  __nv_bfloat162 r = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
template <bool F16>
__device__ __forceinline__ uint32_t bf2_max(uint32_t a, uint32_t b) {
  if (F16) {
    __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
template <bool F16>
__device__ __forceinline__ uint4 bf8_add(uint4 a, uint4 b) {
  return make_uint4(bf2_add<F16>(a.x, b.x), bf2_add<F16>(a.y, b.y), bf2_add<F16>(a.z, b.z), bf2_add<F16>(a.w, b.w));
}
template <bool F16>
__device__ __forceinline__ uint4 bf8_max(uint4 a, uint4 b) {
  return make_uint4(bf2_max<F16>(a.x, b.x), bf2_max<F16>(a.y, b.y), bf2_max<F16>(a.z, b.z), bf2_max<F16>(a.w, b.w));
}

// cover (optional, fs == 32): per pair, the 8x8-grid cells its listed conv3_1 blocks cover; a pooled pixel none of whose 3x3
// neighbourhood cells is covered is never read by HC_GEMM_CONV3_BLOCKS and is skipped (neither loaded nor stored)
template <bool F16>
__global__ void pair_relu_pool_bf16_kernel(const uint4* __restrict__ u, const uint4* __restrict__ v, const int* __restrict__ pair_sub,
                                           const int* __restrict__ pair_obj, long long total_vec, int fs, int cvec,
                                           const unsigned long long* __restrict__ cover, const int4* __restrict__ fp_boxes,
                                           const uint4* __restrict__ u_bg, const uint4* __restrict__ v_bg, int fp_bh,
                                           uint4* __restrict__ out) {
  const int hp = fs / 2;
  const int per_pair = hp * hp * cvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
    const int pr = (int)(i / per_pair);
    const int rem = (int)(i - (long long)pr * per_pair);
    const int cv = rem % cvec, pix = rem / cvec;
    const int px = pix % hp, py = pix / hp;
    if (cover) {
      const int cy0 = max(py - 1, 0) >> 1, cy1 = min(py + 1, hp - 1) >> 1, cx0 = max(px - 1, 0) >> 1, cx1 = min(px + 1, hp - 1) >> 1;
      const unsigned long long rowbits = (cx1 > cx0 ? 3ull : 1ull) << cx0;
      const unsigned long long nbr = (rowbits << (8 * cy0)) | (cy1 > cy0 ? rowbits << (8 * cy1) : 0ull);
      if (!(__ldg(cover + pr) & nbr)) continue;
    }
    const int bs = pair_sub[pr], bo = pair_obj[pr];
    const long long su = (long long)bs * fs * fs, so = (long long)bo * fs * fs;
    // footprint maps (fp_boxes): U / V hold a box's values only inside its conv2_1 footprint rectangle, the background maps apply elsewhere
    Rect ru = {0, fs, 0, fs}, rv = {0, fs, 0, fs};
    if (fp_boxes) { ru = conv2_valid_rect(__ldg(fp_boxes + bs), fs, fp_bh); rv = conv2_valid_rect(__ldg(fp_boxes + bo), fs, fp_bh); }
    uint4 acc = make_uint4(0u, 0u, 0u, 0u);                // +0.0 in every lane: relu folded into the running max
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int y = 2 * py + (q >> 1), x = 2 * px + (q & 1);
      const long long off = ((long long)y * fs + x) * cvec + cv;
      const uint4* up = rect_has(ru, x, y) ? u + su * cvec : u_bg;
      const uint4* vp = rect_has(rv, x, y) ? v + so * cvec : v_bg;
      acc = bf8_max<F16>(acc, bf8_add<F16>(__ldg(up + off), __ldg(vp + off)));
    }
    out[i] = acc;
  }
}

template <bool F16>
__global__ void __launch_bounds__(128, 4)
pair_relu_pool_tiled_bf16_kernel(const uint4* __restrict__ u, const uint4* __restrict__ v, const int* __restrict__ box_off,
                                 const int* __restrict__ lut, int n_max, int img0, int pair_base, int chunk_pairs, int fs, int cvec,
                                 const unsigned long long* __restrict__ cover, const int4* __restrict__ fp_boxes,
                                 const uint4* __restrict__ u_bg, const uint4* __restrict__ v_bg, int fp_bh, uint4* __restrict__ out) {
  // grid = (slabs, subject tiles, images of the chunk); block = 128 consecutive (pooled pixel, channel vector) slots.
  // The (subject, object) -> pair row table and the pairs' cover words are the same for every thread of the block: they are staged
  // in shared memory once (two dependent global loads per (subject, object) and thread became the latency chain of this kernel
  // once the footprint cover had removed most of the stores, profiles/ncu_summary_r01L.json).
  extern __shared__ unsigned long long pp_smem[];
  unsigned long long* s_cov = pp_smem;                                       // [PP_TA][n_max] cover word (0 = pair absent / not in chunk)
  int* s_row = reinterpret_cast<int*>(pp_smem + PP_TA * n_max);              // [PP_TA][n_max] chunk-local pair row
  int4* s_rect = reinterpret_cast<int4*>(pp_smem + PP_TA * n_max + (PP_TA * n_max + 1) / 2);   // [n_max] object footprint rectangles (fp_boxes)
  const int img = img0 + blockIdx.z;
  const int b0 = box_off[img], n = box_off[img + 1] - b0;
  const int a0 = blockIdx.y * PP_TA;
  if (a0 >= n) return;                                                       // block-uniform
  for (int i = threadIdx.x; i < PP_TA * n; i += blockDim.x) {
    const int t = i / n, b = i - t * n;
    int p = (a0 + t < n) ? __ldg(lut + (long long)(b0 + a0 + t) * n_max + b) : -1;
    p = (p >= 0) ? p - pair_base : -1;
    if (p >= chunk_pairs) p = -1;
    s_row[t * n_max + b] = p;
    s_cov[t * n_max + b] = p < 0 ? 0ull : (cover ? __ldg(cover + p) : ~0ull);
  }
  if (fp_boxes)
    for (int b = threadIdx.x; b < n; b += blockDim.x) {
      const Rect r = conv2_valid_rect(__ldg(fp_boxes + b0 + b), fs, fp_bh);
      s_rect[b] = make_int4(r.x0, r.x1, r.y0, r.y1);
    }
  __syncthreads();
  const int hp = fs / 2;
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  const int cv = slot % cvec;
  const int pix = slot / cvec;
  const int px = pix % hp, py = pix / hp;
  if (py >= hp) return;
  long long qoff[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) qoff[q] = ((long long)(2 * py + (q >> 1)) * fs + (2 * px + (q & 1))) * cvec + cv;
  // cover: this pooled pixel is read by a listed conv3_1 block iff one of the cells its 3x3 neighbourhood touches is covered
  // (fs = 32: a cell is 2 x 2 pooled pixels); warp-uniform (a warp holds one pixel)
  unsigned long long nbr = ~0ull;
  if (cover) {
    const int cy0 = max(py - 1, 0) >> 1, cy1 = min(py + 1, hp - 1) >> 1, cx0 = max(px - 1, 0) >> 1, cx1 = min(px + 1, hp - 1) >> 1;
    nbr = 0ull;
    for (int cy = cy0; cy <= cy1; ++cy)
      for (int cx = cx0; cx <= cx1; ++cx) nbr |= 1ull << (8 * cy + cx);
  }
  // subjects of the tile that need this pixel for at least one object: their U tiles are fetched, the others are not
  unsigned need_t = 0u;
  for (int b = 0; b < n; ++b) {
#pragma unroll
    for (int t = 0; t < PP_TA; ++t) need_t |= (unsigned)((s_cov[t * n_max + b] & nbr) != 0ull) << t;
  }
  if (!need_t) return;                                                       // warp-uniform
  uint4 ua[PP_TA][4];
#pragma unroll
  for (int t = 0; t < PP_TA; ++t) {
    if (!((need_t >> t) & 1u)) continue;
    const long long base = (long long)(b0 + min(a0 + t, n - 1)) * fs * fs * cvec;
    Rect ru = {0, fs, 0, fs};
    if (fp_boxes) { const int4 r = s_rect[min(a0 + t, n - 1)]; ru.x0 = r.x; ru.x1 = r.y; ru.y0 = r.z; ru.y1 = r.w; }
#pragma unroll
    for (int q = 0; q < 4; ++q) ua[t][q] = __ldg((rect_has(ru, 2 * px + (q & 1), 2 * py + (q >> 1)) ? u + base : u_bg) + qoff[q]);
  }
  const long long out_slot = (long long)pix * cvec + cv;
  const long long pair_stride = (long long)hp * hp * cvec;
  for (int b = 0; b < n; ++b) {
    int prow[PP_TA];
    bool any = false;
#pragma unroll
    for (int t = 0; t < PP_TA; ++t) {
      prow[t] = (s_cov[t * n_max + b] & nbr) ? s_row[t * n_max + b] : -1;
      any |= prow[t] >= 0;
    }
    if (!any) continue;                                  // uniform per warp
    const long long vb = (long long)(b0 + b) * fs * fs * cvec;
    Rect rv = {0, fs, 0, fs};
    if (fp_boxes) { const int4 r = s_rect[b]; rv.x0 = r.x; rv.x1 = r.y; rv.y0 = r.z; rv.y1 = r.w; }
    uint4 vq[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) vq[q] = __ldg((rect_has(rv, 2 * px + (q & 1), 2 * py + (q >> 1)) ? v + vb : v_bg) + qoff[q]);
#pragma unroll
    for (int t = 0; t < PP_TA; ++t) {
      if (prow[t] < 0) continue;                         // warp-uniform
      uint4 acc = make_uint4(0u, 0u, 0u, 0u);            // relu folded into the running max
#pragma unroll
      for (int q = 0; q < 4; ++q) acc = bf8_max<F16>(acc, bf8_add<F16>(ua[t][q], vq[q]));
      __stcs(out + (long long)prow[t] * pair_stride + out_slot, acc);   // streamed: next read is conv3's TMA, after the chunk
    }
  }
}

}  // namespace hc

using namespace hc;

static int stream_grid(long long total, int block) {
  long long g = (total + block - 1) / block;
  long long cap = (long long)num_sms() * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

extern "C" int hc_pack_pixels(const float* src0, int32_t c0, const float* src1, int32_t c1, int32_t n_img, int32_t hw, int32_t k_pad,
                              void* out_bf16, int32_t operand_f16, hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(src0 && out_bf16, HC_E_NULL, "hc_pack_pixels: NULL pointer");
  HC_REQUIRE(n_img > 0 && hw > 0 && c0 > 0 && c1 >= 0 && k_pad >= c0 + c1, HC_E_SHAPE, "hc_pack_pixels: bad sizes");
  HC_REQUIRE(c1 == 0 || src1, HC_E_NULL, "hc_pack_pixels: src1 is NULL but c1 > 0");
  HC_REQUIRE(n_img <= 65535, HC_E_SHAPE, "hc_pack_pixels: more than 65535 images per call");
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  dim3 grid((hw + 31) / 32, (k_pad + 31) / 32, n_img), block(32, 8);
  pack_pixels_kernel<<<grid, block, 0, stream>>>(src0, c0, src1, c1, hw, k_pad, reinterpret_cast<unsigned short*>(out_bf16), operand_f16 ? 1 : 0);
  return cuda_status("hc_pack_pixels");
}

extern "C" int hc_box_select(const void* t_img, const int32_t* boxes, const int32_t* box_img, int32_t n_box, int32_t fs,
                             int32_t channels, const void* fill, void* out, hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(t_img && boxes && box_img && fill && out, HC_E_NULL, "hc_box_select: NULL pointer");
  HC_REQUIRE(n_box > 0 && fs > 0 && channels > 0 && channels % 8 == 0, HC_E_SHAPE, "hc_box_select: channels must be a multiple of 8");
  HC_REQUIRE(aligned16(t_img) && aligned16(boxes) && aligned16(fill) && aligned16(out), HC_E_ALIGN, "hc_box_select: 16-byte alignment");
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  HC_REQUIRE(n_box <= 65535, HC_E_SHAPE, "hc_box_select: more than 65535 boxes per call");
  dim3 grid((fs + BS_ROWS - 1) / BS_ROWS, n_box);
  box_select_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(t_img), reinterpret_cast<const int4*>(boxes), box_img, fs,
                                              channels / 8, reinterpret_cast<const uint4*>(fill), reinterpret_cast<uint4*>(out));
  return cuda_status("hc_box_select");
}

static int check_footprint(const hc_uv_footprint* fp, const float* bias, int32_t fs, const char* who) {
  if (!fp) return HC_OK;
  HC_REQUIRE(fp->boxes && fp->u_bg && fp->v_bg && aligned16(fp->boxes) && aligned16(fp->u_bg) && aligned16(fp->v_bg), HC_E_NULL, who);
  HC_REQUIRE(!bias && fs >= 8 && (fp->block_rows == 4 || fp->block_rows == 8), HC_E_SHAPE, who);
  return HC_OK;
}

extern "C" int hc_pair_relu_pool(const void* u, const void* v, const float* bias, const int32_t* pair_sub, const int32_t* pair_obj,
                                 int32_t n_pairs, int32_t fs, int32_t channels, const uint64_t* cover, const hc_uv_footprint* fp, void* out,
                                 int32_t operand_f16, hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(u && v && pair_sub && pair_obj && out, HC_E_NULL, "hc_pair_relu_pool: NULL pointer");
  HC_REQUIRE(!operand_f16 || !bias, HC_E_SHAPE, "hc_pair_relu_pool: fp16 operands take the packed path (bias == NULL)");
  HC_REQUIRE(n_pairs > 0 && fs > 0 && fs % 2 == 0 && channels % 8 == 0, HC_E_SHAPE, "hc_pair_relu_pool: bad sizes");
  HC_REQUIRE(!cover || (!bias && fs == 32), HC_E_SHAPE, "hc_pair_relu_pool: the footprint cover needs the packed path (bias == NULL) and feature_size 32");
  {
    int frc = check_footprint(fp, bias, fs, "hc_pair_relu_pool: footprint maps need boxes + both background maps (16-byte aligned), bias == NULL, block_rows 4 or 8");
    if (frc != HC_OK) return frc;
  }
  const int4* fpb = fp ? reinterpret_cast<const int4*>(fp->boxes) : nullptr;
  const uint4* fpu = fp ? reinterpret_cast<const uint4*>(fp->u_bg) : nullptr;
  const uint4* fpv = fp ? reinterpret_cast<const uint4*>(fp->v_bg) : nullptr;
  const int fph = fp ? fp->block_rows : 0;
  HC_REQUIRE(aligned16(u) && aligned16(v) && aligned16(out), HC_E_ALIGN, "hc_pair_relu_pool: 16-byte alignment");
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  long long total = (long long)n_pairs * (fs / 2) * (fs / 2) * (channels / 8);
  if (!bias) {
    if (operand_f16)
      pair_relu_pool_bf16_kernel<true><<<stream_grid(total, 256), 256, 0, stream>>>(reinterpret_cast<const uint4*>(u), reinterpret_cast<const uint4*>(v),
                                                                                    pair_sub, pair_obj, total, fs, channels / 8,
                                                                                    reinterpret_cast<const unsigned long long*>(cover),
                                                                                    fpb, fpu, fpv, fph, reinterpret_cast<uint4*>(out));
    else
      pair_relu_pool_bf16_kernel<false><<<stream_grid(total, 256), 256, 0, stream>>>(reinterpret_cast<const uint4*>(u), reinterpret_cast<const uint4*>(v),
                                                                                     pair_sub, pair_obj, total, fs, channels / 8,
                                                                                     reinterpret_cast<const unsigned long long*>(cover),
                                                                                     fpb, fpu, fpv, fph, reinterpret_cast<uint4*>(out));
    return cuda_status("hc_pair_relu_pool");
  }
  pair_relu_pool_kernel<<<stream_grid(total, 256), 256, 0, stream>>>(reinterpret_cast<const uint4*>(u), reinterpret_cast<const uint4*>(v),
                                                                     bias, pair_sub, pair_obj, total, fs, channels / 8,
                                                                     reinterpret_cast<uint4*>(out));
  return cuda_status("hc_pair_relu_pool");
}

extern "C" int hc_pair_lut_build(const int32_t* pair_sub, const int32_t* pair_obj, const int32_t* pair_img, const int32_t* box_offsets,
                                 int32_t n_pairs, int32_t n_box, int32_t n_max, int32_t* lut, hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(pair_sub && pair_obj && pair_img && box_offsets && lut, HC_E_NULL, "hc_pair_lut_build: NULL pointer");
  HC_REQUIRE(n_box > 0 && n_max > 0 && n_pairs >= 0, HC_E_SHAPE, "hc_pair_lut_build: bad sizes");
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  cudaMemsetAsync(lut, 0xFF, sizeof(int32_t) * (size_t)n_box * n_max, stream);      // -1
  if (n_pairs > 0) pair_lut_kernel<<<stream_grid(n_pairs, 256), 256, 0, stream>>>(pair_sub, pair_obj, pair_img, box_offsets, n_pairs, n_max, lut);
  return cuda_status("hc_pair_lut_build");
}

extern "C" int hc_pair_relu_pool_tiled(const void* u, const void* v, const float* bias, const int32_t* box_offsets, const int32_t* lut,
                                       int32_t n_max, int32_t img0, int32_t n_img, int32_t pair_base, int32_t chunk_pairs, int32_t fs,
                                       int32_t channels, const uint64_t* cover, const hc_uv_footprint* fp, void* out, int32_t operand_f16,
                                       hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(u && v && box_offsets && lut && out, HC_E_NULL, "hc_pair_relu_pool_tiled: NULL pointer");
  HC_REQUIRE(!operand_f16 || !bias, HC_E_SHAPE, "hc_pair_relu_pool_tiled: fp16 operands take the packed path (bias == NULL)");
  HC_REQUIRE(n_img > 0 && n_img <= 65535 && n_max > 0 && chunk_pairs > 0 && fs > 0 && fs % 2 == 0 && channels % 8 == 0, HC_E_SHAPE,
             "hc_pair_relu_pool_tiled: bad sizes");
  HC_REQUIRE(aligned16(u) && aligned16(v) && aligned16(out), HC_E_ALIGN, "hc_pair_relu_pool_tiled: 16-byte alignment");
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  const int cvec = channels / 8, hp = fs / 2;
  const int slots = hp * hp * cvec;
  HC_REQUIRE(slots % 128 == 0, HC_E_SHAPE, "hc_pair_relu_pool_tiled: (fs/2)^2 * channels/8 must be a multiple of 128");
  HC_REQUIRE(!cover || fs == 32, HC_E_SHAPE, "hc_pair_relu_pool_tiled: the footprint cover is defined on the 8x8 cell grid of feature_size 32");
  dim3 grid(slots / 128, (n_max + PP_TA - 1) / PP_TA, n_img);
  if (!bias) {
    const size_t smem = (size_t)PP_TA * n_max * sizeof(unsigned long long) + (size_t)((PP_TA * n_max + 1) / 2) * 8 + (size_t)n_max * 16 + 16;
    HC_REQUIRE(smem <= 48 * 1024, HC_E_SHAPE, "hc_pair_relu_pool_tiled: more than ~750 boxes per image");
    {
      int frc = check_footprint(fp, bias, fs, "hc_pair_relu_pool_tiled: footprint maps need boxes + both background maps (16-byte aligned), bias == NULL, block_rows 4 or 8");
      if (frc != HC_OK) return frc;
    }
    const int4* fpb = fp ? reinterpret_cast<const int4*>(fp->boxes) : nullptr;
    const uint4* fpu = fp ? reinterpret_cast<const uint4*>(fp->u_bg) : nullptr;
    const uint4* fpv = fp ? reinterpret_cast<const uint4*>(fp->v_bg) : nullptr;
    const int fph = fp ? fp->block_rows : 0;
    if (operand_f16)
      pair_relu_pool_tiled_bf16_kernel<true><<<grid, 128, smem, stream>>>(reinterpret_cast<const uint4*>(u), reinterpret_cast<const uint4*>(v),
                                                                       box_offsets, lut, n_max, img0, pair_base, chunk_pairs, fs, cvec,
                                                                       reinterpret_cast<const unsigned long long*>(cover), fpb, fpu, fpv, fph,
                                                                       reinterpret_cast<uint4*>(out));
    else
      pair_relu_pool_tiled_bf16_kernel<false><<<grid, 128, smem, stream>>>(reinterpret_cast<const uint4*>(u), reinterpret_cast<const uint4*>(v),
                                                                        box_offsets, lut, n_max, img0, pair_base, chunk_pairs, fs, cvec,
                                                                        reinterpret_cast<const unsigned long long*>(cover), fpb, fpu, fpv, fph,
                                                                        reinterpret_cast<uint4*>(out));
    return cuda_status("hc_pair_relu_pool_tiled");
  }
  HC_REQUIRE(!cover && !fp, HC_E_SHAPE, "hc_pair_relu_pool_tiled: the footprint cover / footprint maps need the packed path (bias == NULL)");
  pair_relu_pool_tiled_kernel<<<grid, 128, 0, stream>>>(reinterpret_cast<const uint4*>(u), reinterpret_cast<const uint4*>(v), bias,
                                                        box_offsets, lut, n_max, img0, pair_base, chunk_pairs, fs, cvec, grid.y, grid.x,
                                                        reinterpret_cast<uint4*>(out));
  return cuda_status("hc_pair_relu_pool_tiled");
}
