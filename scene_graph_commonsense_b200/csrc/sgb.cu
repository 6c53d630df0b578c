// R14 + N1: the Scene-Graph-Benchmark plug-and-play twin of the path
//   SGB/.../roi_relation_predictors.py:413-459  (pair gather, frequency bias, hierarchical log-softmax)
//   SGB/.../inference.py:246-302                (three candidates per pair, triple score, LLM filter window, re-sort)
//   SGB/.../evaluation/vg/sgg_eval.py:56-99,347-385,528-565 + structures/boxlist_ops.py:54-90 (per-image recall)
// The dense part (post_cat 1024->4096 with the `* union_features` epilogue, BayesHead 4096->54) runs on tc_gemm_kernel.
#include <math.h>

#include <cuda_fp16.h>

#include "hc_common.cuh"

namespace hc {

constexpr int SG_THREADS = 256;
constexpr int SG_MAXR = 128;        // ranked window kept per image (reference evaluates the top 100)
constexpr int SG_NREL = 51;         // 50 predicates + background (SGB vocabulary)

__device__ __forceinline__ float wsum(float v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// bf16x3 operand splitting: x = hi + lo with hi = bf16(x), lo = bf16(x - hi).  A GEMM over the K-concatenated operands
// A' = [A_hi | A_lo | A_hi], B' = [B_hi | B_hi | B_lo] sums the three significant partial products on the bf16 tensor
// cores with fp32 accumulation (error ~2^-16 relative instead of 2^-9), at 3x the (tiny) SGB-tail FLOPs.
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
  __nv_bfloat162* ph = reinterpret_cast<__nv_bfloat162*>(&hi);
  __nv_bfloat162* pl = reinterpret_cast<__nv_bfloat162*>(&lo);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * k], x[2 * k + 1]);
    float2 hf = __bfloat1622float2(h);
    ph[k] = h;
    pl[k] = __floats2bfloat162_rn(x[2 * k] - hf.x, x[2 * k + 1] - hf.y);
  }
}

// prod_rep = cat(head_rep[idx0], tail_rep[idx1]) with edge_rep [n_obj, 2*hidden] = post_emb output viewed as (n_obj, 2, hidden).
// split == 0: out bf16 [n, 2*hidden]; split == 1: out bf16 [n, 3 * 2*hidden] = [hi | lo | hi]; split == 2: out fp16 [n, 2*hidden]
// (saturating at +-65504: the plain-fp16 operand format of hc_tc_gemm, operand_f16 = 1)
__global__ void sgb_pair_gather_kernel(const float* __restrict__ edge_rep, const int* __restrict__ pair_idx, long long total_vec, int hidden,
                                       int split, uint4* __restrict__ out) {
  const int vec_per_row = 2 * hidden / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
    int cv = (int)(i % vec_per_row);
    long long p = i / vec_per_row;
    int col = cv * 8;
    int obj = pair_idx[2 * p + (col >= hidden ? 1 : 0)];
    const float4* src = reinterpret_cast<const float4*>(edge_rep + (long long)obj * 2 * hidden + col);
    float4 a = __ldg(src), b = __ldg(src + 1);
    const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (split == 2) {
      uint32_t w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        __half2 h = __floats2half2_rn(fminf(fmaxf(x[2 * k], -65504.0f), 65504.0f), fminf(fmaxf(x[2 * k + 1], -65504.0f), 65504.0f));
        w[k] = *reinterpret_cast<uint32_t*>(&h);
      }
      out[i] = make_uint4(w[0], w[1], w[2], w[3]);
      continue;
    }
    uint4 hi, lo;
    split8(x, hi, lo);
    if (!split) {
      out[i] = hi;
    } else {
      uint4* row = out + p * 3 * vec_per_row;
      row[cv] = hi; row[vec_per_row + cv] = lo; row[2 * vec_per_row + cv] = hi;
    }
  }
}

// f32 [n, k] (row stride ld) -> bf16 [n, 3k] = [hi | lo | hi]
__global__ void split_bf16x3_kernel(const float* __restrict__ in, long long ld, long long total_vec, int k, uint4* __restrict__ out) {
  const int vec_per_row = k / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
    int cv = (int)(i % vec_per_row);
    long long r = i / vec_per_row;
    const float4* src = reinterpret_cast<const float4*>(in + r * ld + cv * 8);
    float4 a = __ldg(src), b = __ldg(src + 1);
    const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint4 hi, lo;
    split8(x, hi, lo);
    uint4* row = out + r * 3 * vec_per_row;
    row[cv] = hi; row[vec_per_row + cv] = lo; row[2 * vec_per_row + cv] = hi;
  }
}

__device__ __forceinline__ float pick(float v0, float v1, int lane, int a, int b, float other) {
  float r = other;
  if (lane >= a && lane < b) r = v0;
  if (lane + 32 >= a && lane + 32 < b) r = v1;
  return r;
}

// roi_relation_predictors.py:430-459.  logits columns: [0,R) the three heads, [R,R+4) super (background + 3).
__global__ void __launch_bounds__(256)
sgb_hier_softmax_kernel(const float* __restrict__ logits, long long ld, int n_rows, int n_geo, int n_pos, int n_sem,
                        const float* __restrict__ bias_table, int num_obj, const int* __restrict__ pair_pred,
                        const int* __restrict__ label_ids, float* __restrict__ rel, float* __restrict__ sup_out) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int R = n_geo + n_pos + n_sem;
  const float NEG = -INFINITY;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < n_rows; row += gridDim.x * wpb) {
    const float* z = logits + (long long)row * ld;
    float v0 = lane < R + 4 ? z[lane] : 0.f;
    float v1 = lane + 32 < R + 4 ? z[lane + 32] : 0.f;
    float b0 = 0.f, b1 = 0.f;
    if (bias_table) {
      const float* brow = bias_table + ((long long)pair_pred[2 * row] * num_obj + pair_pred[2 * row + 1]) * SG_NREL;
      if (lane < R) b0 = __ldg(brow + label_ids[lane]);
      if (lane + 32 < R) b1 = __ldg(brow + label_ids[lane + 32]);
    }
    const int sa[3] = {0, n_geo, n_geo + n_pos}, sb[3] = {n_geo, n_geo + n_pos, R};
    float sup_logit[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int j = R + k;
      sup_logit[k] = __shfl_sync(0xffffffffu, (j < 32) ? v0 : v1, j & 31);
    }
    if (bias_table) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {            // super_bias = log(sum exp(bias_k)) (:436-447), added to super logits 1..3
        float e = pick(b0, b1, lane, sa[k], sb[k], NEG);
        float s = wsum(e == NEG ? 0.f : expf(e));
        sup_logit[1 + k] += logf(s);
      }
    }
    float sm = fmaxf(fmaxf(sup_logit[0], sup_logit[1]), fmaxf(sup_logit[2], sup_logit[3]));
    float ss = expf(sup_logit[0] - sm) + expf(sup_logit[1] - sm) + expf(sup_logit[2] - sm) + expf(sup_logit[3] - sm);
    float slog = sm + logf(ss);
    if (lane < 4) sup_out[(long long)row * 4 + lane] = (lane == 0 ? sup_logit[0] : lane == 1 ? sup_logit[1] : lane == 2 ? sup_logit[2] : sup_logit[3]) - slog;
    float z0 = v0 + b0, z1 = v1 + b1, o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float x = pick(z0, z1, lane, sa[k], sb[k], NEG);
      float m = wmax(x);
      float e = wsum(x == NEG ? 0.f : expf(x - m));
      float lse = logf(e);
      float sk = sup_logit[1 + k] - slog;
      if (lane >= sa[k] && lane < sb[k]) o0 = (z0 - m - lse) + sk;
      if (lane + 32 >= sa[k] && lane + 32 < sb[k]) o1 = (z1 - m - lse) + sk;
    }
    if (lane < R) rel[(long long)row * R + lane] = o0;
    if (lane + 32 < R) rel[(long long)row * R + lane + 32] = o1;
  }
}

// inference.py:246-281: per pair and head: score = max prob, label = remapped argmax; triple = score * s0 * s1.
// Candidate layout of image i with P_i pairs (torch.cat along dim 0): c = 3*pair_off[i] + k*P_i + r_local.
__global__ void sgb_candidates_kernel(const float* __restrict__ rel, int n_geo, int n_pos, int n_sem, const int* __restrict__ pair_off,
                                      const int* __restrict__ pair_img, const int* __restrict__ pair_idx,
                                      const float* __restrict__ obj_scores, const int* __restrict__ label_ids, int n_rows,
                                      float* __restrict__ cand_score, int* __restrict__ cand_label, int* __restrict__ cand_row) {
  const int R = n_geo + n_pos + n_sem;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += gridDim.x * blockDim.x) {
    const float* x = rel + (long long)r * R;
    const int img = pair_img[r];
    const int p0 = pair_off[img], pn = pair_off[img + 1] - p0;
    const float s0 = obj_scores[pair_idx[2 * r]], s1 = obj_scores[pair_idx[2 * r + 1]];
    const int sa[3] = {0, n_geo, n_geo + n_pos}, sb[3] = {n_geo, n_geo + n_pos, R};
    for (int k = 0; k < 3; ++k) {
      float best = x[sa[k]];
      int arg = sa[k];
      for (int j = sa[k] + 1; j < sb[k]; ++j)
        if (x[j] > best) { best = x[j]; arg = j; }
      float score = expf(best);
      long long c = 3ll * p0 + (long long)k * pn + (r - p0);
      cand_score[c] = __fmul_rn(__fmul_rn(score, s0), s1);
      cand_label[c] = label_ids[arg];
      cand_row[c] = r;
    }
  }
}

// boxlist_ops.py:54-90 in fp32 with the +1 pixel convention, evaluated operation by operation (no FMA contraction)
__device__ __forceinline__ float box_area1(float4 b) { return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f)); }
__device__ __forceinline__ float iou_plus1(float4 a, float4 b) {
  float w = fmaxf(__fadd_rn(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 1.f), 0.f);
  float h = fmaxf(__fadd_rn(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 1.f), 0.f);
  float inter = __fmul_rn(w, h);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(box_area1(a), box_area1(b)), inter));
}

__device__ __forceinline__ uint32_t desc_key32(float f) {
  if (f == 0.0f) f = 0.0f;
  uint32_t u = __float_as_uint(f);
  uint32_t asc = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ~asc;
}

// Second sort of inference.py:292-302 restricted to the ranked window + per-image recall bookkeeping of
// sgg_eval.py:56-99,347-385.  `ranked` [n_img, SG_MAXR] = candidate ids (image-local) in first-sort order
// (score desc, index asc), -1 padded.  `reject` [n_img, SG_MAXR] (optional) marks ranks the validator refused.
struct SgMeta { int s_cls, o_cls, label; float4 sb, ob; };

__global__ void __launch_bounds__(SG_THREADS)
sgb_rank_match_kernel(const int* __restrict__ ranked, const uint8_t* __restrict__ reject, const int* __restrict__ pair_off,
                      const float* __restrict__ cand_score, const int* __restrict__ cand_label, const int* __restrict__ cand_row,
                      const int* __restrict__ pair_idx, const int* __restrict__ pred_cls, const float4* __restrict__ pred_box,
                      const int* __restrict__ gt_off, const int* __restrict__ gt_rel /*[G,3] sub, obj, label (global obj ids)*/,
                      const int* __restrict__ gt_cls, const float4* __restrict__ gt_box, float iou_thresh, int top_max, int k0, int k1,
                      int k2, int* __restrict__ final_rank, int* __restrict__ img_hits, int* __restrict__ img_ngt,
                      int* __restrict__ img_hits_pc, int* __restrict__ img_cnt_pc) {
  __shared__ unsigned long long key[SG_MAXR];
  __shared__ int cand[SG_MAXR];
  __shared__ SgMeta meta[SG_MAXR];
  __shared__ int hits[3], hits_pc[3 * SG_NREL], cnt_pc[SG_NREL];
  const int img = blockIdx.x, tid = threadIdx.x;
  const long long c0 = 3ll * pair_off[img];
  for (int j = tid; j < SG_MAXR; j += SG_THREADS) {
    int c = ranked[(long long)img * SG_MAXR + j];
    unsigned long long kk = ~0ull;
    if (c >= 0) {
      float s = cand_score[c0 + c];
      if (reject && reject[(long long)img * SG_MAXR + j]) s = -INFINITY;      // inference.py:297
      kk = ((unsigned long long)desc_key32(s) << 32) | (uint32_t)j;          // stable: ties keep first-sort rank
    }
    key[j] = kk;
    cand[j] = c;
  }
  for (int i = tid; i < 3; i += SG_THREADS) hits[i] = 0;
  for (int i = tid; i < 3 * SG_NREL; i += SG_THREADS) hits_pc[i] = 0;
  for (int i = tid; i < SG_NREL; i += SG_THREADS) cnt_pc[i] = 0;
  __syncthreads();
  for (int size = 2; size <= SG_MAXR; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (tid < SG_MAXR / 2) {
        int lo = 2 * tid - (tid & (stride - 1)), hi = lo + stride;
        bool up = (lo & size) == 0;
        unsigned long long a = key[lo], b = key[hi];
        if ((a > b) == up) { key[lo] = b; key[hi] = a; }
      }
      __syncthreads();
    }
  int n_sel = 0;
  for (int j = tid; j < SG_MAXR; j += SG_THREADS) {
    unsigned long long kk = key[j];
    int c = kk == ~0ull ? -1 : cand[(int)(kk & 0xFFFFFFFFu)];
    if (j < top_max) {
      if (final_rank) final_rank[(long long)img * top_max + j] = c;
      SgMeta m;
      m.label = -1; m.s_cls = m.o_cls = -1; m.sb = m.ob = make_float4(0, 0, 0, 0);
      if (c >= 0) {
        int row = cand_row[c0 + c];
        int so = pair_idx[2 * row], oo = pair_idx[2 * row + 1];
        m.label = cand_label[c0 + c];
        m.s_cls = pred_cls[so]; m.o_cls = pred_cls[oo];
        m.sb = pred_box[so]; m.ob = pred_box[oo];
      }
      meta[j] = m;
    }
  }
  __syncthreads();
  for (int j = 0; j < top_max && j < SG_MAXR; ++j) n_sel += key[j] != ~0ull;   // valid entries form a dense prefix after the sort
  const int g0 = gt_off[img], G = gt_off[img + 1] - g0;
  const int ks[3] = {k0, k1, k2};
  for (int g = tid; g < G; g += SG_THREADS) {
    const int gs = gt_rel[3 * (g0 + g)], go = gt_rel[3 * (g0 + g) + 1], gl = gt_rel[3 * (g0 + g) + 2];
    const int scls = gt_cls[gs], ocls = gt_cls[go];
    const float4 sbx = gt_box[gs], obx = gt_box[go];
    if (gl >= 0 && gl < SG_NREL) { atomicAdd(&cnt_pc[gl], 1); atomicAdd(&cnt_pc[0], 1); }     // sgg_eval.py:362-365
    int first = -1;
    for (int j = 0; j < n_sel; ++j) {
      const SgMeta& m = meta[j];
      if (m.s_cls != scls || m.label != gl || m.o_cls != ocls) continue;                       // intersect_2d (:534)
      if (iou_plus1(sbx, m.sb) >= iou_thresh && iou_plus1(obx, m.ob) >= iou_thresh) { first = j; break; }
    }
    if (first >= 0)
#pragma unroll
      for (int q = 0; q < 3; ++q)
        if (first < ks[q]) {
          atomicAdd(&hits[q], 1);
          if (gl >= 0 && gl < SG_NREL) { atomicAdd(&hits_pc[q * SG_NREL + gl], 1); atomicAdd(&hits_pc[q * SG_NREL], 1); }
        }
  }
  __syncthreads();
  for (int i = tid; i < 3; i += SG_THREADS) img_hits[img * 3 + i] = hits[i];
  for (int i = tid; i < 3 * SG_NREL; i += SG_THREADS) img_hits_pc[(long long)img * 3 * SG_NREL + i] = hits_pc[i];
  for (int i = tid; i < SG_NREL; i += SG_THREADS) img_cnt_pc[(long long)img * SG_NREL + i] = cnt_pc[i];
  if (tid == 0) img_ngt[img] = G;
}

}  // namespace hc

using namespace hc;

static int sgrid(long long total, int block) {
  long long g = (total + block - 1) / block;
  long long cap = (long long)num_sms() * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

extern "C" int hc_split_bf16x3(const float* in, int64_t ld, int64_t n_rows, int32_t k, void* out, hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(in && out, HC_E_NULL, "hc_split_bf16x3: NULL pointer");
  HC_REQUIRE(n_rows > 0 && k > 0 && k % 8 == 0 && ld >= k && ld % 4 == 0, HC_E_SHAPE, "hc_split_bf16x3: k must be a multiple of 8, ld of 4");
  HC_REQUIRE(aligned16(in) && aligned16(out), HC_E_ALIGN, "hc_split_bf16x3: 16-byte alignment");
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  long long total = n_rows * (k / 8);
  split_bf16x3_kernel<<<sgrid(total, 256), 256, 0, stream>>>(in, ld, total, k, reinterpret_cast<uint4*>(out));
  return cuda_status("hc_split_bf16x3");
}

extern "C" int hc_sgb_pair_gather(const float* edge_rep, const int32_t* pair_idx, int32_t n_pairs, int32_t hidden, int32_t split, void* out,
                                  hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(edge_rep && pair_idx && out, HC_E_NULL, "hc_sgb_pair_gather: NULL pointer");
  HC_REQUIRE(n_pairs > 0 && hidden > 0 && hidden % 8 == 0 && split >= 0 && split <= 2, HC_E_SHAPE,
             "hc_sgb_pair_gather: hidden must be a multiple of 8, split 0 (bf16), 1 (bf16x3) or 2 (fp16)");
  HC_REQUIRE(aligned16(edge_rep) && aligned16(out), HC_E_ALIGN, "hc_sgb_pair_gather: 16-byte alignment");
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  long long total = (long long)n_pairs * (2 * hidden / 8);
  sgb_pair_gather_kernel<<<sgrid(total, 256), 256, 0, stream>>>(edge_rep, pair_idx, total, hidden, split, reinterpret_cast<uint4*>(out));
  return cuda_status("hc_sgb_pair_gather");
}

extern "C" int hc_sgb_hier_softmax(const float* logits, int64_t ld, int32_t n_rows, int32_t n_geo, int32_t n_pos, int32_t n_sem,
                                   const float* bias_table, int32_t num_obj, const int32_t* pair_pred, const int32_t* label_ids,
                                   float* rel, float* super_rel, hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(logits && rel && super_rel, HC_E_NULL, "hc_sgb_hier_softmax: NULL pointer");
  HC_REQUIRE(!bias_table || (pair_pred && label_ids && num_obj > 0), HC_E_NULL, "hc_sgb_hier_softmax: bias needs pair_pred and label_ids");
  HC_REQUIRE(n_geo > 0 && n_pos > 0 && n_sem > 0 && n_geo + n_pos + n_sem + 4 <= 64 && ld >= n_geo + n_pos + n_sem + 4, HC_E_SHAPE,
             "hc_sgb_hier_softmax: bad splits");
  if (n_rows <= 0) return HC_OK;
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  int grid = (n_rows + 7) / 8;
  if (grid > 8 * num_sms()) grid = 8 * num_sms();
  sgb_hier_softmax_kernel<<<grid, 256, 0, stream>>>(logits, ld, n_rows, n_geo, n_pos, n_sem, bias_table, num_obj, pair_pred, label_ids, rel,
                                                    super_rel);
  return cuda_status("hc_sgb_hier_softmax");
}

extern "C" int hc_sgb_candidates(const float* rel, int32_t n_geo, int32_t n_pos, int32_t n_sem, const int32_t* pair_offsets,
                                 const int32_t* pair_img, const int32_t* pair_idx, const float* obj_scores, const int32_t* label_ids,
                                 int32_t n_rows, float* cand_score, int32_t* cand_label, int32_t* cand_row, hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(rel && pair_offsets && pair_img && pair_idx && obj_scores && label_ids && cand_score && cand_label && cand_row, HC_E_NULL,
             "hc_sgb_candidates: NULL pointer");
  if (n_rows <= 0) return HC_OK;
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  sgb_candidates_kernel<<<sgrid(n_rows, 128), 128, 0, stream>>>(rel, n_geo, n_pos, n_sem, pair_offsets, pair_img, pair_idx, obj_scores,
                                                                label_ids, n_rows, cand_score, cand_label, cand_row);
  return cuda_status("hc_sgb_candidates");
}

extern "C" int hc_sgb_rank_match(const int32_t* ranked, const uint8_t* reject, const int32_t* pair_offsets, int32_t n_images,
                                 const float* cand_score, const int32_t* cand_label, const int32_t* cand_row, const int32_t* pair_idx,
                                 const int32_t* pred_cls, const float* pred_box, const int32_t* gt_offsets, const int32_t* gt_rel,
                                 const int32_t* gt_cls, const float* gt_box, float iou_thresh, int32_t top_max, int32_t k0, int32_t k1,
                                 int32_t k2, int32_t* final_rank, int32_t* img_hits, int32_t* img_ngt, int32_t* img_hits_pc,
                                 int32_t* img_cnt_pc, hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(ranked && pair_offsets && cand_score && cand_label && cand_row && pair_idx && pred_cls && pred_box && gt_offsets && gt_rel &&
                 gt_cls && gt_box && img_hits && img_ngt && img_hits_pc && img_cnt_pc,
             HC_E_NULL, "hc_sgb_rank_match: NULL pointer");
  HC_REQUIRE(top_max >= 1 && top_max <= SG_MAXR, HC_E_SHAPE, "hc_sgb_rank_match: top_max must be in [1,128]");
  HC_REQUIRE(aligned16(pred_box) && aligned16(gt_box), HC_E_ALIGN, "hc_sgb_rank_match: boxes must be 16-byte aligned");
  if (n_images <= 0) return HC_OK;
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  sgb_rank_match_kernel<<<n_images, SG_THREADS, 0, stream>>>(ranked, reject, pair_offsets, cand_score, cand_label, cand_row, pair_idx, pred_cls,
                                                            reinterpret_cast<const float4*>(pred_box), gt_offsets, gt_rel, gt_cls,
                                                            reinterpret_cast<const float4*>(gt_box), iou_thresh, top_max, k0, k1, k2,
                                                            final_rank, img_hits, img_ngt, img_hits_pc, img_cnt_pc);
  return cuda_status("hc_sgb_rank_match");
}
