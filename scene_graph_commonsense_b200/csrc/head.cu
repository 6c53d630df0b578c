// R6 tail + R7: label-embedding add, fc2 bias + ReLU, fc3_x/fc4/fc5 heads, Bayesian hierarchical log-softmax
// (model.py:152-168,175-184), and R8/R9: candidate construction with the commonsense bitmap filter
// (evaluator.py:157-179,231-266,646-649).  Both are HBM/latency-bound: one warp per directed pair for the
// head (fp32, coalesced 512-byte row reads, shuffle reductions), one thread per pair for the candidates.
#include <math.h>

#include "hc_common.cuh"

namespace hc {

constexpr int HEAD_HV = 4;          // float4 vectors per lane: hidden = 32 lanes * 4 * HEAD_HV = 512
constexpr int HEAD_MAX_OUT = 64;

__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ void add_row(float4 (&x)[HEAD_HV], const float* __restrict__ row, int lane) {
#pragma unroll
  for (int i = 0; i < HEAD_HV; ++i) {
    float4 w = __ldg(reinterpret_cast<const float4*>(row) + lane + 32 * i);
    x[i].x += w.x; x[i].y += w.y; x[i].z += w.z; x[i].w += w.w;
  }
}

// Super-class multi-hot of utils.py:136-149 (`process_super_class`): one_hot(s[0]) plus, for a list of 2..4 entries, one_hot of
// its LAST entry only - `sc[idx] += one_hot(s[i])` runs over the rows with len(s) == i + 1, so the middle entries of a 3- or
// 4-entry list are never added (13 of the 150 VG classes have 3 super-classes).  box_super rows are the raw lists, left-packed,
// -1 padded; these two helpers pick the entries the reference sums.
__device__ __forceinline__ int super_count(const int8_t* __restrict__ s) {
  return (s[0] >= 0) + (s[1] >= 0) + (s[2] >= 0) + (s[3] >= 0);
}
__device__ __forceinline__ bool super_used(int k, int n) { return k < n && (k == 0 || k == n - 1); }

// value j of a warp-distributed vector lives in lane (j & 31), register (j >> 5)
__device__ __forceinline__ float seg_pick(float v0, float v1, int lane, int a, int b, float other) {
  float r = other;
  if (lane >= a && lane < b) r = v0;
  if (lane + 32 >= a && lane + 32 < b) r = v1;
  return r;
}

constexpr int HEAD_JB = 8;          // outputs per accumulation group (HEAD_RB * HEAD_JB == 32 accumulators per lane)
constexpr int HEAD_RB = 4;          // rows (directed pairs) per warp: every head-weight vector read from shared memory is used RB times

constexpr int HEAD_THREADS = 512;   // 16 warps, one CTA per SM: the 110 KB weight tile + 128 registers/thread fill an SM

// Persistent CTAs stage the whole head matrix [n_out, 512] f32 (<= 128 KB) in shared memory once (reading it through L1
// thrashed and made the kernel L2-bound); one warp owns HEAD_RB consecutive rows.  Outputs are processed HEAD_JB at a time:
// every lane keeps RB x JB = 32 independent fp32 accumulators (one conflict-free LDS.128 of weights feeds 16 FMAs), and a
// single packed butterfly (31 shuffles, 5 dependent levels) reduces all 32 across the warp.  Totals are parked in shared
// memory, from where the log-softmax stage reads them in the "value j in lane j&31, register j>>5" layout.
__global__ void __launch_bounds__(HEAD_THREADS, 1)
hier_head_kernel(const float* __restrict__ raw, long long ld_raw, int n_rows, const float* __restrict__ fc2_bias,
                 const float* __restrict__ emb, int num_obj, int num_super, const int* __restrict__ row_sub,
                 const int* __restrict__ row_obj, const int* __restrict__ box_cat, const int8_t* __restrict__ box_super,
                 const float* __restrict__ w_heads, const float* __restrict__ b_heads, int n_geo, int n_pos, int n_sem, int flat,
                 float it1, float it2, float it3, float* __restrict__ relation, float* __restrict__ super_rel,
                 float* __restrict__ connectivity, float* __restrict__ logsig, float* __restrict__ pred_out,
                 const float* __restrict__ box_emb) {
  const int hidden = 128 * HEAD_HV;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  const int R = n_geo + n_pos + n_sem;
  const int n_out = flat ? R + 1 : R + 4;
  extern __shared__ __align__(16) float head_smem[];
  float* w_s = head_smem;                                                             // [n_out][hidden]
  float (*s_out)[HEAD_RB][HEAD_MAX_OUT] =
      reinterpret_cast<float (*)[HEAD_RB][HEAD_MAX_OUT]>(head_smem + (size_t)n_out * hidden);   // [warps][RB][64]
  for (int i = threadIdx.x; i < n_out * (hidden / 4); i += blockDim.x)
    reinterpret_cast<float4*>(w_s)[i] = __ldg(reinterpret_cast<const float4*>(w_heads) + i);
  __syncthreads();
  for (long long row0 = ((long long)blockIdx.x * warps_per_block + wid) * HEAD_RB; row0 < n_rows;
       row0 += (long long)gridDim.x * warps_per_block * HEAD_RB) {
    float4 x[HEAD_RB][HEAD_HV];
#pragma unroll
    for (int r = 0; r < HEAD_RB; ++r) {
      const long long row = min(row0 + r, (long long)n_rows - 1);      // tail rows are recomputed, never stored
#pragma unroll
      for (int i = 0; i < HEAD_HV; ++i) x[r][i] = __ldg(reinterpret_cast<const float4*>(raw + row * ld_raw) + lane + 32 * i);
      if (fc2_bias) {                   // NULL: `raw` already is the 512-d hidden vector (BayesianHead, model.py:24-34)
        add_row(x[r], fc2_bias, lane);
        // one-hot / multi-hot label columns of fc2 (model.py:153-157) as embedding-row adds
        const int bs = row_sub[row], bo = row_obj[row];
        if (box_emb) {                  // per-box sums of the label columns, precomputed by box_label_embed_kernel
          add_row(x[r], box_emb + (long long)bs * 2 * hidden, lane);
          add_row(x[r], box_emb + (long long)bo * 2 * hidden + hidden, lane);
        } else {
          add_row(x[r], emb + (long long)box_cat[bs] * hidden, lane);
          add_row(x[r], emb + (long long)(num_obj + box_cat[bo]) * hidden, lane);
        }
        if (box_super && !box_emb) {
          const int n1 = super_count(box_super + bs * 4), n2 = super_count(box_super + bo * 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            int s1 = box_super[bs * 4 + k], s2 = box_super[bo * 4 + k];
            if (super_used(k, n1)) add_row(x[r], emb + (long long)(2 * num_obj + s1) * hidden, lane);
            if (super_used(k, n2)) add_row(x[r], emb + (long long)(2 * num_obj + num_super + s2) * hidden, lane);
          }
        }
#pragma unroll
        for (int i = 0; i < HEAD_HV; ++i) {
          x[r][i].x = fmaxf(x[r][i].x, 0.f); x[r][i].y = fmaxf(x[r][i].y, 0.f);
          x[r][i].z = fmaxf(x[r][i].z, 0.f); x[r][i].w = fmaxf(x[r][i].w, 0.f);
        }
      }
      if (pred_out && row0 + r < n_rows) {
#pragma unroll
        for (int i = 0; i < HEAD_HV; ++i) reinterpret_cast<float4*>(pred_out + row * hidden)[lane + 32 * i] = x[r][i];
      }
    }
    __syncwarp();
    // 8 outputs x 4 rows = 32 independent accumulators per lane, then ONE packed butterfly reduces all 32 across the warp
    // (31 shuffles, 5 dependent levels): lane L ends up with the total of accumulator L = (row L>>3, output j0 + (L&7)).
    for (int j0 = 0; j0 < n_out; j0 += HEAD_JB) {
      float v[HEAD_RB * HEAD_JB];
#pragma unroll
      for (int t = 0; t < HEAD_RB * HEAD_JB; ++t) v[t] = 0.f;
#pragma unroll
      for (int i = 0; i < HEAD_HV; ++i) {
#pragma unroll
        for (int jj = 0; jj < HEAD_JB; ++jj) {
          const int j = min(j0 + jj, n_out - 1);             // tail outputs recompute the last row, never stored
          const float4 ww = reinterpret_cast<const float4*>(w_s + j * hidden)[lane + 32 * i];
#pragma unroll
          for (int r = 0; r < HEAD_RB; ++r) {                // per accumulator: i ascending, then x,y,z,w
            float a = v[r * HEAD_JB + jj];
            a = fmaf(x[r][i].x, ww.x, a); a = fmaf(x[r][i].y, ww.y, a);
            a = fmaf(x[r][i].z, ww.z, a); a = fmaf(x[r][i].w, ww.w, a);
            v[r * HEAD_JB + jj] = a;
          }
        }
      }
#pragma unroll
      for (int half = 16; half >= 1; half >>= 1) {
        const bool up = lane & half;
#pragma unroll
        for (int t = 0; t < half; ++t) {
          const float keep = up ? v[t + half] : v[t];
          const float send = up ? v[t] : v[t + half];
          v[t] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
      }
      const int jo = j0 + (lane & (HEAD_JB - 1));
      if (jo < n_out) s_out[wid][lane / HEAD_JB][jo] = v[0] + __ldg(b_heads + jo);
    }
    __syncwarp();
#pragma unroll 1
    for (int r = 0; r < HEAD_RB; ++r) {
      const long long row = row0 + r;
      if (row >= n_rows) break;                              // warp-uniform
      const float v0 = lane < n_out ? s_out[wid][r][lane] : 0.f;
      const float v1 = lane + 32 < n_out ? s_out[wid][r][lane + 32] : 0.f;
      // connectivity = fc4 (model.py:176), logsig = log(sigmoid(.)) as composed by train_utils.py:190
      float conn = __shfl_sync(0xffffffffu, (R < 32) ? v0 : v1, R & 31);
      if (lane == 0) {
        connectivity[row] = conn;
        logsig[row] = logf(1.0f / (1.0f + expf(-conn)));
      }
      if (flat) {
        if (lane < R) relation[row * R + lane] = v0;
        if (lane + 32 < R) relation[row * R + lane + 32] = v1;
        continue;
      }
      // super = log_softmax(fc5) over outputs R+1..R+3 (model.py:177)
      const float NEG = -INFINITY;
      float sv = seg_pick(v0, v1, lane, R + 1, R + 4, NEG);
      float sm = warp_max(sv);
      float ss = warp_sum(sv == NEG ? 0.f : expf(sv - sm));
      float slog = sm + logf(ss);
      float sup[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        int j = R + 1 + k;
        sup[k] = __shfl_sync(0xffffffffu, (j < 32) ? v0 : v1, j & 31) - slog;
      }
      if (lane < 3 && super_rel) super_rel[row * 3 + lane] = sup[lane == 0 ? 0 : (lane == 1 ? 1 : 2)];
      // rel_k = log_softmax(fc3_k / T_k) + super[k] (model.py:179-184)
      const int seg_a[3] = {0, n_geo, n_geo + n_pos};
      const int seg_b[3] = {n_geo, n_geo + n_pos, R};
      const float inv_t[3] = {it1, it2, it3};
      float o0 = 0.f, o1 = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float z = seg_pick(v0, v1, lane, seg_a[k], seg_b[k], NEG);
        if (z != NEG) z *= inv_t[k];
        float m = warp_max(z);
        float e = warp_sum(z == NEG ? 0.f : expf(z - m));
        float lse = logf(e);
        if (lane >= seg_a[k] && lane < seg_b[k]) o0 = (v0 * inv_t[k] - m - lse) + sup[k];
        if (lane + 32 >= seg_a[k] && lane + 32 < seg_b[k]) o1 = (v1 * inv_t[k] - m - lse) + sup[k];
      }
      if (lane < R) relation[row * R + lane] = o0;
      if (lane + 32 < R) relation[row * R + lane + 32] = o1;
    }
    __syncwarp();                                            // s_out is reused by the next row group
  }
}

// Label columns of fc2 (model.py:153-157) summed once per BOX instead of once per pair: out[box] = [ S | O ] with
// S = E[cat] + sum_k E[2*num_obj + super_k] (box as subject), O = E[num_obj + cat] + sum_k E[2*num_obj + num_super + super_k]
// (box as object).  A pair then adds two 2 KB rows instead of up to ten (the gathers were 90 % of the head's L2 traffic).
__global__ void box_label_embed_kernel(const float* __restrict__ emb, int num_obj, int num_super, const int* __restrict__ box_cat,
                                       const int8_t* __restrict__ box_super, int n_box, float* __restrict__ out) {
  const int hidden = 128 * HEAD_HV;
  const int lane = threadIdx.x & 31;
  const int box = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (box >= n_box) return;
  const int c = box_cat[box];
#pragma unroll
  for (int role = 0; role < 2; ++role) {
    float4 x[HEAD_HV];
#pragma unroll
    for (int i = 0; i < HEAD_HV; ++i) x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    add_row(x, emb + (long long)(role * num_obj + c) * hidden, lane);
    if (box_super) {
      const int ns = super_count(box_super + box * 4);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int sc = box_super[box * 4 + k];
        if (super_used(k, ns)) add_row(x, emb + (long long)(2 * num_obj + role * num_super + sc) * hidden, lane);
      }
    }
#pragma unroll
    for (int i = 0; i < HEAD_HV; ++i) reinterpret_cast<float4*>(out + ((long long)box * 2 + role) * hidden)[lane + 32 * i] = x[i];
  }
}

__global__ void candidates_kernel(const float* __restrict__ relation, long long ld_rel, int n_rows, int n_geo, int n_pos, int n_sem,
                                  int hier, const uint8_t* __restrict__ row_ov, const float* __restrict__ logsig,
                                  const float* __restrict__ conf_sub, const float* __restrict__ conf_obj,
                                  const int* __restrict__ row_sub, const int* __restrict__ row_obj, const int* __restrict__ box_cat,
                                  const uint32_t* __restrict__ pass_bitmap, const float* __restrict__ super_rel,
                                  float* __restrict__ cand_conf, int* __restrict__ cand_label, float* __restrict__ t3_conf,
                                  uint8_t* __restrict__ t3_super, int layout) {
  const int R = n_geo + n_pos + n_sem;
  const int K = hier ? 3 : 1;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += gridDim.x * blockDim.x) {
    const float* rel = relation + (long long)r * ld_rel;
    const int seg_a[3] = {0, hier ? n_geo : R, n_geo + n_pos};
    const int seg_b[3] = {hier ? n_geo : R, n_geo + n_pos, R};
    const bool ov = row_ov ? row_ov[r] != 0 : true;
    const float ls = logsig[r];
    const int cs = box_cat[row_sub[r]], co = box_cat[row_obj[r]];
    float t3 = -INFINITY;
    for (int k = 0; k < K; ++k) {
      // torch.max / torch.argmax over the segment: first maximum wins (evaluator.py:160-174)
      float best = rel[seg_a[k]];
      int arg = seg_a[k];
      for (int j = seg_a[k] + 1; j < seg_b[k]; ++j) {
        float v = rel[j];
        if (v > best) { best = v; arg = j; }
      }
      t3 = fmaxf(t3, best);
      float c = best;
      if (conf_sub) c += conf_sub[r] + conf_obj[r];           // evaluator.py:164-166
      if (!ov) c = -INFINITY;                                 // :167-168
      if (pass_bitmap) {                                      // :189-194 / :261-266
        bool pass = cs >= 0 && cs < HC_NUM_OBJ && co >= 0 && co < HC_NUM_OBJ && arg < 50 &&
                    bitmap_test(pass_bitmap, (cs * 50 + arg) * HC_NUM_OBJ + co);
        if (!pass) c = -INFINITY;
      }
      c += ls;                                                // evaluator.py:292
      long long idx = layout == 0 ? (long long)r * K + k : (long long)k * n_rows + r;
      cand_conf[idx] = c;
      cand_label[idx] = arg;
    }
    if (t3_conf) {                                            // Evaluator_Top3 (evaluator.py:646-649,702)
      if (!ov) t3 = -INFINITY;
      t3_conf[r] = t3 + ls;
    }
    if (t3_super && super_rel) {
      const float* s = super_rel + (long long)r * 3;
      int a = 0;
      if (s[1] > s[a]) a = 1;
      if (s[2] > s[a]) a = 2;
      t3_super[r] = (uint8_t)a;
    }
  }
}

}  // namespace hc

using namespace hc;

extern "C" int hc_hier_head(const float* fc2_raw, int64_t ld_raw, int32_t n_rows, int32_t hidden, const float* fc2_bias,
                            const float* emb, int32_t num_obj, int32_t num_super, const int32_t* row_sub, const int32_t* row_obj,
                            const int32_t* box_cat, const int8_t* box_super, const float* w_heads, const float* b_heads, int32_t n_geo,
                            int32_t n_pos, int32_t n_sem, int32_t flat, float t1, float t2, float t3, float* relation,
                            float* super_rel, float* connectivity, float* logsig, float* pred_out, const float* box_emb,
                            hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(fc2_raw && w_heads && b_heads && relation && connectivity && logsig, HC_E_NULL, "hc_hier_head: required pointer is NULL");
  HC_REQUIRE(!fc2_bias || (emb && row_sub && row_obj && box_cat), HC_E_NULL,
             "hc_hier_head: fc2_bias given, so emb/row_sub/row_obj/box_cat are required");
  HC_REQUIRE(hidden == 128 * HEAD_HV, HC_E_SHAPE, "hc_hier_head: hidden must be 512");
  HC_REQUIRE(n_geo > 0 && n_pos >= 0 && n_sem >= 0 && n_geo + n_pos + n_sem + 4 <= HEAD_MAX_OUT, HC_E_SHAPE,
             "hc_hier_head: at most 60 predicate classes");
  HC_REQUIRE(flat || super_rel, HC_E_NULL, "hc_hier_head: super_rel required for the hierarchical head");
  HC_REQUIRE(t1 != 0.f && t2 != 0.f && t3 != 0.f, HC_E_SHAPE, "hc_hier_head: temperatures must be non-zero");
  HC_REQUIRE(!box_emb || (fc2_bias && aligned16(box_emb)), HC_E_ALIGN, "hc_hier_head: box_emb needs fc2_bias and 16-byte alignment");
  HC_REQUIRE(ld_raw % 4 == 0 && aligned16(fc2_raw) && (!fc2_bias || (aligned16(fc2_bias) && aligned16(emb))) && aligned16(w_heads) &&
                 (!pred_out || aligned16(pred_out)),
             HC_E_ALIGN, "hc_hier_head: 16-byte alignment");
  if (n_rows <= 0) return HC_OK;
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  const int rows_per_cta = (HEAD_THREADS / 32) * HEAD_RB;
  int grid = (n_rows + rows_per_cta - 1) / rows_per_cta;
  if (grid > num_sms()) grid = num_sms();
  const int n_out = flat ? n_geo + n_pos + n_sem + 1 : n_geo + n_pos + n_sem + 4;
  const size_t smem = ((size_t)n_out * hidden + (size_t)(HEAD_THREADS / 32) * HEAD_RB * HEAD_MAX_OUT) * sizeof(float);
  static bool configured[HC_MAX_DEVICES] = {};
  const int cfg_dev = current_device();
  if (cfg_dev < 0 || cfg_dev >= HC_MAX_DEVICES) return fail(HC_E_CUDA, "device index out of range");
  if (!configured[cfg_dev]) {
    if (cudaFuncSetAttribute(hier_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)((HEAD_MAX_OUT * 128 * HEAD_HV + (HEAD_THREADS / 32) * HEAD_RB * HEAD_MAX_OUT) * sizeof(float))) != cudaSuccess)
      return cuda_status("cudaFuncSetAttribute(hier_head_kernel)");
    configured[cfg_dev] = true;
  }
  hier_head_kernel<<<grid, HEAD_THREADS, smem, stream>>>(fc2_raw, ld_raw, n_rows, fc2_bias, emb, num_obj, num_super, row_sub, row_obj, box_cat,
                                             box_super, w_heads, b_heads, n_geo, n_pos, n_sem, flat, 1.0f / t1, 1.0f / t2, 1.0f / t3,
                                             relation, super_rel, connectivity, logsig, pred_out, box_emb);
  return cuda_status("hc_hier_head");
}

extern "C" int hc_box_label_embed(const float* emb, int32_t num_obj, int32_t num_super, const int32_t* box_cat, const int8_t* box_super,
                                  int32_t n_box, int32_t hidden, float* out, hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(emb && box_cat && out, HC_E_NULL, "hc_box_label_embed: required pointer is NULL");
  HC_REQUIRE(hidden == 128 * HEAD_HV, HC_E_SHAPE, "hc_box_label_embed: hidden must be 512");
  HC_REQUIRE(aligned16(emb) && aligned16(out), HC_E_ALIGN, "hc_box_label_embed: 16-byte alignment");
  if (n_box <= 0) return HC_OK;
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  box_label_embed_kernel<<<(n_box + 7) / 8, 256, 0, stream>>>(emb, num_obj, num_super, box_cat, box_super, n_box, out);
  return cuda_status("hc_box_label_embed");
}

extern "C" int hc_candidates(const float* relation, int64_t ld_rel, int32_t n_rows, int32_t n_geo, int32_t n_pos, int32_t n_sem,
                             int32_t hier, const uint8_t* row_ov, const float* logsig, const float* conf_sub, const float* conf_obj,
                             const int32_t* row_sub, const int32_t* row_obj, const int32_t* box_cat, const uint32_t* pass_bitmap,
                             const float* super_rel, float* cand_conf, int32_t* cand_label, float* t3_conf, uint8_t* t3_super,
                             int32_t layout, hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(relation && logsig && row_sub && row_obj && box_cat && cand_conf && cand_label, HC_E_NULL,
             "hc_candidates: required pointer is NULL");
  HC_REQUIRE((conf_sub == nullptr) == (conf_obj == nullptr), HC_E_NULL, "hc_candidates: conf_sub and conf_obj go together");
  HC_REQUIRE(n_geo > 0 && n_pos >= 0 && n_sem >= 0 && ld_rel >= n_geo + n_pos + n_sem, HC_E_SHAPE, "hc_candidates: bad splits");
  HC_REQUIRE(layout == 0 || layout == 1, HC_E_SHAPE, "hc_candidates: layout must be 0 or 1");
  if (n_rows <= 0) return HC_OK;
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  int grid = (n_rows + 127) / 128;
  if (grid > 8 * num_sms()) grid = 8 * num_sms();
  candidates_kernel<<<grid, 128, 0, stream>>>(relation, ld_rel, n_rows, n_geo, n_pos, n_sem, hier, row_ov, logsig, conf_sub, conf_obj,
                                              row_sub, row_obj, box_cat, pass_bitmap, super_rel, cand_conf, cand_label, t3_conf,
                                              t3_super, layout);
  return cuda_status("hc_candidates");
}
