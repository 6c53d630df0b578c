// N4: the training-side losses on the fused hierarchical head and their backward.
//   hier_loss_kernel      per CALL (graph_iter, edge_iter, direction) of train_test.py:189-258: commonsense penalty
//                         (train_utils.py:36-60), connectivity BCE (:62-90), hierarchical / flat NLL (:116-157), and the gradient
//                         of the step loss with respect to the head's pre-softmax logits
//   loss_total_kernel     step loss with the reference's running-sum weights (train_test.py:219-230)
//   head_bwd_dpred_kernel d_pred = d_logits @ W_heads                 (fc3_x / fc4 / fc5 backward, model.py:171-183)
//   head_bwd_dw_kernel    d_W = d_logits^T @ pred, d_b = sum d_logits (two-stage, fixed summation order)
// All of it is fp32 SIMT work on a few hundred bytes per directed pair (HBM/latency-bound); there are no float atomics, so the
// losses and gradients are bit-reproducible run to run.
#include <math.h>

#include "hc_common.cuh"

namespace hc {

constexpr int TR_MAX_OUT = 64;
constexpr int TR_HIDDEN = 512;

__device__ __forceinline__ float wsum_f(float v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int wsum_i(int v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// softplus(x) = log(1 + exp(x)) the way BCEWithLogitsLoss evaluates it: max(x,0) + log1p(exp(-|x|))
__device__ __forceinline__ float softplus_f(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

struct LossParams {
  float lam_conn, lam_not_connected, lam_cs, lam_cs_weak, lam_cs_strong;
  float inv_t[3];
  int n_geo, n_pos, n_sem, hier;
};

// segment max / first argmax / sum exp(x - max) of rel[a:b)
__device__ __forceinline__ void seg_stats(const float* __restrict__ rel, int a, int b, float& mx, int& arg, float& se) {
  mx = rel[a]; arg = a;
  for (int j = a + 1; j < b; ++j) { float v = rel[j]; if (v > mx) { mx = v; arg = j; } }
  se = 0.f;
  for (int j = a; j < b; ++j) se += expf(rel[j] - mx);
}

// One warp per call.  Lane l owns rows l, l+32, ... of the call (a call holds at most one row per image of the lock-step batch);
// pass 1 reduces the call's sums with fixed-order butterflies, pass 2 writes the per-row gradient.
__global__ void __launch_bounds__(128)
hier_loss_kernel(const float* __restrict__ relation, long long ld_rel, const float* __restrict__ super_rel,
                 const float* __restrict__ connectivity, const int* __restrict__ row_target, const int* __restrict__ group_offsets,
                 const int* __restrict__ group_rows, int n_groups, const float* __restrict__ group_weight,
                 const float* __restrict__ class_weight, const uint32_t* __restrict__ aligned_bm,
                 const uint32_t* __restrict__ violated_bm, const int* __restrict__ row_sub, const int* __restrict__ row_obj,
                 const int* __restrict__ box_cat, LossParams lp, float* __restrict__ group_loss, float* __restrict__ d_logits,
                 int ld_dl) {
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= n_groups) return;
  const int R = lp.n_geo + lp.n_pos + lp.n_sem;
  const int K = lp.hier ? 3 : 1;
  const int seg_a[3] = {0, lp.hier ? lp.n_geo : R, lp.n_geo + lp.n_pos};
  const int seg_b[3] = {lp.hier ? lp.n_geo : R, lp.n_geo + lp.n_pos, R};
  const int r0 = group_offsets[m], r1 = group_offsets[m + 1];
  const bool cs_on = aligned_bm != nullptr;

  int n_c = 0, n_nc = 0, n_ny = 0, n_no = 0, cnt[3] = {0, 0, 0};
  float bce1 = 0.f, bce0 = 0.f, sup_nll = 0.f, wnll[3] = {0.f, 0.f, 0.f}, wsum[3] = {0.f, 0.f, 0.f}, s_ny = 0.f, s_no = 0.f;
  for (int i = r0 + lane; i < r1; i += 32) {
    const int r = group_rows[i];
    const float* rel = relation + (long long)r * ld_rel;
    const int t = row_target[r];
    const float x = connectivity[r];
    if (t != -1) {
      ++n_c;
      bce1 += softplus_f(-x);                                   // BCEWithLogits(x, 1), train_utils.py:88
      const float w = class_weight[t];
      if (lp.hier) {
        const int k = (t >= seg_a[1]) + (t >= seg_a[2]);        // utils.super_relation_processing
        sup_nll -= super_rel[(long long)r * 3 + k];             // train_utils.py:137
        wnll[k] -= w * rel[t];                                  // NLLLoss(weight): sum w_t (-x_t) / sum w_t, :154
        wsum[k] += w; ++cnt[k];
      } else {
        float mx, se; int arg;
        seg_stats(rel, 0, R, mx, arg, se);
        wnll[0] += w * (mx + logf(se) - rel[t]);                // CrossEntropyLoss(weight), :157
        wsum[0] += w; ++cnt[0];
      }
    } else {
      ++n_nc;
      bce0 += softplus_f(x);                                    // BCEWithLogits(x, 0), :67
    }
    if (cs_on) {                                                // train_utils.py:36-58
      const int cs = box_cat[row_sub[r]], co = box_cat[row_obj[r]];
      for (int k = 0; k < K; ++k) {
        float mx, se; int arg;
        seg_stats(rel, seg_a[k], seg_b[k], mx, arg, se);
        const float p = 1.0f / se;                              // max softmax = exp(mx - mx) / sum exp(. - mx)
        const bool in_range = cs >= 0 && cs < HC_NUM_OBJ && co >= 0 && co < HC_NUM_OBJ && arg < 50;
        const int key = (cs * 50 + arg) * HC_NUM_OBJ + co;
        const bool yes = in_range && bitmap_test(aligned_bm, key);
        const bool no = in_range && bitmap_test(violated_bm, key);
        if (!yes) { s_ny += p; ++n_ny; }
        if (no) { s_no += p; ++n_no; }
      }
    }
  }
  n_c = wsum_i(n_c); n_nc = wsum_i(n_nc); n_ny = wsum_i(n_ny); n_no = wsum_i(n_no);
  bce1 = wsum_f(bce1); bce0 = wsum_f(bce0); sup_nll = wsum_f(sup_nll); s_ny = wsum_f(s_ny); s_no = wsum_f(s_no);
#pragma unroll
  for (int k = 0; k < 3; ++k) { cnt[k] = wsum_i(cnt[k]); wnll[k] = wsum_f(wnll[k]); wsum[k] = wsum_f(wsum[k]); }

  float l_conn = 0.f, l_rel = 0.f, l_cs = 0.f;
  if (n_c > 0) {
    l_conn = bce1 / (float)n_c;                                 // overwrites the not-connected term, train_utils.py:88-90
    if (lp.hier) l_rel = sup_nll / (float)n_c;
#pragma unroll
    for (int k = 0; k < 3; ++k) if (cnt[k] > 0) l_rel += wnll[k] / wsum[k];
  } else if (n_nc > 0) {
    l_conn = lp.lam_not_connected * (bce0 / (float)n_nc);       // :67-68
  }
  if (n_ny > 0) l_cs += lp.lam_cs_weak * (s_ny / (float)n_ny);
  if (n_no > 0) l_cs += lp.lam_cs_strong * (s_no / (float)n_no);
  if (lane == 0) { group_loss[m * 3 + 0] = l_rel; group_loss[m * 3 + 1] = l_conn; group_loss[m * 3 + 2] = l_cs; }
  if (!d_logits) return;

  // ---- backward: d(step loss)/d(logits) with step loss = sum_m gw[m] (l_rel + lam_conn l_conn + lam_cs l_cs) ----
  const float gw = group_weight[m];
  const float c_ny = n_ny > 0 ? gw * lp.lam_cs * lp.lam_cs_weak / (float)n_ny : 0.f;
  const float c_no = n_no > 0 ? gw * lp.lam_cs * lp.lam_cs_strong / (float)n_no : 0.f;
  for (int i = r0 + lane; i < r1; i += 32) {
    const int r = group_rows[i];
    const float* rel = relation + (long long)r * ld_rel;
    float* dl = d_logits + (long long)r * ld_dl;
    const int t = row_target[r];
    const float x = connectivity[r];
    float dconn = 0.f;
    if (n_c > 0) { if (t != -1) dconn = gw * lp.lam_conn * (sigmoid_f(x) - 1.0f) / (float)n_c; }
    else if (t == -1) dconn = gw * lp.lam_conn * lp.lam_not_connected * sigmoid_f(x) / (float)n_nc;
    dl[R] = dconn;
    const int kt = t == -1 ? -1 : (lp.hier ? (t >= seg_a[1]) + (t >= seg_a[2]) : 0);
    int cs = 0, co = 0;
    if (cs_on) { cs = box_cat[row_sub[r]]; co = box_cat[row_obj[r]]; }
    float gsup[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < K; ++k) {
      float mx, se; int arg;
      seg_stats(rel, seg_a[k], seg_b[k], mx, arg, se);
      const float inv_se = 1.0f / se;
      // a: coefficient of (delta_jt - P_j) from the NLL / CE term; c: coefficient of (delta_jj* - P_j) from the commonsense term
      float a = 0.f, c = 0.f;
      if (kt == k) a = -gw * class_weight[t] / wsum[k];
      if (cs_on) {
        const bool in_range = cs >= 0 && cs < HC_NUM_OBJ && co >= 0 && co < HC_NUM_OBJ && arg < 50;
        const int key = (cs * 50 + arg) * HC_NUM_OBJ + co;
        const bool yes = in_range && bitmap_test(aligned_bm, key);
        const bool no = in_range && bitmap_test(violated_bm, key);
        c = ((yes ? 0.f : c_ny) + (no ? c_no : 0.f)) * inv_se;  // times p = max softmax
      }
      const float it = lp.hier ? lp.inv_t[k] : 1.0f;
      const float ac = a + c;
      for (int j = seg_a[k]; j < seg_b[k]; ++j) {
        const float P = expf(rel[j] - mx) * inv_se;             // softmax over the segment (== softmax(fc3_k / T_k))
        float g = -ac * P;
        if (j == t) g += a;
        if (j == arg) g += c;
        dl[j] = g * it;
      }
      if (lp.hier) gsup[k] = a + (kt == k ? -gw / (float)n_c : 0.f);   // d/d super[k]: joint log-prob adds super[k] to every rel_k
    }
    if (lp.hier) {
      const float gs = gsup[0] + gsup[1] + gsup[2];
#pragma unroll
      for (int k = 0; k < 3; ++k) dl[R + 1 + k] = gsup[k] - expf(super_rel[(long long)r * 3 + k]) * gs;
    }
  }
}

// total[0] = sum_m gw[m] (rel + lam_conn conn + lam_cs cs); total[1..3] = sum_m gw[m] * {rel, conn, cs} (the reference's
// running_loss_* bookkeeping before the lambdas); double accumulation, fixed order.
__global__ void loss_total_kernel(const float* __restrict__ group_loss, const float* __restrict__ group_weight, int n_groups,
                                  float lam_conn, float lam_cs, float* __restrict__ total) {
  __shared__ double sh[3][256];
  double a[3] = {0.0, 0.0, 0.0};
  for (int m = threadIdx.x; m < n_groups; m += blockDim.x) {
    const double w = group_weight[m];
#pragma unroll
    for (int k = 0; k < 3; ++k) a[k] += w * (double)group_loss[m * 3 + k];
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) sh[k][threadIdx.x] = a[k];
  __syncthreads();
  for (int s = 128; s; s >>= 1) {
    if ((int)threadIdx.x < s)
#pragma unroll
      for (int k = 0; k < 3; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    total[0] = (float)(sh[0][0] + (double)lam_conn * sh[1][0] + (double)lam_cs * sh[2][0]);
    total[1] = (float)sh[0][0]; total[2] = (float)sh[1][0]; total[3] = (float)sh[2][0];
  }
}

// d_pred[r, :] = scale * sum_j d_logits[r, j] W[j, :].  Persistent CTAs keep W (n_out x 512 f32 <= 128 KB) in shared memory;
// a warp owns BWD_RB rows at a time and a lane 16 of the 512 columns (4 float4), so one LDS.128 feeds 4 * BWD_RB FMAs (with fewer
// rows per warp the kernel is bound by shared-memory bandwidth: an LDS.128 occupies the 128 B/clk port for 4 cycles).
constexpr int BWD_THREADS = 512;
constexpr int BWD_RB = 4;
__global__ void __launch_bounds__(BWD_THREADS, 1)
head_bwd_dpred_kernel(const float* __restrict__ d_logits, int ld_dl, int n_rows, int n_out, const float* __restrict__ w_heads,
                      const float* __restrict__ scale, float* __restrict__ d_pred) {
  extern __shared__ __align__(16) float bwd_smem[];
  float* w_s = bwd_smem;
  for (int i = threadIdx.x; i < n_out * (TR_HIDDEN / 4); i += blockDim.x)
    reinterpret_cast<float4*>(w_s)[i] = __ldg(reinterpret_cast<const float4*>(w_heads) + i);
  __syncthreads();
  const float sc = scale ? __ldg(scale) : 1.0f;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (long long row0 = ((long long)blockIdx.x * nw + wid) * BWD_RB; row0 < n_rows; row0 += (long long)gridDim.x * nw * BWD_RB) {
    float c0[BWD_RB], c1[BWD_RB];
    float4 x[BWD_RB][4];
#pragma unroll
    for (int r = 0; r < BWD_RB; ++r) {
      const long long row = min(row0 + r, (long long)n_rows - 1);       // tail rows are recomputed, never stored
      c0[r] = lane < n_out ? d_logits[row * ld_dl + lane] : 0.f;
      c1[r] = lane + 32 < n_out ? d_logits[row * ld_dl + lane + 32] : 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) x[r][i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int j = 0; j < n_out; ++j) {
      float c[BWD_RB];
#pragma unroll
      for (int r = 0; r < BWD_RB; ++r) c[r] = __shfl_sync(0xffffffffu, j < 32 ? c0[r] : c1[r], j & 31);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 w = reinterpret_cast<const float4*>(w_s + j * TR_HIDDEN)[lane + 32 * i];
#pragma unroll
        for (int r = 0; r < BWD_RB; ++r) {
          x[r][i].x = fmaf(c[r], w.x, x[r][i].x); x[r][i].y = fmaf(c[r], w.y, x[r][i].y);
          x[r][i].z = fmaf(c[r], w.z, x[r][i].z); x[r][i].w = fmaf(c[r], w.w, x[r][i].w);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < BWD_RB; ++r) {
      if (row0 + r >= n_rows) break;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        reinterpret_cast<float4*>(d_pred + (row0 + r) * TR_HIDDEN)[lane + 32 * i] =
            make_float4(x[r][i].x * sc, x[r][i].y * sc, x[r][i].z * sc, x[r][i].w * sc);
    }
  }
}

// stage 1: CTA b sums its contiguous slab of rows: thread c owns column c of pred and n_out accumulators; d_logits rows are
// staged through shared memory 32 at a time.  ws[b][j][c] (+ bias partials ws_b[b][j]).
constexpr int DW_ROWS = 32;
__global__ void __launch_bounds__(TR_HIDDEN, 1)
head_bwd_dw_kernel(const float* __restrict__ d_logits, int ld_dl, const float* __restrict__ pred, long long ld_pred, int n_rows,
                   int n_out, int rows_per_part, float* __restrict__ ws_w, float* __restrict__ ws_b) {
  __shared__ __align__(16) float dl_s[DW_ROWS][TR_MAX_OUT];
  const int c = threadIdx.x;
  float acc[TR_MAX_OUT];
#pragma unroll
  for (int j = 0; j < TR_MAX_OUT; ++j) acc[j] = 0.f;
  float accb = 0.f;
  const long long lo = (long long)blockIdx.x * rows_per_part;
  const long long hi = min(lo + rows_per_part, (long long)n_rows);
  for (long long base = lo; base < hi; base += DW_ROWS) {
    const int nr = (int)min((long long)DW_ROWS, hi - base);
    __syncthreads();
    for (int i = threadIdx.x; i < DW_ROWS * TR_MAX_OUT; i += blockDim.x) {
      const int rr = i / TR_MAX_OUT, j = i % TR_MAX_OUT;
      dl_s[rr][j] = (rr < nr && j < n_out) ? d_logits[(base + rr) * ld_dl + j] : 0.f;
    }
    __syncthreads();
    for (int r4 = 0; r4 < nr; r4 += 4) {                          // 4 global loads in flight before their 256 FMAs
      float p[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) p[u] = r4 + u < nr ? pred[(base + r4 + u) * ld_pred + c] : 0.f;   // rows >= nr: dl_s holds zeros
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int j = 0; j < TR_MAX_OUT; j += 4) {               // broadcast LDS.128: 4 coefficients per shared-memory read
          const float4 d = *reinterpret_cast<const float4*>(&dl_s[r4 + u][j]);
          acc[j] = fmaf(d.x, p[u], acc[j]); acc[j + 1] = fmaf(d.y, p[u], acc[j + 1]);
          acc[j + 2] = fmaf(d.z, p[u], acc[j + 2]); acc[j + 3] = fmaf(d.w, p[u], acc[j + 3]);
        }
        if (c < TR_MAX_OUT) accb += dl_s[r4 + u][c];
      }
    }
  }
  float* out = ws_w + (long long)blockIdx.x * n_out * TR_HIDDEN;
#pragma unroll
  for (int j = 0; j < TR_MAX_OUT; ++j) if (j < n_out) out[j * TR_HIDDEN + c] = acc[j];
  if (c < n_out) ws_b[blockIdx.x * n_out + c] = accb;
}

// stage 2: fixed-order sum over the parts, times the upstream scale
__global__ void head_bwd_dw_reduce_kernel(const float* __restrict__ ws_w, const float* __restrict__ ws_b, int parts, int n_out,
                                          const float* __restrict__ scale, float* __restrict__ d_w, float* __restrict__ d_b) {
  const int n_w = n_out * TR_HIDDEN;
  const float sc = scale ? __ldg(scale) : 1.0f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_w + n_out; i += gridDim.x * blockDim.x) {
    float s = 0.f;
    if (i < n_w) { for (int b = 0; b < parts; ++b) s += ws_w[(long long)b * n_w + i]; d_w[i] = s * sc; }
    else { const int j = i - n_w; for (int b = 0; b < parts; ++b) s += ws_b[b * n_out + j]; d_b[j] = s * sc; }
  }
}

}  // namespace hc

using namespace hc;

extern "C" int hc_hier_loss(const float* relation, int64_t ld_rel, const float* super_rel, const float* connectivity, int32_t n_rows,
                            int32_t n_geo, int32_t n_pos, int32_t n_sem, int32_t hier, float t1, float t2, float t3,
                            const int32_t* row_target, const int32_t* group_offsets, const int32_t* group_rows, int32_t n_groups,
                            const float* group_weight, const float* class_weight, const uint32_t* aligned_bitmap,
                            const uint32_t* violated_bitmap, const int32_t* row_sub, const int32_t* row_obj, const int32_t* box_cat,
                            float lam_conn, float lam_not_connected, float lam_cs, float lam_cs_weak, float lam_cs_strong,
                            float* group_loss, float* total, float* d_logits, int32_t ld_dl, hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(relation && connectivity && row_target && group_offsets && group_rows && group_weight && class_weight && group_loss &&
                 total, HC_E_NULL, "hc_hier_loss: required pointer is NULL");
  HC_REQUIRE(!hier || super_rel, HC_E_NULL, "hc_hier_loss: super_rel required for the hierarchical head");
  HC_REQUIRE((aligned_bitmap == nullptr) == (violated_bitmap == nullptr), HC_E_NULL,
             "hc_hier_loss: aligned and violated bitmaps go together");
  HC_REQUIRE(!aligned_bitmap || (row_sub && row_obj && box_cat), HC_E_NULL, "hc_hier_loss: the commonsense term needs row_sub/row_obj/box_cat");
  const int R = n_geo + n_pos + n_sem;
  HC_REQUIRE(n_geo > 0 && n_pos >= 0 && n_sem >= 0 && R + 4 <= TR_MAX_OUT && ld_rel >= R, HC_E_SHAPE, "hc_hier_loss: bad splits");
  HC_REQUIRE(!d_logits || ld_dl >= R + (hier ? 4 : 1), HC_E_SHAPE, "hc_hier_loss: ld_dl too small");
  HC_REQUIRE(t1 != 0.f && t2 != 0.f && t3 != 0.f, HC_E_SHAPE, "hc_hier_loss: temperatures must be non-zero");
  HC_REQUIRE(n_rows >= 0 && n_groups > 0, HC_E_SHAPE, "hc_hier_loss: n_groups must be positive");
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  LossParams lp;
  lp.lam_conn = lam_conn; lp.lam_not_connected = lam_not_connected; lp.lam_cs = lam_cs; lp.lam_cs_weak = lam_cs_weak;
  lp.lam_cs_strong = lam_cs_strong; lp.inv_t[0] = 1.0f / t1; lp.inv_t[1] = 1.0f / t2; lp.inv_t[2] = 1.0f / t3;
  lp.n_geo = n_geo; lp.n_pos = n_pos; lp.n_sem = n_sem; lp.hier = hier;
  // rows that belong to no call keep a zero gradient
  if (d_logits && n_rows > 0) cudaMemsetAsync(d_logits, 0, (size_t)n_rows * ld_dl * sizeof(float), stream);
  hier_loss_kernel<<<(n_groups + 3) / 4, 128, 0, stream>>>(relation, ld_rel, super_rel, connectivity, row_target, group_offsets,
                                                           group_rows, n_groups, group_weight, class_weight, aligned_bitmap,
                                                           violated_bitmap, row_sub, row_obj, box_cat, lp, group_loss, d_logits, ld_dl);
  loss_total_kernel<<<1, 256, 0, stream>>>(group_loss, group_weight, n_groups, lam_conn, lam_cs, total);
  return cuda_status("hc_hier_loss");
}

extern "C" int hc_hier_head_bwd(const float* d_logits, int32_t ld_dl, const float* pred, int64_t ld_pred, int32_t n_rows,
                                int32_t n_out, const float* w_heads, const float* scale, float* d_pred, float* d_w, float* d_b,
                                float* ws, int32_t parts, hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(d_logits && pred && w_heads, HC_E_NULL, "hc_hier_head_bwd: required pointer is NULL");
  HC_REQUIRE(n_out > 0 && n_out <= TR_MAX_OUT && ld_dl >= n_out && ld_pred >= TR_HIDDEN, HC_E_SHAPE, "hc_hier_head_bwd: bad shapes");
  HC_REQUIRE((d_w == nullptr) == (d_b == nullptr), HC_E_NULL, "hc_hier_head_bwd: d_w and d_b go together");
  HC_REQUIRE(!d_w || (ws && parts > 0 && parts <= 1024), HC_E_NULL, "hc_hier_head_bwd: weight gradients need a workspace and 1..1024 parts");
  HC_REQUIRE(aligned16(w_heads) && (!d_pred || aligned16(d_pred)), HC_E_ALIGN, "hc_hier_head_bwd: 16-byte alignment");
  if (n_rows <= 0) {
    if (d_w) { cudaMemsetAsync(d_w, 0, (size_t)n_out * TR_HIDDEN * sizeof(float), stream); cudaMemsetAsync(d_b, 0, n_out * sizeof(float), stream); }
    return cuda_status("hc_hier_head_bwd");
  }
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  if (d_pred) {
    static bool configured[HC_MAX_DEVICES] = {};
  const int cfg_dev = current_device();
  if (cfg_dev < 0 || cfg_dev >= HC_MAX_DEVICES) return fail(HC_E_CUDA, "device index out of range");
    if (!configured[cfg_dev]) {
      if (cudaFuncSetAttribute(head_bwd_dpred_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)(TR_MAX_OUT * TR_HIDDEN * sizeof(float))) != cudaSuccess)
        return cuda_status("cudaFuncSetAttribute(head_bwd_dpred_kernel)");
      configured[cfg_dev] = true;
    }
    const int rows_per_cta = (BWD_THREADS / 32) * BWD_RB;
    int grid = (n_rows + rows_per_cta - 1) / rows_per_cta;
    if (grid > num_sms()) grid = num_sms();
    head_bwd_dpred_kernel<<<grid, BWD_THREADS, (size_t)n_out * TR_HIDDEN * sizeof(float), stream>>>(d_logits, ld_dl, n_rows, n_out, w_heads,
                                                                                               scale, d_pred);
  }
  if (d_w) {
    const int rows_per_part = (n_rows + parts - 1) / parts;
    float* ws_b = ws + (size_t)parts * n_out * TR_HIDDEN;
    head_bwd_dw_kernel<<<parts, TR_HIDDEN, 0, stream>>>(d_logits, ld_dl, pred, ld_pred, n_rows, n_out, rows_per_part, ws, ws_b);
    head_bwd_dw_reduce_kernel<<<(n_out * TR_HIDDEN + n_out + 255) / 256, 256, 0, stream>>>(ws, ws_b, parts, n_out, scale, d_w, d_b);
  }
  return cuda_status("hc_hier_head_bwd");
}
