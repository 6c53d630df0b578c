// R1 + R2 + R4: rectangular box masks, two-pass ordered pair enumeration and the overlap pre-filter
// (evaluate.py:111-116,132-156; train_test.py:365-410), as four small integer kernels over CSR image segments.
// The 32x32 masks of the reference are never rasterised: two axis-aligned rectangles overlap on the grid iff
// their clamped integer extents intersect (hc_common.cuh: rect_of / rect_inter).
#include "hc_common.cuh"

namespace hc {

__device__ __forceinline__ void tri_decode(int t, int& g, int& e) {
  // t = g(g-1)/2 + e, 0 <= e < g  (reference loop order: for g: for e < g)
  g = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)t)) * 0.5f);
  while (g * (g - 1) / 2 > t) --g;
  while ((g + 1) * g / 2 <= t) ++g;
  e = t - g * (g - 1) / 2;
}

__global__ void pairs_flags_kernel(const int4* __restrict__ boxes, const int* __restrict__ box_off,
                                   const int* __restrict__ tri_off, const int* __restrict__ group_id, int max_tri, int fs,
                                   uint8_t* __restrict__ ov, uint8_t* __restrict__ any) {
  const int img = blockIdx.x;
  const int b0 = box_off[img], n = box_off[img + 1] - b0;
  const int T = n * (n - 1) / 2;
  const int t0 = tri_off[img];
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    int g, e;
    tri_decode(t, g, e);
    Rect a = rect_of(boxes[b0 + g], fs), b = rect_of(boxes[b0 + e], fs);
    uint8_t o = rect_inter(a, b) > 0;
    ov[t0 + t] = o;
    if (group_id && o) any[(size_t)group_id[img] * max_tri + t] = 1;   // every writer stores 1: race-free
  }
}

__device__ __forceinline__ bool survives(const uint8_t* ov, const uint8_t* any, const int* group_id, int img, int t0, int t,
                                         int max_tri) {
  return group_id ? any[(size_t)group_id[img] * max_tri + t] != 0 : ov[t0 + t] != 0;
}

__global__ void pairs_count_kernel(const int* __restrict__ box_off, const int* __restrict__ tri_off,
                                   const int* __restrict__ group_id, int max_tri, const uint8_t* __restrict__ ov,
                                   const uint8_t* __restrict__ any, int* __restrict__ counts) {
  const int img = blockIdx.x;
  const int n = box_off[img + 1] - box_off[img];
  const int T = n * (n - 1) / 2, t0 = tri_off[img];
  int c = 0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) c += survives(ov, any, group_id, img, t0, t, max_tri);
  __shared__ int warp_sum[32];
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) s += warp_sum[w];
    counts[img] = s;
  }
}

// single-block exclusive scan of 2*counts -> directed-pair CSR offsets
__global__ void pairs_scan_kernel(const int* __restrict__ counts, int n_images, int* __restrict__ pair_off, int* __restrict__ total) {
  __shared__ int warp_sum[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_images; base += blockDim.x) {
    int i = base + threadIdx.x;
    int v = i < n_images ? 2 * counts[i] : 0;
    int incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += y;
    }
    if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += warp_sum[w];
    int c = carry;
    if (i < n_images) pair_off[i] = c + woff + incl - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = c + woff + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) { pair_off[n_images] = carry; *total = carry; }
}

__global__ void pairs_fill_kernel(const int* __restrict__ box_off, const int* __restrict__ tri_off,
                                  const int* __restrict__ group_id, int max_tri, const uint8_t* __restrict__ ov,
                                  const uint8_t* __restrict__ any, const int* __restrict__ rel_tri,
                                  const int8_t* __restrict__ dir_tri, const int* __restrict__ pair_off, int* __restrict__ pair_sub,
                                  int* __restrict__ pair_obj, int* __restrict__ pair_img, uint8_t* __restrict__ pair_ov,
                                  int* __restrict__ pair_gt, int* __restrict__ pair_rel) {
  const int img = blockIdx.x;
  const int b0 = box_off[img], n = box_off[img + 1] - b0;
  const int T = n * (n - 1) / 2, t0 = tri_off[img];
  const int out0 = pair_off[img];
  __shared__ int warp_cnt[32];
  __shared__ int base_rank;
  if (threadIdx.x == 0) base_rank = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) / 32;
  for (int start = 0; start < T; start += blockDim.x) {     // ordered compaction: pair_rank follows loop order t
    int t = start + threadIdx.x;
    bool keep = t < T && survives(ov, any, group_id, img, t0, t, max_tri);
    unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[wid] = __popc(m);
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < wid; ++w) woff += warp_cnt[w];
    int rank = base_rank + woff + __popc(m & ((1u << lane) - 1u));
    if (keep) {
      int g, e;
      tri_decode(t, g, e);
      int rel = rel_tri ? rel_tri[t0 + t] : -1;
      int d = dir_tri ? (int)dir_tri[t0 + t] : -1;
      uint8_t o = ov[t0 + t];
      int p = out0 + 2 * rank;
      pair_sub[p] = b0 + g; pair_obj[p] = b0 + e; pair_img[p] = img; pair_ov[p] = o;
      pair_gt[p] = (d == 1) ? rel : -1;
      pair_sub[p + 1] = b0 + e; pair_obj[p + 1] = b0 + g; pair_img[p + 1] = img; pair_ov[p + 1] = o;
      pair_gt[p + 1] = (d == 0) ? rel : -1;
      if (pair_rel) { pair_rel[p] = rel; pair_rel[p + 1] = rel; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int s = 0;
      for (int w = 0; w < nw; ++w) s += warp_cnt[w];
      base_rank += s;
    }
    __syncthreads();
  }
}

__global__ void conn_stats_kernel(const float* __restrict__ conn, const int* __restrict__ gt_dir, const int* __restrict__ gt_undir,
                                  int n, unsigned long long* __restrict__ stats) {
  unsigned long long c[5] = {0, 0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    // train_utils.py:169-183.  "connected" in a direction <=> the directed GT label survived (!= -1); rows whose
    // undirected label is -1 are never connected.  sigmoid(x) >= 0.5 and round(sigmoid(x)) == 1 are evaluated with
    // the same fp32 sigmoid torch uses (1/(1+exp(-x))): x >= 0 <=> sigmoid >= 0.5; round-half-even(0.5) == 0.
    float x = conn[i];
    float sg = 1.0f / (1.0f + expf(-x));
    bool connected = gt_dir[i] != -1;
    bool pred = sg >= 0.5f;
    c[0] += !connected; c[1] += connected; c[2] += pred;
    c[3] += pred && gt_undir[i] != -1;
    c[4] += connected && rintf(sg) == 1.0f;
  }
  for (int k = 0; k < 5; ++k) {
    unsigned long long v = c[k];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(stats + k, v);
  }
}

}  // namespace hc

using namespace hc;

extern "C" int hc_pairs_enumerate(const int32_t* boxes, const int32_t* box_offsets, int32_t n_images, const int32_t* group_id,
                                  int32_t n_groups, int32_t max_tri, const int32_t* rel_tri, const int8_t* dir_tri,
                                  const int32_t* tri_offsets, int32_t feature_size, uint8_t* ws_ov, uint8_t* ws_any,
                                  int32_t* ws_counts, int32_t* pair_offsets, int32_t* pair_sub, int32_t* pair_obj, int32_t* pair_img,
                                  uint8_t* pair_ov, int32_t* pair_gt, int32_t* pair_rel, int32_t* total_out, hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(boxes && box_offsets && tri_offsets && ws_ov && ws_counts && pair_offsets && pair_sub && pair_obj && pair_img &&
                 pair_ov && pair_gt && total_out,
             HC_E_NULL, "hc_pairs_enumerate: required pointer is NULL");
  HC_REQUIRE(n_images > 0 && feature_size > 0, HC_E_SHAPE, "hc_pairs_enumerate: n_images and feature_size must be positive");
  HC_REQUIRE(!group_id || (ws_any && n_groups > 0 && max_tri >= 0), HC_E_NULL, "hc_pairs_enumerate: batch mode needs ws_any/n_groups");
  HC_REQUIRE(aligned16(boxes), HC_E_ALIGN, "hc_pairs_enumerate: boxes must be 16-byte aligned");
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  if (group_id && (size_t)n_groups * max_tri > 0) cudaMemsetAsync(ws_any, 0, (size_t)n_groups * max_tri, stream);
  pairs_flags_kernel<<<n_images, 256, 0, stream>>>(reinterpret_cast<const int4*>(boxes), box_offsets, tri_offsets, group_id, max_tri,
                                                   feature_size, ws_ov, ws_any);
  pairs_count_kernel<<<n_images, 256, 0, stream>>>(box_offsets, tri_offsets, group_id, max_tri, ws_ov, ws_any, ws_counts);
  pairs_scan_kernel<<<1, 1024, 0, stream>>>(ws_counts, n_images, pair_offsets, total_out);
  pairs_fill_kernel<<<n_images, 256, 0, stream>>>(box_offsets, tri_offsets, group_id, max_tri, ws_ov, ws_any, rel_tri, dir_tri,
                                                  pair_offsets, pair_sub, pair_obj, pair_img, pair_ov, pair_gt, pair_rel);
  return cuda_status("hc_pairs_enumerate");
}

extern "C" int hc_connectivity_stats(const float* connectivity, const int32_t* row_gt_directed, const int32_t* row_gt_undirected,
                                     int32_t n_rows, unsigned long long* stats, hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(connectivity && row_gt_directed && row_gt_undirected && stats, HC_E_NULL, "hc_connectivity_stats: NULL pointer");
  if (n_rows <= 0) return HC_OK;
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  int grid = (n_rows + 255) / 256;
  if (grid > 4 * num_sms()) grid = 4 * num_sms();
  conn_stats_kernel<<<grid, 256, 0, stream>>>(connectivity, row_gt_directed, row_gt_undirected, n_rows, stats);
  return cuda_status("hc_connectivity_stats");
}
