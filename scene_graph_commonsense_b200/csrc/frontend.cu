// N2 - SGDET / SGCLS proposal front-end feeding the relation path:
//   evaluate.py:311-370 (eval_sgd) == :545-591 (eval_sgc): DETR logits/boxes -> top-k labels per object query, label remap,
//        cxcywh -> (x1,x2,y1,y2) on the grid, "no object" masking, per-class NMS (torchvision.ops.nms semantics, fp32),
//        super-category lookup;
//   utils.py:376-422 match_object_categories (SGCLS): label every GT box from its two best proposals by grid IoU;
//   utils.py:294-352 match_target_sgd: flat GT triplet lists in (g,e) loop order.
// The reference runs these as Python loops with an `.item()` / `int()` device sync per element; here each is a handful of
// small integer/fp32 kernels over CSR image segments (one CTA per image, one warp per query / GT box).  All of it is
// latency-sized work (a few hundred KB per window): the design goal is "no host round trips", not bandwidth.
#include "hc_common.cuh"

namespace hc {

constexpr int FE_MAX_ENTRIES = 1024;   // n_queries * topk_cat per image
constexpr int FE_THREADS = 256;

// ---------------------------------------------------------------------------------------------------------------
// exclusive scan of per-image counts -> CSR offsets (single CTA)
__global__ void fe_scan_kernel(const int* __restrict__ counts, int n, int* __restrict__ off, int* __restrict__ total) {
  __shared__ int warp_sum[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += blockDim.x) {
    int i = base + threadIdx.x;
    int v = i < n ? counts[i] : 0;
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
    if (i < n) off[i] = c + woff + incl - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = c + woff + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    off[n] = carry;
    if (total) *total = carry;
  }
}

// ordered block-wide exclusive prefix sum of `v` (all threads must call); `carry` (shared) accumulates across calls
__device__ __forceinline__ int block_excl_scan(int v, int* warp_sum, int& carry) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) / 32;
  int incl = v;
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) warp_sum[wid] = incl;
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < wid; ++w) woff += warp_sum[w];
  int res = carry + woff + incl - v;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < nw; ++w) t += warp_sum[w];
    carry += t;
  }
  __syncthreads();
  return res;
}

// ---------------------------------------------------------------------------------------------------------------
// evaluate.py:311-334: one warp per object query.  softmax over num_classes+1 logits, has_object = argmax < num_classes,
// top-k (value desc, index asc) labels/probabilities, label remap, box conversion.  Entry e = query*topk + r.
__global__ void detr_expand_kernel(const float* __restrict__ logits, const float* __restrict__ boxes, int n_rows, int n_cls1,
                                   int topk, const int* __restrict__ label_map, float fs, int* __restrict__ ent_label,
                                   float* __restrict__ ent_conf, float4* __restrict__ ent_box, uint8_t* __restrict__ ent_valid) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const float* x = logits + (size_t)row * n_cls1;
  constexpr int PER = 8;                              // up to 256 classes
  float v[PER];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    int c = lane + 32 * i;
    v[i] = c < n_cls1 ? x[c] : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    int c = lane + 32 * i;
    v[i] = c < n_cls1 ? expf(v[i] - mx) : -1.0f;      // probabilities are > 0; -1 marks "not a class" / "already taken"
    if (c < n_cls1) sum += v[i];
  }
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  // box: (cx,cy,w,h) -> clamp((cx-w/2, cx+w/2, cy-h/2, cy+h/2), 0, 1) * fs, all in fp32 exactly as evaluate.py:327-332
  float4 b = reinterpret_cast<const float4*>(boxes)[row];
  float hw = __fdiv_rn(b.z, 2.0f), hh = __fdiv_rn(b.w, 2.0f);
  float4 g;
  g.x = __fmul_rn(fminf(fmaxf(__fsub_rn(b.x, hw), 0.f), 1.f), fs);
  g.y = __fmul_rn(fminf(fmaxf(__fadd_rn(b.x, hw), 0.f), 1.f), fs);
  g.z = __fmul_rn(fminf(fmaxf(__fsub_rn(b.y, hh), 0.f), 1.f), fs);
  g.w = __fmul_rn(fminf(fmaxf(__fadd_rn(b.y, hh), 0.f), 1.f), fs);
  bool has_object = false;
  for (int r = 0; r < topk; ++r) {
    float best = -1.0f;
    int bi = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < PER; ++i) {                   // ascending class index per lane: strict > keeps the first maximum
      if (v[i] > best) { best = v[i]; bi = lane + 32 * i; }
    }
    for (int o = 16; o; o >>= 1) {
      float ob = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (r == 0) has_object = bi < n_cls1 - 1;
#pragma unroll
    for (int i = 0; i < PER; ++i)                     // mark taken (unrolled select keeps v[] in registers)
      if (lane + 32 * i == bi) v[i] = -1.0f;
    if (lane == 0) {
      int lab = label_map[bi];
      size_t e = (size_t)row * topk + r;
      ent_label[e] = lab;
      ent_conf[e] = __fdiv_rn(best, sum);
      ent_box[e] = g;
      ent_valid[e] = has_object && lab != n_cls1 - 1;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// evaluate.py:341-365: drop label == "no object", per-class NMS, output order = (class asc, confidence desc, entry asc).
// One CTA per image.  Sort key = class | ~orderable(conf) | entry ; NMS runs one warp per class segment.
__device__ __forceinline__ unsigned orderable(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(FE_THREADS) proposals_nms_kernel(
    const int* __restrict__ ent_label, const float* __restrict__ ent_conf, const float4* __restrict__ ent_box,
    const uint8_t* __restrict__ ent_valid, int n_ent, int n_pow2, double thresh, int* __restrict__ st_label,
    float* __restrict__ st_conf, float4* __restrict__ st_box, int* __restrict__ st_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* key = reinterpret_cast<unsigned long long*>(smem_raw);                 // [n_pow2]
  float4* box = reinterpret_cast<float4*>(key + n_pow2);                                       // [n_pow2] sorted (x1,x2,y1,y2)
  float* area = reinterpret_cast<float*>(box + n_pow2);                                        // [n_pow2]
  int* cls = reinterpret_cast<int*>(area + n_pow2);                                            // [n_pow2]
  int* seg_start = cls + n_pow2;                                                               // [n_pow2]
  uint8_t* supp = reinterpret_cast<uint8_t*>(seg_start + n_pow2);                              // [n_pow2]
  __shared__ int n_seg, n_valid_s, base_rank, warp_cnt[32];
  const int img = blockIdx.x;
  const size_t e0 = (size_t)img * n_ent;
  if (threadIdx.x == 0) { n_seg = 0; n_valid_s = 0; base_rank = 0; }
  __syncthreads();
  int local_valid = 0;
  for (int j = threadIdx.x; j < n_pow2; j += blockDim.x) {
    unsigned long long k = ~0ull;                                                              // invalid entries sort last
    if (j < n_ent && ent_valid[e0 + j]) {
      k = ((unsigned long long)(unsigned)ent_label[e0 + j] << 48) | ((unsigned long long)(~orderable(ent_conf[e0 + j])) << 16) |
          (unsigned long long)j;
      ++local_valid;
    }
    key[j] = k;
  }
  if (local_valid) atomicAdd(&n_valid_s, local_valid);
  __syncthreads();
  for (int size = 2; size <= n_pow2; size <<= 1) {                                             // bitonic sort, ascending
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (n_pow2 >> 1); t += blockDim.x) {
        int lo = 2 * t - (t & (stride - 1));
        int hi = lo + stride;
        bool up = (lo & size) == 0;
        unsigned long long a = key[lo], b = key[hi];
        if ((a > b) == up) { key[lo] = b; key[hi] = a; }
      }
      __syncthreads();
    }
  }
  const int nv = n_valid_s;
  for (int j = threadIdx.x; j < nv; j += blockDim.x) {
    int src = (int)(key[j] & 0xffffu);
    float4 b = ent_box[e0 + src];
    box[j] = b;
    // torchvision nms_kernel.cpp: areas = (x2 - x1) * (y2 - y1) on (x1,y1,x2,y2); ours are stored (x1,x2,y1,y2)
    area[j] = __fmul_rn(__fsub_rn(b.y, b.x), __fsub_rn(b.w, b.z));
    cls[j] = (int)(key[j] >> 48);
    supp[j] = 0;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nv; j += blockDim.x)
    if (j == 0 || cls[j] != cls[j - 1]) seg_start[atomicAdd(&n_seg, 1)] = j;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int s = wid; s < n_seg; s += nw) {                                                      // greedy NMS, one warp per class
    const int a = seg_start[s];
    int b = a + 1;
    while (b < nv && cls[b] == cls[a]) ++b;
    for (int i = a; i < b - 1; ++i) {
      __syncwarp();
      if (supp[i]) continue;
      const float4 bi = box[i];
      const float ai = area[i];
      for (int j = i + 1 + lane; j < b; j += 32) {
        if (supp[j]) continue;
        const float4 bj = box[j];
        float w = fmaxf(0.f, __fsub_rn(fminf(bi.y, bj.y), fmaxf(bi.x, bj.x)));
        float h = fmaxf(0.f, __fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.z, bj.z)));
        float inter = __fmul_rn(w, h);
        float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ai, area[j]), inter));
        if ((double)ovr > thresh) supp[j] = 1;                                                         // NaN (0/0) never suppresses
      }
    }
  }
  __syncthreads();
  for (int start = 0; start < nv; start += blockDim.x) {                                       // ordered compaction
    int j = start + threadIdx.x;
    bool keep = j < nv && !supp[j];
    int rank = block_excl_scan(keep ? 1 : 0, warp_cnt, base_rank);
    if (keep) {
      int src = (int)(key[j] & 0xffffu);
      st_label[e0 + rank] = cls[j];
      st_conf[e0 + rank] = ent_conf[e0 + src];
      st_box[e0 + rank] = box[j];
    }
  }
  if (threadIdx.x == 0) st_count[img] = base_rank;
}

// evaluate.py:368-370 + CSR packing: staging -> cats / conf / float boxes / int() boxes / super-categories / image ids
__global__ void proposals_pack_kernel(const int* __restrict__ st_label, const float* __restrict__ st_conf,
                                      const float4* __restrict__ st_box, const int* __restrict__ off, int n_ent,
                                      const int8_t* __restrict__ sub2super, int num_classes, int* __restrict__ cats,
                                      float* __restrict__ conf, float4* __restrict__ box_f, int4* __restrict__ box_i,
                                      int8_t* __restrict__ supers, int* __restrict__ box_img) {
  const int img = blockIdx.x;
  const int o0 = off[img], n = off[img + 1] - o0;
  const size_t e0 = (size_t)img * n_ent;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    int c = st_label[e0 + j];
    float4 b = st_box[e0 + j];
    cats[o0 + j] = c;
    conf[o0 + j] = st_conf[e0 + j];
    box_f[o0 + j] = b;
    box_i[o0 + j] = make_int4((int)b.x, (int)b.y, (int)b.z, (int)b.w);                         // int(): truncation toward zero
    if (supers) {
      char4 s = make_char4(-1, -1, -1, -1);
      if (c >= 0 && c < num_classes) s = reinterpret_cast<const char4*>(sub2super)[c];
      reinterpret_cast<char4*>(supers)[o0 + j] = s;
    }
    if (box_img) box_img[o0 + j] = img;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// utils.py:376-422 match_object_categories.  One warp per GT box: grid IoU (utils.py:58-74; double ratio rounded to
// float32 by torch.tensor(all_ious)) against every proposal of the image, top-2 under (IoU desc, proposal index asc).
__device__ __forceinline__ float grid_iou_f32(const Rect& a, const Rect& b) {
  int inter = rect_inter(a, b);
  int uni = rect_area(a) + rect_area(b) - inter;
  return uni == 0 ? 0.0f : (float)((double)inter / (double)uni);
}

__device__ __forceinline__ void warp_argmax(float& v, int& i) {
  for (int o = 16; o; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
}

__global__ void moc_top2_kernel(const float4* __restrict__ prop_box, const int* __restrict__ prop_off,
                                const int4* __restrict__ gt_box, const int* __restrict__ gt_off, int n_images, int fs,
                                int* __restrict__ best_idx, float* __restrict__ best_iou, int* __restrict__ out_count,
                                int* __restrict__ status) {
  const int img = blockIdx.x;
  const int p0 = prop_off[img], np = prop_off[img + 1] - p0;
  const int g0 = gt_off[img], ng = gt_off[img + 1] - g0;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  if (np < 2) {                                             // utils.py:402-403: `return None, None, None` from inside the GT loop
    if (threadIdx.x == 0) {
      if (ng > 0) atomicExch(status, 1);
      out_count[img] = 0;
    }
    return;
  }
  int local = 0;
  for (int k = wid; k < ng; k += nw) {
    const Rect t = rect_of(gt_box[g0 + k], fs);
    float v0 = -1.f, v1 = -1.f;
    int i0 = 0x7fffffff, i1 = 0x7fffffff;
    for (int j = lane; j < np; j += 32) {
      float4 b = prop_box[p0 + j];
      float v = grid_iou_f32(t, rect_of(make_int4((int)b.x, (int)b.y, (int)b.z, (int)b.w), fs));
      if (v > v0) { v0 = v; i0 = j; }                      // ascending j per lane: strict > keeps the lowest index
    }
    warp_argmax(v0, i0);
    for (int j = lane; j < np; j += 32) {
      if (j == i0) continue;
      float4 b = prop_box[p0 + j];
      float v = grid_iou_f32(t, rect_of(make_int4((int)b.x, (int)b.y, (int)b.z, (int)b.w), fs));
      if (v > v1) { v1 = v; i1 = j; }
    }
    warp_argmax(v1, i1);
    if (lane == 0) {
      best_idx[2 * (g0 + k)] = i0; best_idx[2 * (g0 + k) + 1] = i1;
      best_iou[2 * (g0 + k)] = v0; best_iou[2 * (g0 + k) + 1] = v1;
      local += 1 + (v0 == v1);
    }
  }
  if (lane == 0 && local) atomicAdd(&cnt, local);
  __syncthreads();
  if (threadIdx.x == 0) out_count[img] = cnt;
}

__global__ void moc_fill_kernel(const int* __restrict__ prop_cats, const float* __restrict__ prop_conf,
                                const int* __restrict__ prop_off, const int4* __restrict__ gt_box, const int* __restrict__ gt_off,
                                const int* __restrict__ best_idx, const float* __restrict__ best_iou, const int* __restrict__ out_off,
                                const int8_t* __restrict__ sub2super, int num_classes, int* __restrict__ out_cats,
                                float* __restrict__ out_conf, int4* __restrict__ out_box, int* __restrict__ out_src,
                                int8_t* __restrict__ out_supers, int* __restrict__ out_img) {
  const int img = blockIdx.x;
  const int p0 = prop_off[img];
  const int g0 = gt_off[img], ng = gt_off[img + 1] - g0;
  const int o0 = out_off[img];
  if (out_off[img + 1] == o0) return;
  __shared__ int warp_cnt[32], base_rank;
  if (threadIdx.x == 0) base_rank = 0;
  __syncthreads();
  for (int start = 0; start < ng; start += blockDim.x) {    // ordered: each GT box emits 1 or 2 consecutive rows
    const int k = start + threadIdx.x;
    const bool in = k < ng;
    int i0 = 0, i1 = 0;
    float v0 = 0.f, v1 = 0.f;
    if (in) {
      i0 = best_idx[2 * (g0 + k)]; i1 = best_idx[2 * (g0 + k) + 1];
      v0 = best_iou[2 * (g0 + k)]; v1 = best_iou[2 * (g0 + k) + 1];
    }
    const int emit = in ? 1 + (v0 == v1) : 0;
    const int pos = o0 + block_excl_scan(emit, warp_cnt, base_rank);
    for (int r = 0; r < emit; ++r) {
      const int j = r == 0 ? i0 : i1;
      const float v = r == 0 ? v0 : v1;
      const int c = prop_cats[p0 + j];
      out_cats[pos + r] = c;
      out_conf[pos + r] = __fmul_rn(prop_conf[p0 + j], v);  // utils.py:409-410,417: confidence * IoU in fp32
      out_box[pos + r] = gt_box[g0 + k];                    // utils.py:412-414: the GT box, repeated on a tie
      if (out_src) out_src[pos + r] = g0 + k;
      if (out_supers) {
        char4 sc = make_char4(-1, -1, -1, -1);
        if (c >= 0 && c < num_classes) sc = reinterpret_cast<const char4*>(sub2super)[c];
        reinterpret_cast<char4*>(out_supers)[pos + r] = sc;
      }
      if (out_img) out_img[pos + r] = img;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// utils.py:294-352 match_target_sgd: per image, the (g,e) loop emits a GT triplet where subj_or_obj is 1 (g subject) or
// 0 (e subject).  Reference quirk kept: the outer loop is range(len(relationships[image])) = range(N-1), so g <= N-2.
__device__ __forceinline__ void fe_tri_decode(int t, int& g, int& e) {
  g = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)t)) * 0.5f);
  while (g * (g - 1) / 2 > t) --g;
  while ((g + 1) * g / 2 <= t) ++g;
  e = t - g * (g - 1) / 2;
}

__global__ void targets_count_kernel(const int8_t* __restrict__ dir_tri, const int* __restrict__ tri_off,
                                     const int* __restrict__ box_off, int* __restrict__ counts) {
  const int img = blockIdx.x;
  const int n = box_off[img + 1] - box_off[img];
  const int T = n >= 2 ? (n - 1) * (n - 2) / 2 : 0, t0 = tri_off[img];
  int c = 0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    int d = dir_tri[t0 + t];
    c += (d == 0 || d == 1);
  }
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

__global__ void targets_fill_kernel(const int8_t* __restrict__ dir_tri, const int* __restrict__ rel_tri,
                                    const int* __restrict__ tri_off, const int* __restrict__ box_off,
                                    const int* __restrict__ gt_off, int* __restrict__ gt_label, int* __restrict__ gt_sub,
                                    int* __restrict__ gt_obj) {
  const int img = blockIdx.x;
  const int b0 = box_off[img], n = box_off[img + 1] - b0;
  const int T = n >= 2 ? (n - 1) * (n - 2) / 2 : 0, t0 = tri_off[img];
  const int o0 = gt_off[img];
  __shared__ int warp_cnt[32], base_rank;
  if (threadIdx.x == 0) base_rank = 0;
  __syncthreads();
  for (int start = 0; start < T; start += blockDim.x) {
    const int t = start + threadIdx.x;
    int d = -1;
    if (t < T) d = dir_tri[t0 + t];
    const bool keep = d == 0 || d == 1;
    const int pos = o0 + block_excl_scan(keep ? 1 : 0, warp_cnt, base_rank);
    if (keep) {
      int g, e;
      fe_tri_decode(t, g, e);
      gt_label[pos] = rel_tri[t0 + t];
      gt_sub[pos] = b0 + (d == 1 ? g : e);
      gt_obj[pos] = b0 + (d == 1 ? e : g);
    }
  }
}

}  // namespace hc

using namespace hc;

static int pow2_at_least(int n) {
  int p = 2;
  while (p < n) p <<= 1;
  return p;
}

extern "C" int hc_detr_proposals(const float* pred_logits, const float* pred_boxes, int32_t n_images, int32_t n_queries,
                                 int32_t num_classes, int32_t topk_cat, const int32_t* label_map, int32_t feature_size,
                                 double nms_thresh, int32_t* ws_label, float* ws_conf, float* ws_box, uint8_t* ws_valid,
                                 int32_t* st_label, float* st_conf, float* st_box, int32_t* st_count, int32_t* box_offsets,
                                 hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(pred_logits && pred_boxes && label_map && ws_label && ws_conf && ws_box && ws_valid && st_label && st_conf && st_box &&
                 st_count && box_offsets,
             HC_E_NULL, "hc_detr_proposals: required pointer is NULL");
  HC_REQUIRE(n_images > 0 && n_queries > 0 && topk_cat >= 1 && topk_cat <= 4 && num_classes >= 1 && num_classes + 1 <= 256 &&
                 feature_size > 0,
             HC_E_SHAPE, "hc_detr_proposals: need n_images,n_queries > 0, 1 <= topk_cat <= 4, num_classes + 1 <= 256");
  const int n_ent = n_queries * topk_cat;
  HC_REQUIRE(n_ent <= FE_MAX_ENTRIES, HC_E_SHAPE, "hc_detr_proposals: n_queries * topk_cat must be <= 1024");
  HC_REQUIRE(aligned16(pred_boxes) && aligned16(ws_box) && aligned16(st_box), HC_E_ALIGN, "hc_detr_proposals: box arrays must be 16-byte aligned");
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  const int n_rows = n_images * n_queries;
  detr_expand_kernel<<<(n_rows + 7) / 8, 256, 0, stream>>>(pred_logits, pred_boxes, n_rows, num_classes + 1, topk_cat, label_map,
                                                           (float)feature_size, ws_label, ws_conf, reinterpret_cast<float4*>(ws_box),
                                                           ws_valid);
  const int n_pow2 = pow2_at_least(n_ent);
  const size_t smem = (size_t)n_pow2 * (8 + 16 + 4 + 4 + 4 + 1);
  proposals_nms_kernel<<<n_images, FE_THREADS, smem, stream>>>(ws_label, ws_conf, reinterpret_cast<const float4*>(ws_box), ws_valid,
                                                               n_ent, n_pow2, nms_thresh, st_label, st_conf,
                                                               reinterpret_cast<float4*>(st_box), st_count);
  fe_scan_kernel<<<1, 1024, 0, stream>>>(st_count, n_images, box_offsets, nullptr);
  return cuda_status("hc_detr_proposals");
}

extern "C" int hc_proposals_pack(const int32_t* st_label, const float* st_conf, const float* st_box, const int32_t* box_offsets,
                                 int32_t n_images, int32_t n_entries, const int8_t* sub2super, int32_t num_classes, int32_t* cats,
                                 float* conf, float* box_f, int32_t* box_i, int8_t* supers, int32_t* box_img, hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(st_label && st_conf && st_box && box_offsets && cats && conf && box_f && box_i, HC_E_NULL,
             "hc_proposals_pack: required pointer is NULL");
  HC_REQUIRE(!supers || sub2super, HC_E_NULL, "hc_proposals_pack: supers requested without the sub2super table");
  HC_REQUIRE(n_images > 0 && n_entries > 0, HC_E_SHAPE, "hc_proposals_pack: n_images and n_entries must be positive");
  HC_REQUIRE(aligned16(st_box) && aligned16(box_f) && aligned16(box_i), HC_E_ALIGN, "hc_proposals_pack: box arrays must be 16-byte aligned");
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  proposals_pack_kernel<<<n_images, 128, 0, stream>>>(st_label, st_conf, reinterpret_cast<const float4*>(st_box), box_offsets, n_entries,
                                                      sub2super, num_classes, cats, conf, reinterpret_cast<float4*>(box_f),
                                                      reinterpret_cast<int4*>(box_i), supers, box_img);
  return cuda_status("hc_proposals_pack");
}

extern "C" int hc_match_object_categories(const float* prop_box, const int32_t* prop_offsets, const int32_t* gt_box,
                                          const int32_t* gt_offsets, int32_t n_images, int32_t feature_size, int32_t* ws_idx,
                                          float* ws_iou, int32_t* ws_count, int32_t* out_offsets, int32_t* status,
                                          hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(prop_box && prop_offsets && gt_box && gt_offsets && ws_idx && ws_iou && ws_count && out_offsets && status, HC_E_NULL,
             "hc_match_object_categories: required pointer is NULL");
  HC_REQUIRE(n_images > 0 && feature_size > 0, HC_E_SHAPE, "hc_match_object_categories: n_images and feature_size must be positive");
  HC_REQUIRE(aligned16(prop_box) && aligned16(gt_box), HC_E_ALIGN, "hc_match_object_categories: box arrays must be 16-byte aligned");
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  cudaMemsetAsync(status, 0, sizeof(int32_t), stream);
  moc_top2_kernel<<<n_images, 256, 0, stream>>>(reinterpret_cast<const float4*>(prop_box), prop_offsets,
                                                reinterpret_cast<const int4*>(gt_box), gt_offsets, n_images, feature_size, ws_idx, ws_iou,
                                                ws_count, status);
  fe_scan_kernel<<<1, 1024, 0, stream>>>(ws_count, n_images, out_offsets, nullptr);
  return cuda_status("hc_match_object_categories");
}

extern "C" int hc_match_object_categories_fill(const int32_t* prop_cats, const float* prop_conf, const int32_t* prop_offsets,
                                               const int32_t* gt_box, const int32_t* gt_offsets, int32_t n_images,
                                               const int32_t* ws_idx, const float* ws_iou, const int32_t* out_offsets,
                                               const int8_t* sub2super, int32_t num_classes, int32_t* out_cats, float* out_conf,
                                               int32_t* out_box, int32_t* out_src, int8_t* out_supers, int32_t* out_img,
                                               hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(prop_cats && prop_conf && prop_offsets && gt_box && gt_offsets && ws_idx && ws_iou && out_offsets && out_cats && out_conf &&
                 out_box,
             HC_E_NULL, "hc_match_object_categories_fill: required pointer is NULL");
  HC_REQUIRE(!out_supers || sub2super, HC_E_NULL, "hc_match_object_categories_fill: supers requested without the sub2super table");
  HC_REQUIRE(n_images > 0, HC_E_SHAPE, "hc_match_object_categories_fill: n_images must be positive");
  HC_REQUIRE(aligned16(gt_box) && aligned16(out_box), HC_E_ALIGN, "hc_match_object_categories_fill: box arrays must be 16-byte aligned");
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  moc_fill_kernel<<<n_images, 256, 0, stream>>>(prop_cats, prop_conf, prop_offsets, reinterpret_cast<const int4*>(gt_box), gt_offsets,
                                                ws_idx, ws_iou, out_offsets, sub2super, num_classes, out_cats, out_conf,
                                                reinterpret_cast<int4*>(out_box), out_src, out_supers, out_img);
  return cuda_status("hc_match_object_categories_fill");
}

extern "C" int hc_targets_flat(const int8_t* dir_tri, const int32_t* rel_tri, const int32_t* tri_offsets, const int32_t* box_offsets,
                               int32_t n_images, int32_t* ws_count, int32_t* gt_offsets, int32_t* gt_label, int32_t* gt_sub,
                               int32_t* gt_obj, hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(dir_tri && rel_tri && tri_offsets && box_offsets && ws_count && gt_offsets && gt_label && gt_sub && gt_obj, HC_E_NULL,
             "hc_targets_flat: required pointer is NULL");
  HC_REQUIRE(n_images > 0, HC_E_SHAPE, "hc_targets_flat: n_images must be positive");
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  targets_count_kernel<<<n_images, 256, 0, stream>>>(dir_tri, tri_offsets, box_offsets, ws_count);
  fe_scan_kernel<<<1, 1024, 0, stream>>>(ws_count, n_images, gt_offsets, nullptr);
  targets_fill_kernel<<<n_images, 256, 0, stream>>>(dir_tri, rel_tri, tri_offsets, box_offsets, gt_offsets, gt_label, gt_sub, gt_obj);
  return cuda_status("hc_targets_flat");
}
