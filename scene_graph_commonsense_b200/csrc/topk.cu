// R10-R13: per-image top-K selection, first-match scan and integer hit/GT counters
// (evaluator.py:294-356 Evaluator.compute, :704-766 Evaluator_Top3.compute).
//
// One CTA per image.  Selection order is (confidence descending, candidate index ascending) == a stable
// descending sort; -inf candidates are legal members of the top-K.  Steps:
//   1. 4-pass radix select (8 bits/pass, shared-memory histograms) for the key of the K-th best candidate
//   2. ordered compaction (ballot scans in candidate-index order) of everything better than the threshold plus the
//      first ties in index order
//   3. bitonic sort of the <=128 survivors on 64-bit (key, index)
//   4. one thread per GT slot scans the sorted survivors held in shared memory; hits go to shared-memory
//      counters, flushed once per CTA with 64-bit global atomics (non-zero slots only).
#include "hc_common.cuh"

namespace hc {

constexpr int TK_THREADS = 256;
constexpr int TK_MAX = HC_TOP_MAX;       // 128
constexpr int NPRED = 50;
constexpr int NKK = 3;
// counter layouts (must match scene_graph_commonsense_b200/tables.py)
constexpr int EV_HITS = 0, EV_HITS_PC = 3, EV_NGT = 3 + 3 * NPRED, EV_NGT_PC = EV_NGT + 1, EV_BLOCK = EV_NGT_PC + NPRED;
constexpr int T3_HITS = 0, T3_HITS_PC = 3, T3_TOP1 = 3 + 3 * NPRED, T3_TOP1_PC = T3_TOP1 + 3, T3_NGT = T3_TOP1_PC + 3 * NPRED,
              T3_NGT_PC = T3_NGT + 1, T3_SIZE = T3_NGT_PC + NPRED;
constexpr int CNT_MAX = 2 * EV_BLOCK;    // 408 >= 357

// descending-order key: smaller key == larger confidence; -0.0 is canonicalised to +0.0 (torch compares them equal)
__device__ __forceinline__ uint32_t desc_key(float f) {
  if (f == 0.0f) f = 0.0f;
  uint32_t u = __float_as_uint(f);
  uint32_t asc = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ~asc;
}

struct SelMeta {
  int label;           // mode 0: candidate label
  int lab3[3];         // mode 1: the three per-head labels of the row
  int sup;             // mode 1: argmax super-category of the row
  int cs, co;
  Rect rs, ro;
};

__global__ void __launch_bounds__(TK_THREADS)
topk_match_kernel(const int* __restrict__ cand_off, const float* __restrict__ cand_conf, const int* __restrict__ cand_label,
                  const int* __restrict__ cand_row, int K, const int* __restrict__ row_sub, const int* __restrict__ row_obj,
                  const int* __restrict__ pred_cat, const int4* __restrict__ pred_box, const int* __restrict__ gt_off,
                  const int* __restrict__ gt_label, const int* __restrict__ gt_sub, const int* __restrict__ gt_obj,
                  const int* __restrict__ gt_cat, const int4* __restrict__ gt_box, const uint8_t* __restrict__ synonyms, int num_obj,
                  const uint32_t* __restrict__ zs_bitmap, int fs, double iou_thresh, int top_max, int k0, int k1, int k2, int mode,
                  const int* __restrict__ t3_labels, const uint8_t* __restrict__ t3_super, unsigned long long* __restrict__ counters,
                  int* __restrict__ topk_out, int select_only) {
  __shared__ int hist[256];
  __shared__ int warp_cnt[2][TK_THREADS / 32];
  __shared__ unsigned long long sel[TK_MAX];
  __shared__ SelMeta meta[TK_MAX];
  __shared__ int cnt[CNT_MAX];
  __shared__ uint32_t s_prefix;
  __shared__ int s_remaining, s_base_lt, s_base_eq, s_ngt;

  const int img = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int c0 = cand_off[img];
  const int C = cand_off[img + 1] - c0;
  if (topk_out)
    for (int j = tid; j < top_max; j += TK_THREADS) topk_out[(long long)img * top_max + j] = -1;
  if (C <= 0) return;                                   // image absent from torch.unique(which_in_batch) (evaluator.py:294)
  const int g0 = select_only ? 0 : gt_off[img], G = select_only ? 0 : gt_off[img + 1] - g0;
  if (G <= 0 && !topk_out) return;
  const int n_sel = min(top_max, C);                    // evaluator.py:315-316
  const float* conf = cand_conf + c0;

  // ---- 1. radix select: key of the n_sel-th best
  if (tid == 0) { s_prefix = 0; s_remaining = n_sel; }
  for (int pass = 3; pass >= 0; --pass) {
    for (int i = tid; i < 256; i += TK_THREADS) hist[i] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    const uint32_t hi_mask = pass == 3 ? 0u : (0xFFFFFFFFu << (8 * (pass + 1)));
    for (int c = tid; c < C; c += TK_THREADS) {
      uint32_t k = desc_key(conf[c]);
      if ((k & hi_mask) == (prefix & hi_mask)) atomicAdd(&hist[(k >> (8 * pass)) & 255], 1);
    }
    __syncthreads();
    if (tid == 0) {
      int rem = s_remaining, b = 0, acc = 0;
      for (; b < 256; ++b) {
        if (acc + hist[b] >= rem) break;
        acc += hist[b];
      }
      s_remaining = rem - acc;
      s_prefix = prefix | ((uint32_t)b << (8 * pass));
    }
    __syncthreads();
  }
  const uint32_t thr = s_prefix;                        // exact key of the n_sel-th element
  const int take_eq = s_remaining;                      // how many ties (in index order) make the cut

  // ---- 2. ordered compaction
  if (tid == 0) { s_base_lt = 0; s_base_eq = 0; }
  for (int i = tid; i < TK_MAX; i += TK_THREADS) sel[i] = ~0ull;
  __syncthreads();
  const int n_lt_total = n_sel - take_eq;
  for (int start = 0; start < C; start += TK_THREADS) {
    int c = start + tid;
    uint32_t k = c < C ? desc_key(conf[c]) : 0xFFFFFFFFu;
    bool lt = c < C && k < thr;
    bool eq = c < C && k == thr;
    unsigned mlt = __ballot_sync(0xffffffffu, lt), meq = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) { warp_cnt[0][wid] = __popc(mlt); warp_cnt[1][wid] = __popc(meq); }
    __syncthreads();
    int olt = s_base_lt, oeq = s_base_eq;
    for (int w = 0; w < wid; ++w) { olt += warp_cnt[0][w]; oeq += warp_cnt[1][w]; }
    olt += __popc(mlt & ((1u << lane) - 1u));
    oeq += __popc(meq & ((1u << lane) - 1u));
    if (lt) sel[olt] = ((unsigned long long)k << 32) | (uint32_t)c;
    if (eq && oeq < take_eq) sel[n_lt_total + oeq] = ((unsigned long long)k << 32) | (uint32_t)c;
    __syncthreads();
    if (tid == 0) {
      for (int w = 0; w < TK_THREADS / 32; ++w) { s_base_lt += warp_cnt[0][w]; s_base_eq += warp_cnt[1][w]; }
    }
    __syncthreads();
  }

  // ---- 3. bitonic sort of 128 (key, index) words; padding = all ones sorts last
  for (int size = 2; size <= TK_MAX; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (tid < TK_MAX / 2) {
        int lo = 2 * tid - (tid & (stride - 1));
        int hi = lo + stride;
        bool up = (lo & size) == 0;
        unsigned long long a = sel[lo], b = sel[hi];
        if ((a > b) == up) { sel[lo] = b; sel[hi] = a; }
      }
      __syncthreads();
    }
  }

  // ---- survivors' metadata into shared memory
  for (int j = tid; j < n_sel; j += TK_THREADS) {
    int c = (int)(sel[j] & 0xFFFFFFFFu);
    if (topk_out) topk_out[(long long)img * top_max + j] = c;
    if (select_only) continue;
    int gc = c0 + c;
    int row = cand_row ? cand_row[gc] : gc / K;
    SelMeta m;
    m.label = cand_label ? cand_label[gc] : -1;
    m.lab3[0] = m.lab3[1] = m.lab3[2] = -1;
    m.sup = 0;
    if (mode == 1) {
      m.lab3[0] = t3_labels[(long long)row * 3]; m.lab3[1] = t3_labels[(long long)row * 3 + 1]; m.lab3[2] = t3_labels[(long long)row * 3 + 2];
      m.sup = t3_super[row];
    }
    int bs = row_sub[row], bo = row_obj[row];
    m.cs = pred_cat[bs]; m.co = pred_cat[bo];
    m.rs = rect_of(pred_box[bs], fs); m.ro = rect_of(pred_box[bo], fs);
    meta[j] = m;
  }
  for (int i = tid; i < CNT_MAX; i += TK_THREADS) cnt[i] = 0;
  if (tid == 0) s_ngt = 0;
  __syncthreads();
  if (G <= 0) return;

  // Evaluator_Top3 cut-off uses the number of connected targets of the image (evaluator.py:716,739)
  if (mode == 1) {
    int n = 0;
    for (int g = tid; g < G; g += TK_THREADS) n += gt_label[g0 + g] != -1;
    for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if (lane == 0 && n) atomicAdd(&s_ngt, n);
    __syncthreads();
  }
  const int ngt_img = s_ngt;
  const int ks[NKK] = {k0, k1, k2};

  // ---- 4. first-match scan, one thread per GT slot
  for (int g = tid; g < G; g += TK_THREADS) {
    const int t = gt_label[g0 + g];
    if (t == -1) continue;                              // evaluator.py:307
    const int a = gt_sub[g0 + g], b = gt_obj[g0 + g];
    const int tcs = gt_cat[a], tco = gt_cat[b];
    const Rect trs = rect_of(gt_box[a], fs), tro = rect_of(gt_box[b], fs);
    const bool t_ok = t >= 0 && t < NPRED;
    bool zs = false;
    if (mode == 0 && zs_bitmap && t_ok && tcs >= 0 && tcs < HC_NUM_OBJ && tco >= 0 && tco < HC_NUM_OBJ)
      zs = bitmap_test(zs_bitmap, (tcs * 50 + t) * HC_NUM_OBJ + tco);
    bool found = false, found1 = false;
    for (int j = 0; j < n_sel; ++j) {
      const SelMeta& m = meta[j];
      bool lab;
      if (synonyms)                                     // utils.compare_object_cat (evaluator.py:324-325)
        lab = synonyms[tcs * num_obj + m.cs] && synonyms[tco * num_obj + m.co];
      else
        lab = tcs == m.cs && tco == m.co;               // evaluator.py:321-322
      if (!lab) continue;
      if (!(grid_iou_ge(trs, m.rs, iou_thresh) && grid_iou_ge(tro, m.ro, iou_thresh))) continue;
      if (mode == 0) {
        if (t == m.label) {                             // evaluator.py:331-348
#pragma unroll
          for (int q = 0; q < NKK; ++q)
            if (j < ks[q]) {
              atomicAdd(&cnt[EV_HITS + q], 1);
              if (t_ok) atomicAdd(&cnt[EV_HITS_PC + q * NPRED + t], 1);
              if (zs) {
                atomicAdd(&cnt[EV_BLOCK + EV_HITS + q], 1);
                atomicAdd(&cnt[EV_BLOCK + EV_HITS_PC + q * NPRED + t], 1);
              }
            }
          break;
        }
      } else {
        if (!found && (t == m.lab3[0] || t == m.lab3[1] || t == m.lab3[2])) {   // evaluator.py:730-744
#pragma unroll
          for (int q = 0; q < NKK; ++q)
            if (j < max(ks[q], ngt_img)) {
              atomicAdd(&cnt[T3_HITS + q], 1);
              if (t_ok) atomicAdd(&cnt[T3_HITS_PC + q * NPRED + t], 1);
            }
          found = true;
        }
        if (!found1 && t == (m.sup == 0 ? m.lab3[0] : (m.sup == 1 ? m.lab3[1] : m.lab3[2]))) {   // :746-760
#pragma unroll
          for (int q = 0; q < NKK; ++q)
            if (j < max(ks[q], ngt_img)) {
              atomicAdd(&cnt[T3_TOP1 + q], 1);
              if (t_ok) atomicAdd(&cnt[T3_TOP1_PC + q * NPRED + t], 1);
            }
          found1 = true;
        }
        if (found && found1) break;
      }
    }
    if (mode == 0) {                                    // evaluator.py:350-356
      atomicAdd(&cnt[EV_NGT], 1);
      if (t_ok) atomicAdd(&cnt[EV_NGT_PC + t], 1);
      if (zs) { atomicAdd(&cnt[EV_BLOCK + EV_NGT], 1); atomicAdd(&cnt[EV_BLOCK + EV_NGT_PC + t], 1); }
    } else {                                            // evaluator.py:765-766
      atomicAdd(&cnt[T3_NGT], 1);
      if (t_ok) atomicAdd(&cnt[T3_NGT_PC + t], 1);
    }
  }
  __syncthreads();
  const int n_cnt = mode == 0 ? 2 * EV_BLOCK : T3_SIZE;
  for (int i = tid; i < n_cnt; i += TK_THREADS)
    if (cnt[i]) atomicAdd(counters + i, (unsigned long long)cnt[i]);
}

}  // namespace hc

using namespace hc;

extern "C" int hc_topk_match(const int32_t* cand_offsets, int32_t n_images, const float* cand_conf, const int32_t* cand_label,
                             const int32_t* cand_row, int32_t k_per_row, const int32_t* row_sub, const int32_t* row_obj,
                             const int32_t* pred_cat, const int32_t* pred_box, const int32_t* gt_offsets, const int32_t* gt_label,
                             const int32_t* gt_sub, const int32_t* gt_obj, const int32_t* gt_cat, const int32_t* gt_box,
                             const uint8_t* synonyms, int32_t num_obj, int32_t num_pred, const uint32_t* zs_bitmap,
                             int32_t feature_size, double iou_thresh, int32_t top_max, int32_t k0, int32_t k1, int32_t k2, int32_t mode,
                             const int32_t* t3_labels, const uint8_t* t3_super, unsigned long long* counters, int32_t* topk_out,
                             hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(cand_offsets && cand_conf && row_sub && row_obj && pred_cat && pred_box && gt_offsets && gt_label && gt_sub && gt_obj &&
                 gt_cat && gt_box && counters,
             HC_E_NULL, "hc_topk_match: required pointer is NULL");
  HC_REQUIRE(mode == 0 || mode == 1, HC_E_SHAPE, "hc_topk_match: mode must be 0 (Evaluator) or 1 (Evaluator_Top3)");
  HC_REQUIRE(mode == 0 ? cand_label != nullptr : (t3_labels && t3_super && k_per_row == 1), HC_E_NULL,
             "hc_topk_match: mode 0 needs cand_label; mode 1 needs t3_labels, t3_super and k_per_row == 1");
  HC_REQUIRE(k_per_row >= 1, HC_E_SHAPE, "hc_topk_match: k_per_row must be >= 1");
  HC_REQUIRE(top_max >= 1 && top_max <= HC_TOP_MAX, HC_E_SHAPE, "hc_topk_match: top_max must be in [1,128]");
  HC_REQUIRE(num_pred == NPRED, HC_E_SHAPE, "hc_topk_match: counters are laid out for 50 predicate classes");
  HC_REQUIRE(num_obj > 0 && num_obj <= HC_NUM_OBJ && feature_size > 0, HC_E_SHAPE, "hc_topk_match: bad num_obj/feature_size");
  HC_REQUIRE(aligned16(pred_box) && aligned16(gt_box), HC_E_ALIGN, "hc_topk_match: box tables must be 16-byte aligned");
  if (n_images <= 0) return HC_OK;
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  topk_match_kernel<<<n_images, TK_THREADS, 0, stream>>>(
      cand_offsets, cand_conf, cand_label, cand_row, k_per_row, row_sub, row_obj, pred_cat, reinterpret_cast<const int4*>(pred_box),
      gt_offsets, gt_label, gt_sub, gt_obj, gt_cat, reinterpret_cast<const int4*>(gt_box), synonyms, num_obj, zs_bitmap, feature_size,
      iou_thresh, top_max, k0, k1, k2, mode, t3_labels, t3_super, counters, topk_out, 0);
  return cuda_status("hc_topk_match");
}

extern "C" int hc_topk_select(const int32_t* cand_offsets, int32_t n_images, const float* cand_conf, int32_t top_max, int32_t* topk_out,
                              hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(cand_offsets && cand_conf && topk_out, HC_E_NULL, "hc_topk_select: NULL pointer");
  HC_REQUIRE(top_max >= 1 && top_max <= HC_TOP_MAX, HC_E_SHAPE, "hc_topk_select: top_max must be in [1,128]");
  if (n_images <= 0) return HC_OK;
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  topk_match_kernel<<<n_images, TK_THREADS, 0, stream>>>(cand_offsets, cand_conf, nullptr, nullptr, 1, nullptr, nullptr, nullptr, nullptr,
                                                         nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 1, nullptr, 32, 0.5,
                                                         top_max, 0, 0, 0, 0, nullptr, nullptr, nullptr, topk_out, 1);
  return cuda_status("hc_topk_select");
}
