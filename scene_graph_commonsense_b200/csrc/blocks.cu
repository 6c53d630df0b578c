// Active-region work list for the block-sparse conv3_1 (HC_GEMM_CONV3_BLOCKS) and the background pre-fill of its output.
//
// train_test.py:391,398 multiplies the feature map by a rectangular box mask, so after conv1_x + tanh every pixel outside the
// box holds tanh(bias): a constant of the WEIGHTS.  Each later stage (3x3 conv, 2x2 pool, 3x3 conv, 2x2 pool, model.py:143-146)
// grows the region that can differ from that weights-only background by its receptive field; outside it the pooled conv3_1
// output of every pair equals one [8,8,1024] tensor computed once per checkpoint.  The GEMM therefore only has to visit the
// dilated footprint of the pair's two boxes; everything else is a broadcast copy.
#include "hc_common.cuh"

namespace hc {

// box interval [lo,hi) on the 32-grid -> half-open interval of 8-grid cells whose pooled conv3_1 output may differ from background
__device__ __forceinline__ void active_cells(int lo, int hi, int& a, int& b) {
  if (hi <= lo) { a = b = 0; return; }
  int qlo = max(0, (lo - 1) >> 1), qhi = min(15, hi >> 1);      // pooled conv2 pixels touched by the box +-1 (arithmetic shift: -1>>1 = -1)
  qlo = max(0, qlo - 1); qhi = min(15, qhi + 1);                // +-1 again: conv3_1's 3x3 window
  a = qlo >> 1; b = (qhi >> 1) + 1;
}

__device__ __forceinline__ unsigned long long cell_mask(const Rect& r) {
  int xa, xb, ya, yb;
  active_cells(r.x0, r.x1, xa, xb);
  active_cells(r.y0, r.y1, ya, yb);
  if (xb <= xa || yb <= ya) return 0ull;
  const unsigned long long rowbits = ((1ull << (xb - xa)) - 1ull) << xa;
  const unsigned long long rows = (yb - ya >= 8) ? ~0ull : (((1ull << (8 * (yb - ya))) - 1ull) << (8 * ya));
  return (0x0101010101010101ull & rows) * rowbits;
}

// greedy cover with blocks of wc x hc cells (wide x tall); emit(entry) is called once per block, returns the count
template <typename F>
__device__ __forceinline__ int cover_greedy(unsigned long long m, int hc, int wc, F emit) {
  int n = 0;
  const unsigned long long rows = (hc == 4) ? 0x01010101ull : ((hc == 2) ? 0x0101ull : 0x01ull);
  const unsigned long long cols = (1ull << wc) - 1ull;
  while (m) {
    const int bit = __ffsll((long long)m) - 1;
    const int y = min(bit >> 3, 8 - hc), x = min(bit & 7, 8 - wc);
    m &= ~((rows * cols) << (8 * y + x));
    emit((y << 4) | x);
    ++n;
  }
  return n;
}

// the greedy cover can need one block more than the aligned tiling of the whole map (two staggered rectangles): never emit
// more than the dense 2 x (8/hc) tiling, so a pair costs at most what the dense kernel would
template <typename F>
__device__ __forceinline__ int cover(unsigned long long m, int hc, int wc, F emit) {
  const int full = (8 / wc) * (8 / hc);
  if (cover_greedy(m, hc, wc, [](int) {}) <= full) return cover_greedy(m, hc, wc, emit);
  for (int y = 0; y < 8; y += hc)
    for (int x = 0; x < 8; x += wc) emit((y << 4) | x);
  return full;
}

__global__ void __launch_bounds__(1024)
conv3_blocks_kernel(const int4* __restrict__ boxes, const int* __restrict__ pair_sub, const int* __restrict__ pair_obj, int n_pairs, int fs,
                    int hc, int wc, bool both, int* __restrict__ blocks, int* __restrict__ n_blocks) {
  __shared__ int s_scan[1024];
  const int t = threadIdx.x;
  const int per = (n_pairs + 1023) / 1024;
  const int p0 = min(n_pairs, t * per), p1 = min(n_pairs, p0 + per);
  int cnt = 0;
  // both: only the cells BOTH boxes can reach (everything else comes from per-box maps, see p3_assemble_kernel)
  auto pair_mask = [&](int p) {
    const unsigned long long ms = cell_mask(rect_of(__ldg(boxes + pair_sub[p]), fs)), mo = cell_mask(rect_of(__ldg(boxes + pair_obj[p]), fs));
    return both ? (ms & mo) : (ms | mo);
  };
  for (int p = p0; p < p1; ++p) cnt += cover(pair_mask(p), hc, wc, [](int) {});
  s_scan[t] = cnt;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {                          // inclusive Hillis-Steele scan over the 1024 per-thread counts
    const int v = t >= d ? s_scan[t - d] : 0;
    __syncthreads();
    s_scan[t] += v;
    __syncthreads();
  }
  int o = s_scan[t] - cnt;
  if (t == 1023) n_blocks[0] = s_scan[t];
  for (int p = p0; p < p1; ++p) cover(pair_mask(p), hc, wc, [&](int e) { blocks[o++] = (p << 8) | e; });
}

// cells covered by the blocks conv3_blocks_kernel lists for each pair (same cover function, so the two can never disagree): the
// tiled pooling kernel writes a pooled conv2 pixel of a pair only if a listed block reads it (block + 1-pixel halo)
__global__ void pair_cover_masks_kernel(const int4* __restrict__ boxes, const int* __restrict__ pair_sub, const int* __restrict__ pair_obj,
                                        int n_pairs, int fs, int hc, int wc, bool both, unsigned long long* __restrict__ masks) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  const unsigned long long ms = cell_mask(rect_of(__ldg(boxes + pair_sub[p]), fs)), mo = cell_mask(rect_of(__ldg(boxes + pair_obj[p]), fs));
  const unsigned long long rows = (hc == 4) ? 0x01010101ull : ((hc == 2) ? 0x0101ull : 0x01ull);
  const unsigned long long blk = rows * ((1ull << wc) - 1ull);
  unsigned long long c = 0ull;
  cover(both ? (ms & mo) : (ms | mo), hc, wc, [&](int e) { c |= blk << (8 * (e >> 4) + (e & 15)); });
  masks[p] = c;
}


// conv2_1 halves on the BOX footprint: a box's conv2 output differs from the background map only within one pixel of the box
// rectangle (3x3 conv of a map that equals tanh(bias) outside the box).  Work list of 8 x bh-pixel blocks (even origins, clamped
// into the 32 x 32 map) covering that rectangle, same entry format as the conv3_1 lists; an empty box lists nothing.
__global__ void __launch_bounds__(1024)
conv2_blocks_kernel(const int4* __restrict__ boxes, int n_box, int fs, int bh, int* __restrict__ blocks, int* __restrict__ n_blocks) {
  __shared__ int s_scan[1024];
  const int t = threadIdx.x;
  const int per = (n_box + 1023) / 1024;
  const int b0 = min(n_box, t * per), b1 = min(n_box, b0 + per);
  auto walk = [&](int b, auto emit) {
    const Rect r = rect_of(__ldg(boxes + b), fs);
    if (r.x1 <= r.x0 || r.y1 <= r.y0) return 0;
    const int xlo = max(0, r.x0 - 1) & ~1, xhi = min(fs, r.x1 + 1), ylo = max(0, r.y0 - 1) & ~1, yhi = min(fs, r.y1 + 1);
    int n = 0;
    for (int y = ylo; y < yhi; y += bh)
      for (int x = xlo; x < xhi; x += 8) {
        emit((b << 8) | ((min(y, fs - bh) >> 1) << 4) | (min(x, fs - 8) >> 1));
        ++n;
      }
    return n;
  };
  int cnt = 0;
  for (int b = b0; b < b1; ++b) cnt += walk(b, [](int) {});
  s_scan[t] = cnt;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    const int v = t >= d ? s_scan[t - d] : 0;
    __syncthreads();
    s_scan[t] += v;
    __syncthreads();
  }
  int o = s_scan[t] - cnt;
  if (t == 1023) n_blocks[0] = s_scan[t];
  for (int b = b0; b < b1; ++b) walk(b, [&](int e) { blocks[o++] = e; });
}

// Pooled conv3_1 output of a pair outside the cells both boxes reach: a cell only the subject's box reaches equals the map of
// the pair (subject, EMPTY box), one only the object's box reaches equals (EMPTY box, object), the rest is the background.
// One CTA per pair; a cell is 1024 channels = 128 uint4, so 256 threads move two cells per iteration.  Cells both boxes
// reach are left alone: the work list of hc_conv3_shared_blocks covers them and HC_GEMM_CONV3_BLOCKS writes them afterwards.
__global__ void __launch_bounds__(256)
p3_assemble_kernel(const uint4* __restrict__ bg, const uint4* __restrict__ sub_maps, const uint4* __restrict__ obj_maps,
                   const int4* __restrict__ boxes, const int* __restrict__ pair_sub, const int* __restrict__ pair_obj, int n_pairs, int fs,
                   uint4* __restrict__ out) {
  const int lane128 = threadIdx.x & 127, half = threadIdx.x >> 7;
  for (int p = blockIdx.x; p < n_pairs; p += gridDim.x) {
    const int s = pair_sub[p], o = pair_obj[p];
    const unsigned long long ms = cell_mask(rect_of(__ldg(boxes + s), fs)), mo = cell_mask(rect_of(__ldg(boxes + o), fs));
    const uint4* ps = sub_maps + (long long)s * 8192;
    const uint4* po = obj_maps + (long long)o * 8192;
    uint4* dst = out + (long long)p * 8192;
#pragma unroll 4
    for (int c = half; c < 64; c += 2) {
      const bool in_s = (ms >> c) & 1ull, in_o = (mo >> c) & 1ull;
      if (in_s && in_o) continue;
      const uint4* src = in_s ? ps : (in_o ? po : bg);
      __stcs(dst + c * 128 + lane128, __ldg(src + c * 128 + lane128));
    }
  }
}

// a thread owns one 16-byte column of the row: one load, then a streaming store per destination row
__global__ void broadcast_rows_kernel(const uint4* __restrict__ src, long long row_vecs, long long n_rows, uint4* __restrict__ out) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= row_vecs) return;
  const uint4 v = __ldg(src + c);
  for (long long r = blockIdx.y; r < n_rows; r += gridDim.y) __stcs(out + r * row_vecs + c, v);
}


// ---- shared-footprint fc1: row order, per-tile K-cell masks, zero fill of the operand --------------------------------------------
// A directed pair's fc1 operand d (HC_EPI_POOL_DIFF_BF16) is non-zero only in the cells BOTH boxes reach - the intersection of two
// cell rectangles, itself a rectangle.  Rows are sorted by that rectangle so the 256 rows of a CTA M tile share most of their
// cells; the tile's mask is the union.  key = ((y0*8 + y1)*8 + x0)*8 + x1 (inclusive cell bounds), pairs with no common cell last.
__global__ void pair_cell_keys_kernel(const int4* __restrict__ boxes, const int* __restrict__ pair_sub, const int* __restrict__ pair_obj,
                                      int n_pairs, int fs, int* __restrict__ keys) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  const unsigned long long m = cell_mask(rect_of(__ldg(boxes + pair_sub[p]), fs)) & cell_mask(rect_of(__ldg(boxes + pair_obj[p]), fs));
  int key = 4096;
  if (m) {
    const int lo = __ffsll((long long)m) - 1, hi = 63 - __clzll((long long)m);      // first / last set cell of a rectangle = its corners
    key = ((((lo >> 3) << 3 | (hi >> 3)) << 3 | (lo & 7)) << 3) | (hi & 7);
  }
  keys[p] = key;
}

// one warp per tile of `rows_per_tile` sorted rows: OR of the rows' cell masks
__global__ void tile_cell_masks_kernel(const int4* __restrict__ boxes, const int* __restrict__ row_sub, const int* __restrict__ row_obj,
                                       int n_rows, int fs, int rows_per_tile, int n_tiles, unsigned long long* __restrict__ masks) {
  const int tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (tile >= n_tiles) return;
  unsigned long long m = 0ull;
  const int r1 = min(n_rows, (tile + 1) * rows_per_tile);
  for (int r = tile * rows_per_tile + lane; r < r1; r += 32)
    m |= cell_mask(rect_of(__ldg(boxes + row_sub[r]), fs)) & cell_mask(rect_of(__ldg(boxes + row_obj[r]), fs));
#pragma unroll
  for (int d = 16; d; d >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, d);
  if (lane == 0) masks[tile] = m;
}

// zero every cell of a row that its tile's mask visits (the difference epilogue then overwrites the cells the pair itself computes);
// cells outside the mask are never read.  One CTA per row, a cell = `cell_vecs` uint4.
__global__ void __launch_bounds__(128)
cells_zero_kernel(const unsigned long long* __restrict__ masks, int rows_per_tile, long long n_rows, int n_cells, int cell_vecs,
                  uint4* __restrict__ out) {
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (long long r = blockIdx.x; r < n_rows; r += gridDim.x) {
    uint4* dst = out + r * (long long)n_cells * cell_vecs;
    for (unsigned long long m = __ldg(masks + r / rows_per_tile); m; m &= m - 1) {
      const int c = __ffsll((long long)m) - 1;
      for (int i = threadIdx.x; i < cell_vecs; i += blockDim.x) dst[(long long)c * cell_vecs + i] = z;
    }
  }
}

}  // namespace hc

using namespace hc;

static int conv3_blocks_list(const int32_t* boxes, const int32_t* pair_sub, const int32_t* pair_obj, int32_t n_pairs, int32_t feature_size,
                             int32_t block_rows, int32_t block_cols, bool both, int32_t* blocks, int32_t* n_blocks, hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(n_blocks && (n_pairs <= 0 || (boxes && pair_sub && pair_obj && blocks)), HC_E_NULL, "hc_conv3_active_blocks: NULL operand");
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  HC_REQUIRE(feature_size == 32, HC_E_SHAPE, "hc_conv3_active_blocks: built for feature_size 32 (8x8 pooled conv3 cells)");
  if (block_cols == 0) block_cols = 8;
  HC_REQUIRE(((block_rows == 8 || block_rows == 4) && (block_cols == 8 || (block_cols == 4 && block_rows == 4))) || (block_rows == 2 && block_cols == 4),
             HC_E_SHAPE, "hc_conv3_active_blocks: blocks are 8x8, 8x4, 4x4 or 4x2 pixels (block_cols x block_rows)");
  HC_REQUIRE(n_pairs >= 0 && n_pairs < (1 << 23), HC_E_SHAPE, "hc_conv3_active_blocks: n_pairs must be below 2^23");
  HC_REQUIRE(aligned16(boxes), HC_E_ALIGN, "hc_conv3_active_blocks: boxes must be 16-byte aligned");   // n_pairs == 0 still writes n_blocks = 0
  conv3_blocks_kernel<<<1, 1024, 0, stream>>>(reinterpret_cast<const int4*>(boxes), pair_sub, pair_obj, n_pairs, feature_size, block_rows / 2,
                                              block_cols / 2, both, blocks, n_blocks);
  return cuda_status("conv3_blocks_kernel launch");
}

extern "C" int hc_conv3_active_blocks(const int32_t* boxes, const int32_t* pair_sub, const int32_t* pair_obj, int32_t n_pairs,
                                      int32_t feature_size, int32_t block_rows, int32_t block_cols, int32_t* blocks, int32_t* n_blocks,
                                      hc_stream_t stream_) {
  return conv3_blocks_list(boxes, pair_sub, pair_obj, n_pairs, feature_size, block_rows, block_cols, false, blocks, n_blocks, stream_);
}

extern "C" int hc_conv3_shared_blocks(const int32_t* boxes, const int32_t* pair_sub, const int32_t* pair_obj, int32_t n_pairs,
                                      int32_t feature_size, int32_t block_rows, int32_t block_cols, int32_t* blocks, int32_t* n_blocks,
                                      hc_stream_t stream_) {
  return conv3_blocks_list(boxes, pair_sub, pair_obj, n_pairs, feature_size, block_rows, block_cols, true, blocks, n_blocks, stream_);
}

extern "C" int hc_p3_assemble(const void* background, const void* sub_maps, const void* obj_maps, const int32_t* boxes,
                              const int32_t* pair_sub, const int32_t* pair_obj, int32_t n_pairs, int32_t feature_size, void* out,
                              hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(n_pairs <= 0 || (background && sub_maps && obj_maps && boxes && pair_sub && pair_obj && out), HC_E_NULL,
             "hc_p3_assemble: NULL operand");
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  HC_REQUIRE(feature_size == 32, HC_E_SHAPE, "hc_p3_assemble: built for feature_size 32 (8x8 pooled conv3 cells of 1024 channels)");
  HC_REQUIRE(aligned16(background) && aligned16(sub_maps) && aligned16(obj_maps) && aligned16(boxes) && aligned16(out), HC_E_ALIGN,
             "hc_p3_assemble: operands must be 16-byte aligned");
  if (n_pairs <= 0) return HC_OK;
  const int grid = n_pairs < num_sms() * 8 ? n_pairs : num_sms() * 8;
  p3_assemble_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(background), reinterpret_cast<const uint4*>(sub_maps),
                                               reinterpret_cast<const uint4*>(obj_maps), reinterpret_cast<const int4*>(boxes), pair_sub,
                                               pair_obj, n_pairs, feature_size, reinterpret_cast<uint4*>(out));
  return cuda_status("p3_assemble_kernel launch");
}

extern "C" int hc_broadcast_rows(const void* src, int64_t row_bytes, int64_t n_rows, void* out, hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(src && out, HC_E_NULL, "hc_broadcast_rows: NULL operand");
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  HC_REQUIRE(row_bytes > 0 && row_bytes % 16 == 0 && n_rows >= 0, HC_E_SHAPE, "hc_broadcast_rows: row_bytes must be a positive multiple of 16");
  HC_REQUIRE(aligned16(src) && aligned16(out), HC_E_ALIGN, "hc_broadcast_rows: src/out must be 16-byte aligned");
  if (n_rows == 0) return HC_OK;
  const long long row_vecs = row_bytes / 16;
  const long long gx = (row_vecs + 255) / 256;
  HC_REQUIRE(gx < (1ll << 31), HC_E_SHAPE, "hc_broadcast_rows: row too long");
  long long gy = ((long long)num_sms() * 16 + gx - 1) / gx;
  if (gy > n_rows) gy = n_rows;
  if (gy > 65535) gy = 65535;
  broadcast_rows_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, stream>>>(reinterpret_cast<const uint4*>(src), row_vecs, n_rows,
                                                                            reinterpret_cast<uint4*>(out));
  return cuda_status("broadcast_rows_kernel launch");
}

extern "C" int hc_pair_cell_keys(const int32_t* boxes, const int32_t* pair_sub, const int32_t* pair_obj, int32_t n_pairs, int32_t feature_size,
                                 int32_t* keys, hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(n_pairs <= 0 || (boxes && pair_sub && pair_obj && keys), HC_E_NULL, "hc_pair_cell_keys: NULL operand");
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  HC_REQUIRE(feature_size == 32, HC_E_SHAPE, "hc_pair_cell_keys: built for feature_size 32 (8x8 pooled conv3 cells)");
  HC_REQUIRE(aligned16(boxes), HC_E_ALIGN, "hc_pair_cell_keys: boxes must be 16-byte aligned");
  if (n_pairs <= 0) return HC_OK;
  pair_cell_keys_kernel<<<(n_pairs + 255) / 256, 256, 0, stream>>>(reinterpret_cast<const int4*>(boxes), pair_sub, pair_obj, n_pairs,
                                                                   feature_size, keys);
  return cuda_status("pair_cell_keys_kernel launch");
}

extern "C" int hc_tile_cell_masks(const int32_t* boxes, const int32_t* row_sub, const int32_t* row_obj, int32_t n_rows, int32_t feature_size,
                                  int32_t rows_per_tile, uint64_t* masks, hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(n_rows <= 0 || (boxes && row_sub && row_obj && masks), HC_E_NULL, "hc_tile_cell_masks: NULL operand");
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  HC_REQUIRE(feature_size == 32, HC_E_SHAPE, "hc_tile_cell_masks: built for feature_size 32 (8x8 pooled conv3 cells)");
  HC_REQUIRE(rows_per_tile > 0, HC_E_SHAPE, "hc_tile_cell_masks: rows_per_tile must be positive");
  HC_REQUIRE(aligned16(boxes), HC_E_ALIGN, "hc_tile_cell_masks: boxes must be 16-byte aligned");
  if (n_rows <= 0) return HC_OK;
  const int n_tiles = (n_rows + rows_per_tile - 1) / rows_per_tile;
  tile_cell_masks_kernel<<<(n_tiles + 7) / 8, 256, 0, stream>>>(reinterpret_cast<const int4*>(boxes), row_sub, row_obj, n_rows, feature_size,
                                                                rows_per_tile, n_tiles, reinterpret_cast<unsigned long long*>(masks));
  return cuda_status("tile_cell_masks_kernel launch");
}

extern "C" int hc_cells_zero(const uint64_t* masks, int32_t rows_per_tile, int64_t n_rows, int32_t n_cells, int64_t cell_bytes, void* out,
                             hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(n_rows <= 0 || (masks && out), HC_E_NULL, "hc_cells_zero: NULL operand");
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  HC_REQUIRE(rows_per_tile > 0 && n_cells > 0 && n_cells <= 64 && cell_bytes > 0 && cell_bytes % 16 == 0, HC_E_SHAPE,
             "hc_cells_zero: 1..64 cells of a positive multiple of 16 bytes");
  HC_REQUIRE(aligned16(out), HC_E_ALIGN, "hc_cells_zero: out must be 16-byte aligned");
  if (n_rows <= 0) return HC_OK;
  const long long cap = (long long)num_sms() * 16;
  cells_zero_kernel<<<(unsigned)(n_rows < cap ? n_rows : cap), 128, 0, stream>>>(reinterpret_cast<const unsigned long long*>(masks),
                                                                                rows_per_tile, n_rows, n_cells, (int)(cell_bytes / 16),
                                                                                reinterpret_cast<uint4*>(out));
  return cuda_status("cells_zero_kernel launch");
}

extern "C" int hc_pair_cover_masks(const int32_t* boxes, const int32_t* pair_sub, const int32_t* pair_obj, int32_t n_pairs, int32_t feature_size,
                                   int32_t block_rows, int32_t block_cols, int32_t shared, uint64_t* masks, hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(n_pairs <= 0 || (boxes && pair_sub && pair_obj && masks), HC_E_NULL, "hc_pair_cover_masks: NULL operand");
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  HC_REQUIRE(feature_size == 32, HC_E_SHAPE, "hc_pair_cover_masks: built for feature_size 32 (8x8 pooled conv3 cells)");
  if (block_cols == 0) block_cols = 8;
  HC_REQUIRE(((block_rows == 8 || block_rows == 4) && (block_cols == 8 || (block_cols == 4 && block_rows == 4))) || (block_rows == 2 && block_cols == 4),
             HC_E_SHAPE, "hc_pair_cover_masks: blocks are 8x8, 8x4, 4x4 or 4x2 pixels (block_cols x block_rows)");
  HC_REQUIRE(aligned16(boxes), HC_E_ALIGN, "hc_pair_cover_masks: boxes must be 16-byte aligned");
  if (n_pairs <= 0) return HC_OK;
  pair_cover_masks_kernel<<<(n_pairs + 255) / 256, 256, 0, stream>>>(reinterpret_cast<const int4*>(boxes), pair_sub, pair_obj, n_pairs,
                                                                     feature_size, block_rows / 2, block_cols / 2, shared != 0,
                                                                     reinterpret_cast<unsigned long long*>(masks));
  return cuda_status("pair_cover_masks_kernel launch");
}

extern "C" int hc_conv2_box_blocks(const int32_t* boxes, int32_t n_box, int32_t feature_size, int32_t block_rows, int32_t* blocks,
                                   int32_t* n_blocks, hc_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  HC_REQUIRE(n_blocks && (n_box <= 0 || (boxes && blocks)), HC_E_NULL, "hc_conv2_box_blocks: NULL operand");
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  HC_REQUIRE(feature_size == 32, HC_E_SHAPE, "hc_conv2_box_blocks: built for feature_size 32 (block origins are packed in 4 bits of even pixels)");
  HC_REQUIRE(block_rows == 8 || block_rows == 4, HC_E_SHAPE, "hc_conv2_box_blocks: block_rows must be 8 or 4");
  HC_REQUIRE(n_box >= 0 && n_box < (1 << 23), HC_E_SHAPE, "hc_conv2_box_blocks: n_box must be below 2^23");
  HC_REQUIRE(aligned16(boxes), HC_E_ALIGN, "hc_conv2_box_blocks: boxes must be 16-byte aligned");
  conv2_blocks_kernel<<<1, 1024, 0, stream>>>(reinterpret_cast<const int4*>(boxes), n_box, feature_size, block_rows, blocks, n_blocks);
  return cuda_status("conv2_blocks_kernel launch");
}
