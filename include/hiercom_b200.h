/* hiercom_b200 - C ABI of the B200-native HIERCOM relation-prediction hot path.
 *
 * Drop-in boundary: plain pointers + sizes + a CUDA stream in, status code out.  Every pointer
 * named "device" is a device pointer owned by the caller (the PyTorch caching allocator in the Python host);
 * kernels never allocate.  All entry points are asynchronous on `stream` unless noted, re-entrant, and return
 * HC_OK or a negative HC_E_* code; `hc_last_error()` returns the text for the calling thread.  There is no CPU
 * fallback: on a device that is not sm_100 every compute entry point returns HC_E_ARCH.
 *
 * Citations are into the reference repository (bowen-upenn/scene_graph_commonsense @ 3388036f).
 */
#ifndef HIERCOM_B200_H_
#define HIERCOM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HC_ABI_VERSION 5

#define HC_OK 0
#define HC_E_SHAPE (-1) /* bad size / unsupported shape          */
#define HC_E_ALIGN (-2) /* pointer or stride not 16-byte aligned */
#define HC_E_ARCH (-3)  /* current device is not sm_100          */
#define HC_E_CUDA (-4)  /* CUDA runtime / driver error           */
#define HC_E_NULL (-5)  /* required pointer is NULL              */

#define HC_NUM_OBJ 150
#define HC_NUM_PRED_MAX 64
#define HC_TRIPLET_SPACE (150 * 50 * 150)
#define HC_BITMAP_WORDS ((HC_TRIPLET_SPACE + 31) / 32)
#define HC_TOP_MAX 128 /* largest supported top-K cut (reference uses 100) */

typedef void* hc_stream_t; /* cudaStream_t */

const char* hc_last_error(void);
int hc_abi_version(void);
/* HC_OK iff the current CUDA device is compute capability 10.x */
int hc_device_check(void);

/* ---------------------------------------------------------------------------------------------------------
 * R9 (host) - commonsense sets -> one "passes the filter" bitmap: bit(key) = key in aligned AND key not in
 * violated, key = (s*50+p)*150+o.  Replaces the two python dict lookups per candidate of
 * evaluator.py:189-194,261-266 (and train_utils.py:53-54).  HOST function, synchronous.
 */
int hc_cs_bitmap_build(const int64_t* aligned_keys, int64_t n_aligned, const int64_t* violated_keys,
                       int64_t n_violated, uint32_t* bitmap_out /* host, HC_BITMAP_WORDS words */);

/* ---------------------------------------------------------------------------------------------------------
 * R1 + R2 + R4 - rectangular masks, two-pass ordered-pair enumeration and the overlap pre-filter
 * (evaluate.py:111-116,132-156; train_test.py:365-410).  Images are CSR segments of `boxes`
 * (int32 xmin,xmax,ymin,ymax on the feature grid, already truncated toward zero like `int()`).
 *
 * For image i with N_i boxes, unordered pairs (g,e), e<g are visited in the reference's loop order
 * t = g(g-1)/2 + e.  ov(i,g,e) = masks overlap.  Skip rule:
 *   group_id == NULL ("per_image", = reference at batch_size 1): keep iff ov(i,g,e)
 *   group_id != NULL ("batch"): keep iff ANY image j of the same group with N_j > g has ov(j,g,e)
 * Each kept pair emits two directed pairs: (sub=g,obj=e) then (sub=e,obj=g), with
 *   pair_ov  = ov(i,g,e)                                       (iou_mask, evaluate.py:154)
 *   pair_gt  = rel_tri[t] if dir_tri[t] == 1 (first) / == 0 (second) else -1   (train_utils.py:169-187)
 *   pair_rel = rel_tri[t] (undirected label, optional output, used by hc_connectivity_stats)
 * Outputs must be sized for sum_i N_i(N_i-1) entries; `total_out[0]` receives the number written.
 * Workspaces: ws_ov [sum T_i] bytes, ws_any [n_groups*max_tri] bytes (batch mode), ws_counts [n_images].
 */
int hc_pairs_enumerate(const int32_t* boxes, const int32_t* box_offsets, int32_t n_images,
                       const int32_t* group_id, int32_t n_groups, int32_t max_tri, const int32_t* rel_tri,
                       const int8_t* dir_tri, const int32_t* tri_offsets, int32_t feature_size, uint8_t* ws_ov,
                       uint8_t* ws_any, int32_t* ws_counts, int32_t* pair_offsets, int32_t* pair_sub,
                       int32_t* pair_obj, int32_t* pair_img, uint8_t* pair_ov, int32_t* pair_gt,
                       int32_t* pair_rel, int32_t* total_out, hc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * R3 (fused away) + R5/R6 dense contractions - one tcgen05/TMEM kernel, bf16 operands, fp32 accumulate.
 *
 *   out[m, n] = epilogue( sum_k A[m,k] * B[n,k] )        B is [N,K] row-major ("K-major"), bf16
 *
 * mode HC_GEMM_PLAIN : A is [M,K] row-major bf16 with row stride lda (elements).
 * mode HC_GEMM_CONV3 : implicit 3x3 / pad 1 / stride 1 convolution.  A is an NHWC activation tensor
 *     [n_img, H, W, c_total] bf16; input channels [c_base, c_base+c_in) are used; M = n_img*H*W output pixels,
 *     K = 9*c_in with k = (ky*3+kx)*c_in + c (B must be packed in that order).  H, W multiples of 16.
 * mode HC_GEMM_CONV3_BLOCKS : the same convolution evaluated only on a device-side WORK LIST of 8-pixel-wide, `block_rows`-tall
 *     (8 or 4) output blocks: `blocks[i] = img << 8 | (y0/2) << 4 | (x0/2)` (even origins), `n_blocks[0]` entries (read on the
 *     device: no host round trip).  Pooled epilogues or HC_EPI_BF16; output pixels outside the listed blocks are NOT written (the caller
 *     pre-fills them, see hc_conv3_active_blocks / hc_broadcast_rows).  Same K order as HC_GEMM_CONV3: listed pixels are
 *     bit-identical to the dense result.
 * epilogue HC_EPI_BF16      : out bf16 [M, ldc] (+col offset c_off): act(acc + bias)
 *          HC_EPI_F32       : out f32  [M, ldc]: acc (+ bias if non-NULL)
 *          HC_EPI_POOL_BF16 : conv only: 2x2/stride-2 max-pool of relu(acc + bias) -> NHWC bf16
 *                             [n_img, H/2, W/2, ldc] (model.py:143-146 conv+ReLU+maxpool)
 *          HC_EPI_SPLIT3_BF16 : plain only: x = act(acc + bias) * mul written as the bf16x3 A operand of a following
 *                             GEMM, out bf16 [M, ldc >= 3N] = [hi | lo | hi], hi = bf16(x), lo = bf16(x - hi)
 *                             (see hc_split_bf16x3); the f32 intermediate never reaches HBM
 *          HC_EPI_POOL_DIFF_BF16 : CONV3_BLOCKS only: with x = the HC_EPI_POOL_BF16 value (rounded to bf16) of local pair i,
 *                             writes d = (x - diff_sub[pair_sub[i]]) - (diff_obj[pair_obj[i]] - diff_bg) (fp32, then bf16) into row
 *                             pair_row[i] of out [rows, H/2, W/2, ldc]; diff_* are pooled maps of the same layout.  d is exactly 0
 *                             in a cell that only one box of the pair (or none) reaches - the operand of the K-cell-sparse fc1 below.
 * K-cell-sparse PLAIN GEMM (k_masks != NULL): the K axis is cut into K / k_cell <= 64 cells; bit c of k_masks[t] says that some row
 *     of CTA M tile t (m_sub*128 rows) is non-zero in cell c, and only those cells' K blocks are visited, in ascending order.  The
 *     caller guarantees A is ZERO in every (row, cell) of a visited cell the row does not use, so the result equals the dense GEMM
 *     bit for bit per row, whatever tile the row sits in.  A tile with an empty mask skips the tensor core altogether.
 *     (model.py:149 fc1 is linear: fc1(p3) = fc1(sub map) + fc1(obj map) - fc1(background) + fc1(d), and d is zero outside the cells
 *     both boxes reach - hc_pair_cell_keys / hc_tile_cell_masks below build the row order and the masks.)
 * add_a/add_b (EPI_BF16, PLAIN): act(acc + bias + add_a[add_a_rows[r]] + add_b[add_b_rows[r]]), f32 tables [*, ld_add].
 * out_rows (PLAIN, no mul): GEMM row r is written to output row out_rows[r] (a permutation; NULL = identity).
 * Requirements: K % 64 == 0, N % 128 == 0, all bases 16-byte aligned, lda/ldc/c_total multiples of 8.
 */
#define HC_GEMM_PLAIN 0
#define HC_GEMM_CONV3 1
#define HC_GEMM_CONV3_BLOCKS 2
#define HC_EPI_BF16 0
#define HC_EPI_F32 1
#define HC_EPI_POOL_BF16 2
#define HC_EPI_SPLIT3_BF16 3
#define HC_EPI_POOL_DIFF_BF16 4
#define HC_ACT_NONE 0
#define HC_ACT_RELU 1
#define HC_ACT_TANH 2

typedef struct hc_gemm_desc {
  const void* a;     /* device, bf16 */
  const void* b;     /* device, bf16 [N,K] */
  const float* bias; /* device, f32 [N] or NULL */
  void* out;         /* device */
  int64_t m, n, k;
  int64_t lda;       /* PLAIN: A row stride in elements */
  int64_t ldc;       /* output row stride in elements */
  int64_t c_off;     /* output column offset in elements */
  int32_t mode, epilogue, act;
  int32_t n_img, h, w, c_total, c_base, c_in; /* CONV3 */
  int32_t group_m;   /* tile rasterisation: m-blocks per band (0 = default) */
  int32_t m_sub;     /* 128-row sub-tiles per CTA tile: 1 or 2 (0 = default) */
  const float* mul;  /* optional f32 [M, ld_mul]: out = act(acc + bias) * mul (PLAIN, non-pooled); SGB `* union_features` */
  int64_t ld_mul;
  const int32_t* blocks;   /* CONV3_BLOCKS: device work list */
  const int32_t* n_blocks; /* CONV3_BLOCKS: device scalar, number of entries */
  int32_t block_rows;      /* CONV3_BLOCKS: 8 or 4; 2 (blocks one pooled cell tall) with block_cols 4 and cta_pairs */
  int32_t block_cols;      /* CONV3_BLOCKS: 8 (or 0), or 4 with block_rows 4 or 2 */
  int32_t cta_pairs;       /* 1 = tcgen05 cta_group::2, clusters of 2 CTAs, UMMA M = 256 across the pair (same K order: bit-identical results).
                              CONV3_BLOCKS (4-pixel-wide blocks, pooled epilogue): weights on the M side, the tile's 256 pixels shared by
                              the pair, each CTA stages half of them.  PLAIN (N % 256 == 0, m_sub 1): a pair owns one 256 x 256 tile, each
                              CTA staging its 128 rows of A and 128 columns of B per K step; k_masks stays per 256-row tile */
  const uint64_t* k_masks; /* PLAIN: per CTA M tile, bitmap of visited K cells (NULL = dense) */
  int64_t k_cell;          /* PLAIN + k_masks: K elements per cell (multiple of 64, K / k_cell <= 64) */
  const float* add_a;      /* EPI_BF16 row gathers, both or neither */
  const int32_t* add_a_rows;
  const float* add_b;
  const int32_t* add_b_rows;
  int64_t ld_add;
  const int32_t* out_rows; /* PLAIN: output row per GEMM row (NULL = identity) */
  const void* diff_sub;    /* HC_EPI_POOL_DIFF_BF16: bf16 [n_sub, H/2, W/2, ldc] */
  const void* diff_obj;    /*                        bf16 [n_obj, H/2, W/2, ldc] */
  const void* diff_bg;     /*                        bf16 [H/2, W/2, ldc] */
  const int32_t* pair_sub; /*                        per local pair: row of diff_sub */
  const int32_t* pair_obj; /*                        per local pair: row of diff_obj */
  const int32_t* pair_row; /*                        per local pair: output row */
  void* scratch;           /* cta_pairs + HC_EPI_POOL_DIFF_BF16: bf16 [n_img, H/2, W/2, ldc] work map (the pair kernel leaves the pooled
                              values there by local pair, a second launch forms the differences); NULL = single-CTA kernel */
  const int32_t* m_order;  /* PLAIN: visiting order of the CTA M tiles (m_order[i] = tile visited i-th; a permutation of 0 .. tiles-1, tile =
                              m_sub * 128 rows, with cta_pairs m_sub * 256 rows); NULL = ascending.  Longest-first for K-cell-sparse launches */
  int32_t operand_f16;     /* 16-bit operand format: 0 = bf16 (default), 1 = IEEE fp16 - A, B, the 16-bit outputs ("BF16" epilogues
                              then write fp16, saturating at +-65504) and the difference maps.  Same tensor-core rate
                              (tcgen05 kind::f16); 3 more mantissa bits = 8x smaller operand rounding error.  Not with SPLIT3. */
} hc_gemm_desc;

int hc_tc_gemm(const hc_gemm_desc* desc, hc_stream_t stream);

/* R3/R5 - which part of conv3_1's output a directed pair actually has to compute.
 * `feature*mask` (train_test.py:391,398) leaves tanh(conv1_x.bias) outside a box, so every activation after it equals a
 * weights-only BACKGROUND wherever the receptive field misses both boxes (model.py:139-146: 3x3 conv, 2x2 pool, 3x3 conv, 2x2
 * pool).  For a box [lo,hi) on the 32-grid the pooled conv3_1 output (8-grid "cells") can differ from the background only in
 * cells [ (max(0,(lo-1)>>1) - 1 clamped) >> 1 , (min(15, min(15, hi>>1) + 1)) >> 1 ] per axis; a pair's active set is the union of
 * its two boxes' cell rectangles.  This entry point covers that set greedily (first uncovered cell in row-major order, block
 * origin clamped into the map) with blocks of 4 x (block_rows/2) cells = 8 x block_rows conv3 pixels and writes the work list
 * HC_GEMM_CONV3_BLOCKS consumes: blocks[i] = local_pair << 8 | cell_y << 4 | cell_x, pairs in order, n_blocks[0] = count.
 * block_cols (0 = 8) is the block width in conv3 pixels: blocks are 8x8, 8x4 or 4x4 pixels (block_cols x block_rows) = 4x4, 4x2 or
 * 2x2 cells; smaller blocks hug the cell rectangle more tightly (fewer computed pixels, more TMA boxes per tile).
 * `blocks` must hold n_pairs * 256 / (block_rows*block_cols) entries.  feature_size must be 32.  One CTA, deterministic order. */
int hc_conv3_active_blocks(const int32_t* boxes, const int32_t* pair_sub, const int32_t* pair_obj, int32_t n_pairs,
                           int32_t feature_size, int32_t block_rows, int32_t block_cols, int32_t* blocks, int32_t* n_blocks,
                           hc_stream_t stream);

/* Shared-footprint variant (the default): a pooled conv3_1 cell that only the SUBJECT's box reaches equals the same cell of the
 * pair (subject, empty box), one only the OBJECT's box reaches equals (empty box, object) - both are per-BOX maps, computed once
 * per box of the window instead of once per pair (model.py:143-146 see the object only through its masked features,
 * train_test.py:391,398).  Only the cells BOTH boxes reach depend on the pair: hc_conv3_shared_blocks lists the cover of that
 * intersection (same cover, same entry format and capacity as hc_conv3_active_blocks). */
int hc_conv3_shared_blocks(const int32_t* boxes, const int32_t* pair_sub, const int32_t* pair_obj, int32_t n_pairs,
                           int32_t feature_size, int32_t block_rows, int32_t block_cols, int32_t* blocks, int32_t* n_blocks,
                           hc_stream_t stream);

/* out[p] ([8,8,1024] bf16 per pair) = per cell: sub_maps[pair_sub[p]] where only the subject's box reaches the cell,
 * obj_maps[pair_obj[p]] where only the object's box does, `background` where neither does; cells both reach are NOT written
 * (HC_GEMM_CONV3_BLOCKS over hc_conv3_shared_blocks' list writes them afterwards on the same stream).  sub_maps / obj_maps:
 * [n_box,8,8,1024] bf16 = pooled conv3_1 output of (box, empty) / (empty, box); background: [8,8,1024] bf16. */
int hc_p3_assemble(const void* background, const void* sub_maps, const void* obj_maps, const int32_t* boxes,
                   const int32_t* pair_sub, const int32_t* pair_obj, int32_t n_pairs, int32_t feature_size, void* out,
                   hc_stream_t stream);

/* out[i, :] = src[:] for i < n_rows (row_bytes a multiple of 16; both 16-byte aligned): pre-fills the pooled conv3_1 output of
 * every pair with the background before HC_GEMM_CONV3_BLOCKS overwrites the active blocks. */
int hc_broadcast_rows(const void* src, int64_t row_bytes, int64_t n_rows, void* out, hc_stream_t stream);

/* model.py:143 on the box footprint: conv2_1's per-box halves (see hc_tc_gemm) differ from the weights-only background
 * map only within one pixel of the box rectangle.  Lists 8 x block_rows-pixel blocks (even origins, clamped into the 32 x 32 map)
 * covering that rectangle for every box: blocks[i] = box << 8 | (y0/2) << 4 | (x0/2), n_blocks[0] = count (device side); `blocks`
 * must hold n_box * 4 * 32 / block_rows entries.  HC_GEMM_CONV3_BLOCKS with HC_EPI_BF16 then writes the listed pixels of an output
 * the caller pre-filled with the background (hc_broadcast_rows); listed pixels are bit-identical to the dense convolution. */
int hc_conv2_box_blocks(const int32_t* boxes, int32_t n_box, int32_t feature_size, int32_t block_rows, int32_t* blocks,
                        int32_t* n_blocks, hc_stream_t stream);

/* model.py:149 - shared-footprint fc1 (the K-cell-sparse mode of hc_tc_gemm).  fc1 is linear, and the pooled conv3_1 output of a
 * pair differs from "subject map + object map - background" (hc_p3_assemble) only in the 8-grid cells BOTH boxes reach, a
 * rectangle of cells.  hc_pair_cell_keys writes a sort key of that rectangle per directed pair, ((y0*8 + y1)*8 + x0)*8 + x1 with
 * inclusive cell bounds, 4096 when the boxes share no cell: sorting rows by it makes the rows of a GEMM tile share their cells.
 * hc_tile_cell_masks ORs the cell masks of every `rows_per_tile` consecutive sorted rows (row_sub / row_obj = box ids in sorted
 * order) into masks[tile] - the `k_masks` of hc_tc_gemm (rows_per_tile = m_sub*128).  hc_cells_zero zeroes, in every row of the
 * operand out [n_rows, n_cells, cell_bytes], the cells its tile's mask visits; HC_EPI_POOL_DIFF_BF16 then overwrites the cells the
 * pair computes.  feature_size must be 32. */
int hc_pair_cell_keys(const int32_t* boxes, const int32_t* pair_sub, const int32_t* pair_obj, int32_t n_pairs, int32_t feature_size,
                      int32_t* keys, hc_stream_t stream);
int hc_tile_cell_masks(const int32_t* boxes, const int32_t* row_sub, const int32_t* row_obj, int32_t n_rows, int32_t feature_size,
                       int32_t rows_per_tile, uint64_t* masks, hc_stream_t stream);
int hc_cells_zero(const uint64_t* masks, int32_t rows_per_tile, int64_t n_rows, int32_t n_cells, int64_t cell_bytes, void* out,
                  hc_stream_t stream);

/* [B,C0,hw] f32 (+ optional [B,C1,hw] f32) NCHW maps -> [B*hw, k_pad] bf16 pixel-major rows, zero padded
 * (the A operand of the 1x1 convolutions, model.py:139-140; also packs the legacy pre-masked [bs,257,32,32]). */
int hc_pack_pixels(const float* src0, int32_t c0, const float* src1, int32_t c1, int32_t n_img, int32_t hw,
                   int32_t k_pad, void* out_bf16, int32_t operand_f16, hc_stream_t stream);

/* R3: per-box masked conv1 activations.  tanh(conv1(x*mask)) == mask ? tanh(conv1(x)) : tanh(bias)
 * (train_test.py:391,398 + model.py:139-140).  t_img [n_img, fs*fs, C] bf16,
 * fill [C] bf16 = tanh(bias), out [n_box, fs, fs, C] bf16. */
int hc_box_select(const void* t_img, const int32_t* boxes, const int32_t* box_img, int32_t n_box, int32_t fs,
                  int32_t channels, const void* fill, void* out, hc_stream_t stream);

/* model.py:143-144 after the subject/object split of conv2_1:
 * out[p] = maxpool2x2(relu(U[pair_sub[p]] + V[pair_obj[p]] + bias)),  U,V [n_box, fs, fs, C] bf16,
 * out [n_pairs, fs/2, fs/2, C] bf16.  bias == NULL: V already includes the conv2 bias (added in fp32 in the object-half
 * GEMM epilogue); the kernel then runs on packed bf16x2 adds/max - bit-identical to rounding the fp32 sum, 1/3 of the
 * instructions, HBM-bound instead of issue-bound. */
/* hc_uv_footprint (optional, NULL = u / v are complete maps): u / v were written by HC_GEMM_CONV3_BLOCKS over hc_conv2_box_blocks' list
 * WITHOUT a background pre-fill, i.e. a box's map is defined only inside the rectangle its listed blocks cover; outside it the value
 * is taken from the background maps u_bg / v_bg [fs, fs, channels] (the conv2_1 halves of an all-tanh(bias) map), which is what the
 * complete map holds there bit for bit.  Saves writing 2 MiB of background per box. */
typedef struct hc_uv_footprint {
  const int32_t* boxes;  /* [n_box,4] the boxes u / v were computed for (same rows) */
  const void* u_bg;
  const void* v_bg;
  int32_t block_rows;    /* of the hc_conv2_box_blocks list: 4 or 8 */
} hc_uv_footprint;
int hc_pair_relu_pool(const void* u, const void* v, const float* bias, const int32_t* pair_sub,
                      const int32_t* pair_obj, int32_t n_pairs, int32_t fs, int32_t channels, const uint64_t* cover,
                      const hc_uv_footprint* fp, void* out, int32_t operand_f16,
                      hc_stream_t stream);     /* cover: as hc_pair_relu_pool_tiled (bias == NULL), or NULL */

/* Same stage for pair lists produced by hc_pairs_enumerate, tiled as an outer sum over the boxes of an image: a
 * thread block keeps the U tiles of 4 subject boxes in registers and streams every object box's V tile once, so
 * reads per pair drop from 2 MiB to ~0.3 MiB.  lut [n_box, n_max] int32 maps (subject box, local object index) to
 * the directed pair index (-1 = pair skipped); built by hc_pair_lut_build.  Processes images [img0, img0+n_img)
 * and writes out[p - pair_base] for pair_base <= p < pair_base + chunk_pairs. */
int hc_pair_lut_build(const int32_t* pair_sub, const int32_t* pair_obj, const int32_t* pair_img,
                      const int32_t* box_offsets, int32_t n_pairs, int32_t n_box, int32_t n_max, int32_t* lut,
                      hc_stream_t stream);
int hc_pair_relu_pool_tiled(const void* u, const void* v, const float* bias, const int32_t* box_offsets,
                            const int32_t* lut, int32_t n_max, int32_t img0, int32_t n_img, int32_t pair_base,
                            int32_t chunk_pairs, int32_t fs, int32_t channels, const uint64_t* cover, const hc_uv_footprint* fp,
                            void* out, int32_t operand_f16, hc_stream_t stream);

/* Footprint-aware pooling: masks[p] = the 8x8-grid cells covered by the blocks hc_conv3_active_blocks (shared = 0) or
 * hc_conv3_shared_blocks (shared = 1) lists for pair p (same cover function).  Passed as `cover` (chunk-local, bias == NULL,
 * feature_size 32) to hc_pair_relu_pool_tiled, a pooled pixel of a pair is written only if a listed conv3_1 block reads it (block +
 * 1-pixel halo); HC_GEMM_CONV3_BLOCKS never reads the others.  cover == NULL writes every pixel. */
int hc_pair_cover_masks(const int32_t* boxes, const int32_t* pair_sub, const int32_t* pair_obj, int32_t n_pairs, int32_t feature_size,
                        int32_t block_rows, int32_t block_cols, int32_t shared, uint64_t* masks, hc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * R6 tail + R7 - label-embedding add, fc2 bias + ReLU, fc3_x / fc4 / fc5 heads and the Bayesian hierarchical
 * log-softmax (model.py:152-168,175-184; flat variant model.py:97-101).
 *   pred = relu(fc2_raw + fc2_bias + E[c_sub] + E[150+c_obj] + sum' E[300+s_sub] + sum' E[317+s_obj])
 *   sum' = utils.py:136-149 `process_super_class`: the FIRST entry of the box's super-class list plus, for a list of
 *   2..4 entries, its LAST entry (the middle entries of 3- and 4-entry lists are never added by the reference).
 *   box_super holds the raw lists (left-packed, -1 padded); the kernels apply the rule.
 *   hier : super = log_softmax(fc5 pred); rel_k = log_softmax(fc3_k pred / T_k) + super[k]; conn = fc4 pred
 *   flat : relation = fc3 pred (raw logits), conn = fc4 pred
 * emb is fc2.weight[:, 4096:].T, f32 [n_emb, hidden].  w_heads rows: hier [fc3_1; fc3_2; fc3_3; fc4; fc5],
 * flat [fc3; fc4].  row_sub/row_obj index box_cat / box_super ([n_box,4] int8, -1 padded; NULL = no
 * super-class columns, model.py:125-128).  logsig = log(sigmoid(conn)) (train_utils.py:190).
 * fc2_bias == NULL: fc2_raw already is the hidden vector - no bias/embedding/ReLU (BayesianHead, model.py:24-34).
 * box_emb (optional) f32 [n_box, 2*hidden] from hc_box_label_embed: the label columns summed once per box
 * ([as subject | as object]); when given, a row adds box_emb[row_sub][0:hidden] + box_emb[row_obj][hidden:] instead of
 * gathering up to ten embedding rows (same sum, different fp32 association).
 */
int hc_hier_head(const float* fc2_raw, int64_t ld_raw, int32_t n_rows, int32_t hidden, const float* fc2_bias,
                 const float* emb, int32_t num_obj, int32_t num_super, const int32_t* row_sub,
                 const int32_t* row_obj, const int32_t* box_cat, const int8_t* box_super, const float* w_heads,
                 const float* b_heads, int32_t n_geo, int32_t n_pos, int32_t n_sem, int32_t flat, float t1,
                 float t2, float t3, float* relation, float* super_rel, float* connectivity, float* logsig,
                 float* pred_out, const float* box_emb, hc_stream_t stream);
int hc_box_label_embed(const float* emb, int32_t num_obj, int32_t num_super, const int32_t* box_cat,
                       const int8_t* box_super, int32_t n_box, int32_t hidden, float* out, hc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * R8 + R9 (+ R10's connectivity add) - candidate construction (evaluator.py:124-138,157-179,231-266,646-649).
 * Per row r and super-category k: conf = max(relation[r, seg_k]) (+ conf_sub[r] + conf_obj[r] when non-NULL),
 * label = first argmax + offset_k; conf = -inf if !row_ov[r]; conf = -inf if pass_bitmap != NULL and
 * (cat_sub,label,cat_obj) does not pass; finally conf += logsig[r] (evaluator.py:292).
 * cand arrays hold K = (hier ? 3 : 1) entries per row, index r*K + k (layout 0) or k*n_rows + r (layout 1,
 * the reference's per-call append order).  t3_conf[r] = max_k conf_k with only the overlap mask
 * (Evaluator_Top3, no commonsense filter) + logsig; t3_super[r] = argmax(super_rel[r]).
 */
int hc_candidates(const float* relation, int64_t ld_rel, int32_t n_rows, int32_t n_geo, int32_t n_pos,
                  int32_t n_sem, int32_t hier, const uint8_t* row_ov, const float* logsig, const float* conf_sub,
                  const float* conf_obj, const int32_t* row_sub, const int32_t* row_obj, const int32_t* box_cat,
                  const uint32_t* pass_bitmap, const float* super_rel, float* cand_conf, int32_t* cand_label,
                  float* t3_conf, uint8_t* t3_super, int32_t layout, hc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * R10-R13 - per-image top-K selection under (confidence desc, candidate index asc), first-match scan of
 * every GT triplet, integer hit / GT counters (evaluator.py:294-356 and :704-766).
 * Image i owns candidates [cand_offsets[i], cand_offsets[i+1]) and GT slots [gt_offsets[i], gt_offsets[i+1]).
 * Candidate c belongs to row cand_row[c] (NULL: c / K); a row's subject/object are entries row_sub/row_obj of
 * the pred tables (pred_cat, pred_box[,4] = int xmin,xmax,ymin,ymax); GT slot g references gt tables the
 * same way; gt_label == -1 slots are skipped.  synonyms NULL: exact category equality (PredCLS), else
 * utils.compare_object_cat truth table [num_obj*num_obj].  zs_bitmap: zero-shot triplet set or NULL.
 * mode 0: Evaluator counters (layout tables.EV_*, 408 slots), mode 1: Evaluator_Top3 (357 slots):
 * one candidate per row (k_per_row == 1, cand_label unused), t3_labels [n_rows*3] holds the three per-head
 * labels of each row (= hc_candidates' cand_label in layout 0) and t3_super its argmax super-category.
 * topk_out (optional) [n_images, top_max] receives the selected candidate ids in rank order, -1 padded.
 */
int hc_topk_match(const int32_t* cand_offsets, int32_t n_images, const float* cand_conf,
                  const int32_t* cand_label, const int32_t* cand_row, int32_t k_per_row, const int32_t* row_sub,
                  const int32_t* row_obj, const int32_t* pred_cat, const int32_t* pred_box,
                  const int32_t* gt_offsets, const int32_t* gt_label, const int32_t* gt_sub,
                  const int32_t* gt_obj, const int32_t* gt_cat, const int32_t* gt_box, const uint8_t* synonyms,
                  int32_t num_obj, int32_t num_pred, const uint32_t* zs_bitmap, int32_t feature_size,
                  double iou_thresh, int32_t top_max, int32_t k0, int32_t k1, int32_t k2, int32_t mode,
                  const int32_t* t3_labels, const uint8_t* t3_super, unsigned long long* counters,
                  int32_t* topk_out, hc_stream_t stream);

/* Selection only: topk_out [n_images, top_max] = image-local candidate ids ordered by (confidence desc, index asc),
 * -1 padded (the prefix of `torch.sort(descending=True, stable=True)`; evaluator.py:304, inference.py:282). */
int hc_topk_select(const int32_t* cand_offsets, int32_t n_images, const float* cand_conf, int32_t top_max,
                   int32_t* topk_out, hc_stream_t stream);

/* train_utils.py:169-183 side statistics over directed rows: stats[0..4] += num_not_connected, num_connected,
 * num_connected_pred (sigmoid(conn) >= 0.5), connectivity_precision (#pred-connected rows whose undirected GT
 * label != -1), connectivity_recall (sum round(sigmoid(conn)) over connected rows). */
int hc_connectivity_stats(const float* connectivity, const int32_t* row_gt_directed,
                          const int32_t* row_gt_undirected, int32_t n_rows, unsigned long long* stats,
                          hc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * R14 / N1 - Scene-Graph-Benchmark plug-and-play twin (PredCLS, config 5).  Dense stages use hc_tc_gemm:
 * post_cat (1024 -> 4096) with `mul = union_features` (roi_relation_predictors.py:421-427) and the BayesHead
 * (4096 -> 15/11/24/4 raw logits, model_motifs_hierarchical.py:22-33).
 */
/* prod_rep[p] = cat(head_rep[idx0], tail_rep[idx1]) (roi_relation_predictors.py:413-419); edge_rep f32
 * [n_obj, 2*hidden] = post_emb output, pair_idx int32 [n_pairs,2] global object ids, out bf16 [n_pairs, 2*hidden]
 * (split = 0), [n_pairs, 3 * 2*hidden] in the bf16x3 layout (split = 1), or fp16 [n_pairs, 2*hidden] saturating at +-65504
 * (split = 2: the A operand of hc_tc_gemm with operand_f16 = 1). */
int hc_sgb_pair_gather(const float* edge_rep, const int32_t* pair_idx, int32_t n_pairs, int32_t hidden, int32_t split,
                       void* out, hc_stream_t stream);
/* bf16x3 operand splitting for near-fp32 accuracy on the bf16 tensor cores: f32 [n,k] -> bf16 [n,3k] = [hi | lo | hi]
 * with hi = bf16(x), lo = bf16(x - hi); pair it with weights packed as [W_hi | W_hi | W_lo] so that one hc_tc_gemm over
 * K' = 3k sums A_hi*W_hi + A_lo*W_hi + A_hi*W_lo in the fp32 accumulator.  (hc_sgb_pair_gather with split = 1 emits the
 * same layout directly.) */
int hc_split_bf16x3(const float* in, int64_t ld, int64_t n_rows, int32_t k, void* out, hc_stream_t stream);
/* frequency-bias gather + log-sum-exp super bias + hierarchical log-softmax with super indices 1..3
 * (roi_relation_predictors.py:430-459).  logits [n, ld]: columns [0,R) heads, [R,R+4) super.  bias_table f32
 * [num_obj^2, 51] or NULL, pair_pred int32 [n,2] object labels, label_ids int32 [R] 51-vocabulary id of each head
 * column (:376-382).  rel f32 [n,R] log-joint, super_rel f32 [n,4]. */
int hc_sgb_hier_softmax(const float* logits, int64_t ld, int32_t n_rows, int32_t n_geo, int32_t n_pos, int32_t n_sem,
                        const float* bias_table, int32_t num_obj, const int32_t* pair_pred, const int32_t* label_ids,
                        float* rel, float* super_rel, hc_stream_t stream);
/* HierarchPostProcessor candidates (inference.py:246-281): three per pair, score = max prob * obj_score0 *
 * obj_score1, label remapped through label_ids; image i with P_i pairs owns candidates [3*off_i, 3*off_{i+1}) in the
 * reference's torch.cat order c = 3*off_i + k*P_i + r.  cand_row = global pair row of each candidate. */
int hc_sgb_candidates(const float* rel, int32_t n_geo, int32_t n_pos, int32_t n_sem, const int32_t* pair_offsets,
                      const int32_t* pair_img, const int32_t* pair_idx, const float* obj_scores,
                      const int32_t* label_ids, int32_t n_rows, float* cand_score, int32_t* cand_label,
                      int32_t* cand_row, hc_stream_t stream);
/* Second (post-validator) sort restricted to the ranked window + per-image recall bookkeeping
 * (inference.py:292-302; sgg_eval.py:56-99,347-385,528-565; boxlist_ops.py:54-90 fp32 IoU with +1).
 * ranked int32 [n_images,128]: image-local candidate ids in first-sort order (hc_topk_match's topk_out with
 * top_max = 128), reject uint8 [n_images,128] or NULL.  gt_rel int32 [G,3] (sub id, obj id, label) with global ids
 * into gt_cls / gt_box (f32 xyxy); pred_cls / pred_box indexed by pair_idx.  Outputs per image:
 * final_rank [n_images, top_max] (optional), img_hits [n,3], img_ngt [n], img_hits_pc [n,3,51], img_cnt_pc [n,51]. */
int hc_sgb_rank_match(const int32_t* ranked, const uint8_t* reject, const int32_t* pair_offsets, int32_t n_images,
                      const float* cand_score, const int32_t* cand_label, const int32_t* cand_row,
                      const int32_t* pair_idx, const int32_t* pred_cls, const float* pred_box,
                      const int32_t* gt_offsets, const int32_t* gt_rel, const int32_t* gt_cls, const float* gt_box,
                      float iou_thresh, int32_t top_max, int32_t k0, int32_t k1, int32_t k2, int32_t* final_rank,
                      int32_t* img_hits, int32_t* img_ngt, int32_t* img_hits_pc, int32_t* img_cnt_pc,
                      hc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * N2 - SGDET / SGCLS proposal front-end that feeds the path (evaluate.py:311-370 == :545-591; the reference runs it
 * as Python loops with one device sync per element).
 *
 * hc_detr_proposals: DETR head outputs -> per-image NMS-ed proposal lists in padded staging + CSR offsets.
 *   pred_logits f32 [n_images, n_queries, num_classes+1] (last class = "no object"), pred_boxes f32
 *   [n_images, n_queries, 4] = (cx,cy,w,h) in 0..1.  Per query: softmax, has_object = argmax < num_classes
 *   (evaluate.py:311-312), top-`topk_cat` labels/probabilities in (value desc, class asc) order (:313-316), labels
 *   remapped through label_map int32 [num_classes+1] (dataset_utils.py:606-614, :319-323), boxes converted to
 *   (x1,x2,y1,y2) = clamp(c -/+ size/2, 0, 1) * feature_size in fp32 (:327-334); entries whose label maps to
 *   num_classes are dropped (:325,343-347); per-class greedy NMS with torchvision.ops.nms semantics in fp32
 *   (suppress iff inter/(area_i+area_j-inter) > nms_thresh, visit order = confidence desc, stable) (:350-367).
 *   Output order per image = (label asc, confidence desc, entry asc), the reference's hstack order.
 *   E = n_queries*topk_cat <= 1024.  Workspaces ws_* and staging st_* hold n_images*E entries (boxes 4 floats each);
 *   st_count [n_images]; box_offsets [n_images+1] receives the CSR offsets of the surviving proposals.
 *   Deviation: an image without any object query keeps its slot with 0 proposals (the reference drops it from its
 *   lists, misaligning them with the per-image targets).
 * hc_proposals_pack: staging -> CSR arrays sized box_offsets[n_images]: cats int32, conf f32, box_f f32 [n,4],
 *   box_i int32 [n,4] (int() truncation, the form hc_pairs_enumerate / hc_topk_match take), supers int8 [n,4] from
 *   sub2super int8 [num_classes,4] (evaluate.py:368-370; NULL to skip), box_img int32 (NULL to skip).
 */
int hc_detr_proposals(const float* pred_logits, const float* pred_boxes, int32_t n_images, int32_t n_queries,
                      int32_t num_classes, int32_t topk_cat, const int32_t* label_map, int32_t feature_size,
                      double nms_thresh, int32_t* ws_label, float* ws_conf, float* ws_box, uint8_t* ws_valid,
                      int32_t* st_label, float* st_conf, float* st_box, int32_t* st_count, int32_t* box_offsets,
                      hc_stream_t stream);
int hc_proposals_pack(const int32_t* st_label, const float* st_conf, const float* st_box, const int32_t* box_offsets,
                      int32_t n_images, int32_t n_entries, const int8_t* sub2super, int32_t num_classes, int32_t* cats,
                      float* conf, float* box_f, int32_t* box_i, int8_t* supers, int32_t* box_img, hc_stream_t stream);

/* utils.py:376-422 match_object_categories (SGCLS, evaluate.py:605): for every GT box, the two proposals of largest
 * rasterised-grid IoU (utils.py:58-74; double ratio rounded to fp32) under (IoU desc, proposal index asc); if the two
 * IoUs are equal both labels are emitted and the GT box is repeated, else the best one; confidence = proposal
 * confidence * IoU (fp32).  Phase 1 fills ws_idx/ws_iou [n_gt,2], out_offsets [n_images+1] and *status (device int32):
 * 1 when an image with GT boxes has fewer than two proposals - the reference then returns (None, None, None) and
 * skips the whole batch (utils.py:402-403, evaluate.py:606-607).  Phase 2 writes the out_offsets[n_images] rows:
 * out_box = the (repeated) GT boxes, out_src = GT box index of each row, out_supers / out_img optional. */
int hc_match_object_categories(const float* prop_box, const int32_t* prop_offsets, const int32_t* gt_box,
                               const int32_t* gt_offsets, int32_t n_images, int32_t feature_size, int32_t* ws_idx,
                               float* ws_iou, int32_t* ws_count, int32_t* out_offsets, int32_t* status,
                               hc_stream_t stream);
int hc_match_object_categories_fill(const int32_t* prop_cats, const float* prop_conf, const int32_t* prop_offsets,
                                    const int32_t* gt_box, const int32_t* gt_offsets, int32_t n_images,
                                    const int32_t* ws_idx, const float* ws_iou, const int32_t* out_offsets,
                                    const int8_t* sub2super, int32_t num_classes, int32_t* out_cats, float* out_conf,
                                    int32_t* out_box, int32_t* out_src, int8_t* out_supers, int32_t* out_img,
                                    hc_stream_t stream);

/* utils.py:294-352 match_target_sgd: flat GT triplet tables in the reference's (g,e) loop order from the packed
 * triangle arrays (rel_tri / dir_tri at t = g(g-1)/2+e as in hc_pairs_enumerate): a triplet where dir == 1
 * (g subject) or dir == 0 (e subject).  Reference quirk kept: g stops at N-2 (utils.py:312 loops over
 * range(len(relationships[image])) = N-1 rows).  gt_label/gt_sub/gt_obj must hold sum_i T_i entries; gt_sub/gt_obj are
 * global box ids; gt_offsets [n_images+1]. */
int hc_targets_flat(const int8_t* dir_tri, const int32_t* rel_tri, const int32_t* tri_offsets,
                    const int32_t* box_offsets, int32_t n_images, int32_t* ws_count, int32_t* gt_offsets,
                    int32_t* gt_label, int32_t* gt_sub, int32_t* gt_obj, hc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * N4 - training-side losses on the hierarchical head and their backward (train_utils.py:23-113 train_one_direction,
 * :116-157 calculate_losses_on_relationships; criteria of train_test.py:105-117; step accumulation train_test.py:187-258).
 * The reference evaluates the losses once per CALL (graph_iter, edge_iter, direction) over the images of the lock-step
 * batch that own box graph_iter.  Here every directed pair is a ROW (relation [n_rows, ld_rel] = the head's joint log-probs,
 * or raw logits for the flat head; super_rel [n_rows,3]; connectivity [n_rows]; all from hc_hier_head) and call m owns rows
 * group_rows[group_offsets[m] : group_offsets[m+1]].  row_target = directed predicate label, -1 = not connected (:62-73).
 *   loss_connectivity  connected rows exist: BCEWithLogits(conn[connected], 1) (it overwrites the other term, :88-90), else
 *                      lam_not_connected * BCEWithLogits(conn[not connected], 0) (:67-68)
 *   loss_relationship  hier: NLL(super[connected], super target) + sum_k NLLLoss(weight)(rel_k[connected_k], t - off_k);
 *                      flat: CrossEntropyLoss(weight)(relation[connected], t)
 *   loss_commonsense   bitmaps non-NULL (run_mode 'train_cs'): p = max softmax(rel_k) for each (row, k), triplet
 *                      (cat_sub, argmax + off_k, cat_obj): lam_cs_weak * mean(p[not in aligned]) + lam_cs_strong *
 *                      mean(p[in violated]) (:36-60); aligned_bitmap / violated_bitmap are plain membership bitmaps over keys
 *                      (s*50+p)*150+o (hc_cs_bitmap_build(keys, n, NULL, 0, out) builds one)
 * group_loss [n_groups,3] = (relationship, connectivity, commonsense) of each call, as train_one_direction returns them.
 * total [4]: total[0] = sum_m group_weight[m] * (rel_m + lam_conn*conn_m + lam_cs*cs_m) - with group_weight[m] = M - m this
 * is the reference's step loss, whose running-sum accumulation (`losses += loss_relationship + ...` with cumulative operands,
 * train_test.py:219-230) weights call m by the number of calls that follow it; total[1..3] = sum_m group_weight[m] * each part.
 * d_logits (optional) [n_rows, ld_dl]: gradient of total[0] with respect to the head's pre-softmax outputs in hc_hier_head's
 * w_heads row order (hier: fc3_1 | fc3_2 | fc3_3 | fc4 | fc5; flat: fc3 | fc4); rows in no call get zeros.
 * No float atomics: losses and gradients are bit-reproducible.
 */
int hc_hier_loss(const float* relation, int64_t ld_rel, const float* super_rel, const float* connectivity, int32_t n_rows,
                 int32_t n_geo, int32_t n_pos, int32_t n_sem, int32_t hier, float t1, float t2, float t3,
                 const int32_t* row_target, const int32_t* group_offsets, const int32_t* group_rows, int32_t n_groups,
                 const float* group_weight, const float* class_weight, const uint32_t* aligned_bitmap,
                 const uint32_t* violated_bitmap, const int32_t* row_sub, const int32_t* row_obj, const int32_t* box_cat,
                 float lam_conn, float lam_not_connected, float lam_cs, float lam_cs_weak, float lam_cs_strong,
                 float* group_loss, float* total, float* d_logits, int32_t ld_dl, hc_stream_t stream);

/* Backward of the heads fc3_x / fc4 / fc5 (model.py:171-183; what autograd does for the reference):
 *   d_pred [n_rows,512] = scale * d_logits @ w_heads,  d_w [n_out,512] = scale * d_logits^T @ pred,  d_b [n_out] = scale * sum_r.
 * scale: device scalar (the upstream gradient of the step loss) or NULL = 1.  d_pred and/or (d_w, d_b) may be NULL.
 * ws: parts * n_out * 513 floats of workspace for the two-stage weight-gradient sum (fixed order, no atomics). */
int hc_hier_head_bwd(const float* d_logits, int32_t ld_dl, const float* pred, int64_t ld_pred, int32_t n_rows, int32_t n_out,
                     const float* w_heads, const float* scale, float* d_pred, float* d_w, float* d_b, float* ws,
                     int32_t parts, hc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * R12 across ranks (SURVEY §8b "counts_allreduce", §8e): ONE ncclAllReduce(sum, int64) of the counter vector
 * (765 slots: Evaluator {hits[3], hits_pc[3x50], n_gt, n_gt_pc[50]} x {normal, zero-shot} + Evaluator_Top3), in place, on
 * `stream`.  nccl_comm is an ncclComm_t of the caller's NCCL (bound with dlopen at first use: no link-time dependency).
 * Counters are CUMULATIVE: reduce a window's delta (or reset before the window), never the same totals twice.
 * The reference has no cross-rank reduction (every rank writes its own JSON, utils.py:486).
 * hc_nccl_unique_id / hc_nccl_comm_create / hc_nccl_comm_destroy wrap ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy
 * for hosts that do not link NCCL themselves (one communicator per process, one process per GPU). */
typedef struct HcNcclId { char internal[128]; } HcNcclId;
int hc_counts_allreduce(void* nccl_comm, int64_t* counters, int64_t n, hc_stream_t stream);
int hc_nccl_unique_id(HcNcclId* id_out);
int hc_nccl_comm_create(const HcNcclId* id, int32_t n_ranks, int32_t rank, void** comm_out);
int hc_nccl_comm_destroy(void* comm);

/* ---------------------------------------------------------------------------------------------------------
 * Sizing queries (SURVEY §8b "Ownership"): the caller owns every allocation, kernels never allocate.  Host functions,
 * no GPU needed.  Return bytes / entries (>= 0) or a negative status.
 *   hc_pairs_enumerate_workspace_bytes : ws_ov / ws_any / ws_counts of hc_pairs_enumerate (sum_tri = sum_i N_i(N_i-1)/2)
 *   hc_conv3_blocks_capacity           : int32 entries of a conv3_1 work list for n_pairs directed pairs
 *   hc_conv2_box_blocks_capacity       : int32 entries of the conv2_1 box-footprint work list
 *   hc_relation_workspace_bytes        : every device buffer of the batched relation path for one window */
int64_t hc_pairs_enumerate_workspace_bytes(int64_t sum_tri, int32_t n_images, int32_t n_groups, int32_t max_tri,
                                           int64_t* ws_ov_bytes, int64_t* ws_any_bytes, int64_t* ws_counts_bytes);
int64_t hc_conv3_blocks_capacity(int64_t n_pairs, int32_t block_rows, int32_t block_cols);
int64_t hc_conv2_box_blocks_capacity(int64_t n_box, int32_t block_rows);
typedef struct hc_relation_workspace {
  int64_t pixels_packed, conv1_out, box_select, conv2_halves, pooled_conv2, work_lists, box_maps, box_fc1_rows, fc1_operand,
      row_maps, pooled_conv3, fc1_out, fc2_raw, head_out, candidates, total;
} hc_relation_workspace;
int hc_relation_workspace_bytes(int32_t n_images, int64_t n_box, int64_t n_pairs, int64_t chunk_pairs, int32_t shared_fc1,
                                hc_relation_workspace* out);

#ifdef __cplusplus
}
#endif
#endif /* HIERCOM_B200_H_ */
