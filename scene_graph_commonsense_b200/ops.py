"""`torch.ops.hiercom.*` - the torch custom-op layer over the C ABI (north_star "thin C-ABI torch custom-op layer").

One operator per entry point of include/hiercom_b200.h, declared with `torch.library` and implemented for the **CUDA dispatch key
only**: a CPU tensor reaches no kernel (`NotImplementedError` from the dispatcher - there is no CPU fallback to fall into).  The
CUDA implementations are the ctypes binders of `_abi_ops` (raw pointers + sizes + the current stream into libhiercom_b200.so).
The module-level functions below keep the call signatures the rest of the package uses (optional operands, dict returns) and
route every call through `torch.ops.hiercom.<name>`, so model.py, evaluator.py, pipeline.py, sgb.py and frontend.py all reach
the kernels through the dispatcher.  Operators that fill a caller-owned buffer declare it `Tensor(a!)`.
"""
import numpy as np
import torch

from . import _abi_ops as _A
from . import tables
from ._abi_ops import LAUNCHES, PROFILE, cs_bitmap_build, pack_weight_bf16x3  # noqa: F401  (host-side helpers, no kernel)
from ._lib import (ACT_NONE, ACT_RELU, ACT_TANH, EPI_BF16, EPI_F32, EPI_POOL_BF16, EPI_POOL_DIFF_BF16, EPI_SPLIT3_BF16,  # noqa: F401
                   GEMM_CONV3, GEMM_CONV3_BLOCKS, GEMM_PLAIN)

NAMESPACE = "hiercom"
_LIB = torch.library.Library(NAMESPACE, "DEF")
SCHEMAS = {}


def _op(name, schema):
    """Declare `hiercom::<name><schema>` and register the decorated function as its CUDA kernel."""
    def deco(fn):
        _LIB.define(name + schema)
        _LIB.impl(name, fn, "CUDA")
        SCHEMAS[name] = schema
        return fn
    return deco


def _call(name):
    return getattr(getattr(torch.ops, NAMESPACE), name)


def _some(t, like):
    """None -> a 0-element tensor (operator returns cannot be optional)."""
    return t if t is not None else torch.empty(0, device=like.device)


# ------------------------------------------------------------------------------------------------ R1/R2/R4 pair enumeration
@_op("pairs_enumerate", "(Tensor boxes, Tensor box_offsets, Tensor tri_offsets, int p_max, Tensor? rel_tri, Tensor? dir_tri, "
     "Tensor? group_id, int n_groups, int max_tri, int feature_size, int[]? offsets_host) -> Tensor[]")
def _pairs_enumerate(boxes, box_offsets, tri_offsets, p_max, rel_tri, dir_tri, group_id, n_groups, max_tri, feature_size, offsets_host):
    known = None if offsets_host is None else np.asarray(offsets_host, dtype=np.int32)
    d = _A.pairs_enumerate(boxes, box_offsets, tri_offsets, p_max, rel_tri, dir_tri, group_id, n_groups, max_tri, feature_size, offsets_host=known)
    return [d["offsets"], torch.from_numpy(np.ascontiguousarray(d["offsets_host"])), d["sub"], d["obj"], d["img"], d["ov"], d["gt"], d["rel"]]


def pairs_enumerate(boxes, box_offsets, tri_offsets, p_max, rel_tri=None, dir_tri=None, group_id=None, n_groups=0, max_tri=0,
                    feature_size=32, offsets_host=None):
    """`offsets_host` (int32 [B+1], pipeline.host_pair_offsets): the per-image pair offsets counted on the host - with them the call
    reads nothing back from the device."""
    known = None if offsets_host is None else [int(v) for v in offsets_host]
    off, off_host, sub, obj, img, ov, gt, rel = _call("pairs_enumerate")(boxes, box_offsets, tri_offsets, p_max, rel_tri, dir_tri,
                                                                        group_id, n_groups, max_tri, feature_size, known)
    off_host = off_host.numpy()
    return dict(n=int(off_host[-1]), offsets=off, offsets_host=off_host, sub=sub, obj=obj, img=img, ov=ov, gt=gt, rel=rel)


# ------------------------------------------------------------------------------------------------ R5/R6 dense contractions
@_op("tc_gemm", "(Tensor a, Tensor b, Tensor(a!) out, int m, int n, int k, Tensor? bias, Tensor? mul, int lda, int ldc, int c_off, "
     "int mode, int epilogue, int act, int n_img, int h, int w, int c_total, int c_base, int c_in, int group_m, int m_sub, "
     "str tag, Tensor? blocks, Tensor? n_blocks, int block_rows, int block_cols, Tensor? k_masks, int k_cell, Tensor? add_a, Tensor? add_a_rows, "
     "Tensor? add_b, Tensor? add_b_rows, Tensor? out_rows, Tensor? diff_sub, Tensor? diff_obj, Tensor? diff_bg, Tensor? pair_sub, "
     "Tensor? pair_obj, Tensor? pair_row, int cta_pairs, Tensor(b!)? scratch, Tensor? m_order) -> ()")
def _tc_gemm(a, b, out, m, n, k, bias, mul, lda, ldc, c_off, mode, epilogue, act, n_img, h, w, c_total, c_base, c_in, group_m, m_sub,
             tag, blocks, n_blocks, block_rows, block_cols, k_masks, k_cell, add_a, add_a_rows, add_b, add_b_rows, out_rows, diff_sub, diff_obj,
             diff_bg, pair_sub, pair_obj, pair_row, cta_pairs, scratch, m_order):
    _A.tc_gemm(a, b, out, m, n, k, bias=bias, lda=lda, ldc=ldc, c_off=c_off, mode=mode, epilogue=epilogue, act=act, n_img=n_img, h=h,
               w=w, c_total=c_total, c_base=c_base, c_in=c_in, group_m=group_m, m_sub=m_sub, tag=tag, mul=mul, blocks=blocks,
               n_blocks=n_blocks, block_rows=block_rows, block_cols=block_cols, k_masks=k_masks, k_cell=k_cell, add_a=add_a, add_a_rows=add_a_rows, add_b=add_b,
               add_b_rows=add_b_rows, out_rows=out_rows, diff_sub=diff_sub, diff_obj=diff_obj, diff_bg=diff_bg, pair_sub=pair_sub,
               pair_obj=pair_obj, pair_row=pair_row, cta_pairs=cta_pairs, scratch=scratch, m_order=m_order)


def tc_gemm(a, b, out, m, n, k, *, bias=None, lda=0, ldc=None, c_off=0, mode=GEMM_PLAIN, epilogue=EPI_BF16, act=ACT_NONE, n_img=0,
            h=0, w=0, c_total=0, c_base=0, c_in=0, group_m=0, m_sub=0, tag="tc_gemm", mul=None, blocks=None, n_blocks=None,
            block_rows=0, block_cols=0, k_masks=None, k_cell=0, add_a=None, add_a_rows=None, add_b=None, add_b_rows=None, out_rows=None, diff_sub=None,
            diff_obj=None, diff_bg=None, pair_sub=None, pair_obj=None, pair_row=None, cta_pairs=0, scratch=None, m_order=None):
    """out = epilogue(A @ B^T) on tcgen05 (include/hiercom_b200.h hc_tc_gemm)."""
    _call("tc_gemm")(a, b, out, m, n, k, bias, mul, lda, n if ldc is None else ldc, c_off, mode, epilogue, act, n_img, h, w, c_total,
                     c_base, c_in, group_m, m_sub, tag, blocks, n_blocks, block_rows, block_cols, k_masks, k_cell, add_a, add_a_rows, add_b, add_b_rows,
                     out_rows, diff_sub, diff_obj, diff_bg, pair_sub, pair_obj, pair_row, cta_pairs, scratch, m_order)
    return out


# ------------------------------------------------------------------------------------------------ block-sparse conv3_1 support
@_op("conv3_active_blocks", "(Tensor boxes, Tensor pair_sub, Tensor pair_obj, int block_rows, int fs, Tensor(a!) blocks, "
     "Tensor(b!) n_blocks, int block_cols) -> ()")
def _conv3_active_blocks(boxes, pair_sub, pair_obj, block_rows, fs, blocks, n_blocks, block_cols):
    _A.conv3_active_blocks(boxes, pair_sub, pair_obj, block_rows, fs, blocks=blocks, n_blocks=n_blocks, block_cols=block_cols)


def conv3_active_blocks(boxes, pair_sub, pair_obj, block_rows=8, fs=32, blocks=None, n_blocks=None, block_cols=8):
    """Device work list (blocks, n_blocks) of the conv3_1 output blocks a set of directed pairs has to compute."""
    if blocks is None:
        blocks = torch.empty(max(pair_sub.numel() * (256 // (block_rows * block_cols)), 1), dtype=torch.int32, device=boxes.device)
    if n_blocks is None:
        n_blocks = torch.empty(1, dtype=torch.int32, device=boxes.device)
    _call("conv3_active_blocks")(boxes, pair_sub, pair_obj, block_rows, fs, blocks, n_blocks, block_cols)
    return blocks, n_blocks


@_op("conv3_shared_blocks", "(Tensor boxes, Tensor pair_sub, Tensor pair_obj, int block_rows, int fs, Tensor(a!) blocks, "
     "Tensor(b!) n_blocks, int block_cols) -> ()")
def _conv3_shared_blocks(boxes, pair_sub, pair_obj, block_rows, fs, blocks, n_blocks, block_cols):
    _A.conv3_active_blocks(boxes, pair_sub, pair_obj, block_rows, fs, blocks=blocks, n_blocks=n_blocks, shared=True, block_cols=block_cols)


def conv3_shared_blocks(boxes, pair_sub, pair_obj, block_rows=4, fs=32, blocks=None, n_blocks=None, block_cols=8):
    """Device work list of the conv3_1 output blocks that depend on BOTH boxes of a pair (the rest comes from `p3_assemble`)."""
    if blocks is None:
        blocks = torch.empty(max(pair_sub.numel() * (256 // (block_rows * block_cols)), 1), dtype=torch.int32, device=boxes.device)
    if n_blocks is None:
        n_blocks = torch.empty(1, dtype=torch.int32, device=boxes.device)
    _call("conv3_shared_blocks")(boxes, pair_sub, pair_obj, block_rows, fs, blocks, n_blocks, block_cols)
    return blocks, n_blocks


@_op("p3_assemble", "(Tensor background, Tensor sub_maps, Tensor obj_maps, Tensor boxes, Tensor pair_sub, Tensor pair_obj, "
     "Tensor(a!) out, int fs) -> ()")
def _p3_assemble(background, sub_maps, obj_maps, boxes, pair_sub, pair_obj, out, fs):
    _A.p3_assemble(background, sub_maps, obj_maps, boxes, pair_sub, pair_obj, out, fs)


def p3_assemble(background, sub_maps, obj_maps, boxes, pair_sub, pair_obj, out, fs=32):
    _call("p3_assemble")(background, sub_maps, obj_maps, boxes, pair_sub, pair_obj, out, fs)
    return out


@_op("conv2_box_blocks", "(Tensor boxes, int block_rows, int fs, Tensor(a!) blocks, Tensor(b!) n_blocks) -> ()")
def _conv2_box_blocks(boxes, block_rows, fs, blocks, n_blocks):
    _A.conv2_box_blocks(boxes, block_rows, fs, blocks=blocks, n_blocks=n_blocks)


def conv2_box_blocks(boxes, block_rows=4, fs=32, blocks=None, n_blocks=None):
    """Device work list (blocks, n_blocks) of the conv2_1 output blocks within one pixel of each box (everything else is background)."""
    if blocks is None:
        blocks = torch.empty(max(boxes.shape[0] * 4 * (32 // block_rows), 1), dtype=torch.int32, device=boxes.device)
    if n_blocks is None:
        n_blocks = torch.empty(1, dtype=torch.int32, device=boxes.device)
    _call("conv2_box_blocks")(boxes, block_rows, fs, blocks, n_blocks)
    return blocks, n_blocks


@_op("pair_cell_keys", "(Tensor boxes, Tensor pair_sub, Tensor pair_obj, int fs) -> Tensor")
def _pair_cell_keys(boxes, pair_sub, pair_obj, fs):
    return _A.pair_cell_keys(boxes, pair_sub, pair_obj, fs)


def pair_cell_keys(boxes, pair_sub, pair_obj, fs=32):
    """Sort key of the cell rectangle both boxes of each directed pair reach (row order of the shared-footprint fc1)."""
    return _call("pair_cell_keys")(boxes, pair_sub, pair_obj, fs)


@_op("tile_cell_masks", "(Tensor boxes, Tensor row_sub, Tensor row_obj, int rows_per_tile, int fs) -> Tensor")
def _tile_cell_masks(boxes, row_sub, row_obj, rows_per_tile, fs):
    return _A.tile_cell_masks(boxes, row_sub, row_obj, rows_per_tile, fs)


def tile_cell_masks(boxes, row_sub, row_obj, rows_per_tile, fs=32):
    """Per GEMM tile of `rows_per_tile` sorted rows: int64 bitmap of the K cells the tile visits (`k_masks` of tc_gemm)."""
    return _call("tile_cell_masks")(boxes, row_sub, row_obj, rows_per_tile, fs)


@_op("cells_zero", "(Tensor masks, int rows_per_tile, int n_rows, Tensor(a!) out) -> ()")
def _cells_zero(masks, rows_per_tile, n_rows, out):
    _A.cells_zero(masks, rows_per_tile, n_rows, out)


def cells_zero(masks, rows_per_tile, n_rows, out):
    _call("cells_zero")(masks, rows_per_tile, n_rows, out)
    return out


@_op("broadcast_rows", "(Tensor src, int n_rows, Tensor(a!) out) -> ()")
def _broadcast_rows(src, n_rows, out):
    _A.broadcast_rows(src, n_rows, out)


def broadcast_rows(src, n_rows, out):
    _call("broadcast_rows")(src, n_rows, out)
    return out


# ------------------------------------------------------------------------------------------------ R3 gather / pooling producers
@_op("pack_pixels", "(Tensor src0, Tensor? src1, int k_pad, bool f16) -> Tensor")
def _pack_pixels(src0, src1, k_pad, f16):
    return _A.pack_pixels(src0, src1, k_pad, dtype=torch.float16 if f16 else torch.bfloat16)


def pack_pixels(src0, src1, k_pad, out=None, dtype=torch.bfloat16):
    if out is not None:
        raise RuntimeError("hiercom_b200: pack_pixels allocates its result")
    return _call("pack_pixels")(src0, src1, k_pad, dtype == torch.float16)


@_op("box_select", "(Tensor t_img, Tensor boxes, Tensor box_img, Tensor fill, int fs) -> Tensor")
def _box_select(t_img, boxes, box_img, fill, fs):
    return _A.box_select(t_img, boxes, box_img, fill, fs)


def box_select(t_img, boxes, box_img, fill, fs=32, out=None):
    if out is not None:
        raise RuntimeError("hiercom_b200: box_select allocates its result")
    return _call("box_select")(t_img, boxes, box_img, fill, fs)


def _fp_pack(fp):
    return (None, None, None, 0) if fp is None else (fp[0], fp[1], fp[2], int(fp[3]))


def _fp_unpack(boxes, u_bg, v_bg, rows):
    return None if boxes is None else (boxes, u_bg, v_bg, rows)


@_op("pair_relu_pool", "(Tensor u, Tensor v, Tensor? bias, Tensor pair_sub, Tensor pair_obj, int fs, Tensor(a!) out, Tensor? cover, "
     "Tensor? fp_boxes, Tensor? u_bg, Tensor? v_bg, int fp_block_rows) -> ()")
def _pair_relu_pool(u, v, bias, pair_sub, pair_obj, fs, out, cover, fp_boxes, u_bg, v_bg, fp_block_rows):
    _A.pair_relu_pool(u, v, bias, pair_sub, pair_obj, fs, out=out, cover=cover, fp=_fp_unpack(fp_boxes, u_bg, v_bg, fp_block_rows))


def pair_relu_pool(u, v, bias, pair_sub, pair_obj, fs=32, out=None, cover=None, fp=None):
    """`cover` (int64 per pair, `pair_cover_masks`): write only the pooled pixels a listed conv3_1 block reads (packed path, bias None).
    `fp` = (boxes, u_bg, v_bg, conv2 block rows): u / v hold a box's values only inside its conv2_1 footprint (no background pre-fill);
    the background maps are read elsewhere (`PackedHead.conv2_halves_sparse(prefill=False)`)."""
    if out is None:
        out = torch.empty(pair_sub.numel(), fs // 2, fs // 2, u.shape[-1], dtype=u.dtype, device=u.device)
    _call("pair_relu_pool")(u, v, bias, pair_sub, pair_obj, fs, out, cover, *_fp_pack(fp))
    return out


@_op("pair_lut_build", "(Tensor pair_sub, Tensor pair_obj, Tensor pair_img, Tensor box_offsets, int n_box, int n_max) -> Tensor")
def _pair_lut_build(pair_sub, pair_obj, pair_img, box_offsets, n_box, n_max):
    return _A.pair_lut_build(pair_sub, pair_obj, pair_img, box_offsets, n_box, n_max)


def pair_lut_build(pair_sub, pair_obj, pair_img, box_offsets, n_box, n_max):
    return _call("pair_lut_build")(pair_sub, pair_obj, pair_img, box_offsets, n_box, n_max)


@_op("pair_cover_masks", "(Tensor boxes, Tensor pair_sub, Tensor pair_obj, int block_rows, int block_cols, bool shared, int fs, "
     "Tensor(a!) out) -> ()")
def _pair_cover_masks(boxes, pair_sub, pair_obj, block_rows, block_cols, shared, fs, out):
    _A.pair_cover_masks(boxes, pair_sub, pair_obj, block_rows, block_cols, shared, fs, out=out)


def pair_cover_masks(boxes, pair_sub, pair_obj, block_rows, block_cols, shared, fs=32, out=None):
    """Per pair: int64 bitmap of the cells its listed conv3_1 blocks cover (the `cover` of `pair_relu_pool_tiled`)."""
    if out is None:
        out = torch.empty(max(pair_sub.numel(), 1), dtype=torch.int64, device=boxes.device)
    _call("pair_cover_masks")(boxes, pair_sub, pair_obj, block_rows, block_cols, shared, fs, out)
    return out


@_op("pair_relu_pool_tiled", "(Tensor u, Tensor v, Tensor? bias, Tensor box_offsets, Tensor lut, int img0, int n_img, int pair_base, "
     "int chunk_pairs, int fs, Tensor(a!) out, Tensor? cover, Tensor? fp_boxes, Tensor? u_bg, Tensor? v_bg, int fp_block_rows) -> ()")
def _pair_relu_pool_tiled(u, v, bias, box_offsets, lut, img0, n_img, pair_base, chunk_pairs, fs, out, cover, fp_boxes, u_bg, v_bg, fp_block_rows):
    _A.pair_relu_pool_tiled(u, v, bias, box_offsets, lut, img0, n_img, pair_base, chunk_pairs, fs, out=out, cover=cover,
                            fp=_fp_unpack(fp_boxes, u_bg, v_bg, fp_block_rows))


def pair_relu_pool_tiled(u, v, bias, box_offsets, lut, img0, n_img, pair_base, chunk_pairs, fs=32, out=None, cover=None, fp=None):
    """`cover` (int64 per pair of the chunk, `pair_cover_masks`): write only the pooled pixels a listed conv3_1 block reads.
    `fp`: footprint-only u / v, see `pair_relu_pool`."""
    if out is None:
        out = torch.empty(chunk_pairs, fs // 2, fs // 2, u.shape[-1], dtype=u.dtype, device=u.device)
    _call("pair_relu_pool_tiled")(u, v, bias, box_offsets, lut, img0, n_img, pair_base, chunk_pairs, fs, out, cover, *_fp_pack(fp))
    return out


# ------------------------------------------------------------------------------------------------ R6 tail + R7 hierarchical head
@_op("hier_head", "(Tensor fc2_raw, Tensor? fc2_bias, Tensor? emb, Tensor? row_sub, Tensor? row_obj, Tensor? box_cat, "
     "Tensor? box_super, Tensor w_heads, Tensor b_heads, int[] splits, bool flat, float[] temps, int num_obj, int num_super, "
     "bool want_pred) -> (Tensor, Tensor, Tensor, Tensor, Tensor)")
def _hier_head(fc2_raw, fc2_bias, emb, row_sub, row_obj, box_cat, box_super, w_heads, b_heads, splits, flat, temps, num_obj, num_super,
               want_pred):
    relation, sup, conn, logsig, pred = _A.hier_head(fc2_raw, fc2_bias, emb, row_sub, row_obj, box_cat, box_super, w_heads, b_heads,
                                                     tuple(splits), flat, tuple(temps), num_obj, num_super, want_pred)
    return relation, _some(sup, relation), conn, logsig, _some(pred, relation)


def hier_head(fc2_raw, fc2_bias, emb, row_sub, row_obj, box_cat, box_super, w_heads, b_heads, splits, flat=False,
              temps=(1.0, 1.0, 1.0), num_obj=150, num_super=17, want_pred=False):
    relation, sup, conn, logsig, pred = _call("hier_head")(fc2_raw, fc2_bias, emb, row_sub, row_obj, box_cat, box_super, w_heads,
                                                            b_heads, list(splits), bool(flat), [float(t) for t in temps], num_obj,
                                                            num_super, bool(want_pred))
    return relation, (None if flat else sup), conn, logsig, (pred if want_pred else None)


# ------------------------------------------------------------------------------------------------ N4 training losses + head backward
@_op("hier_loss", "(Tensor relation, Tensor? super_rel, Tensor connectivity, Tensor row_target, Tensor group_offsets, Tensor group_rows, "
     "Tensor group_weight, Tensor class_weight, int[] splits, bool hier, float[] temps, Tensor? aligned_bitmap, Tensor? violated_bitmap, "
     "Tensor? row_sub, Tensor? row_obj, Tensor? box_cat, float[] lambdas, bool want_grad) -> (Tensor, Tensor, Tensor)")
def _hier_loss(relation, super_rel, connectivity, row_target, group_offsets, group_rows, group_weight, class_weight, splits, hier, temps,
               aligned_bitmap, violated_bitmap, row_sub, row_obj, box_cat, lambdas, want_grad):
    gl, total, dl = _A.hier_loss(relation, super_rel, connectivity, row_target, group_offsets, group_rows, group_weight, class_weight,
                                 tuple(splits), hier, tuple(temps), aligned_bitmap, violated_bitmap, row_sub, row_obj, box_cat,
                                 tuple(lambdas), want_grad)
    return gl, total, _some(dl, relation)


def hier_loss(relation, super_rel, connectivity, row_target, group_offsets, group_rows, group_weight, class_weight, splits, hier=True,
              temps=(1.0, 1.0, 1.0), aligned_bitmap=None, violated_bitmap=None, row_sub=None, row_obj=None, box_cat=None,
              lambdas=(0.1, 1.0, 1.0, 0.1, 10.0), want_grad=True):
    """Per-call training losses of train_utils.train_one_direction + d(step loss)/d(head logits) (hc_hier_loss)."""
    gl, total, dl = _call("hier_loss")(relation, super_rel, connectivity, row_target, group_offsets, group_rows, group_weight, class_weight,
                                       list(splits), bool(hier), [float(t) for t in temps], aligned_bitmap, violated_bitmap, row_sub,
                                       row_obj, box_cat, [float(x) for x in lambdas], bool(want_grad))
    return gl, total, (dl if want_grad else None)


@_op("hier_head_bwd", "(Tensor d_logits, Tensor pred, Tensor w_heads, Tensor? scale, bool want_pred, bool want_weights) "
     "-> (Tensor, Tensor, Tensor)")
def _hier_head_bwd(d_logits, pred, w_heads, scale, want_pred, want_weights):
    d_pred, d_w, d_b = _A.hier_head_bwd(d_logits, pred, w_heads, scale, want_pred, want_weights)
    return _some(d_pred, pred), _some(d_w, pred), _some(d_b, pred)


def hier_head_bwd(d_logits, pred, w_heads, scale=None, want_pred=True, want_weights=True):
    d_pred, d_w, d_b = _call("hier_head_bwd")(d_logits, pred, w_heads, scale, bool(want_pred), bool(want_weights))
    return (d_pred if want_pred else None), (d_w if want_weights else None), (d_b if want_weights else None)


# ------------------------------------------------------------------------------------------------ R8/R9/R10 candidates
@_op("candidates", "(Tensor relation, int[] splits, bool hier, Tensor row_ov, Tensor logsig, Tensor row_sub, Tensor row_obj, "
     "Tensor box_cat, Tensor? pass_bitmap, Tensor? super_rel, Tensor? conf_sub, Tensor? conf_obj, int layout, bool want_top3) "
     "-> (Tensor, Tensor, Tensor, Tensor)")
def _candidates(relation, splits, hier, row_ov, logsig, row_sub, row_obj, box_cat, pass_bitmap, super_rel, conf_sub, conf_obj, layout,
                want_top3):
    conf, label, t3_conf, t3_super = _A.candidates(relation, tuple(splits), hier, row_ov, logsig, row_sub, row_obj, box_cat, pass_bitmap,
                                                   super_rel, conf_sub, conf_obj, layout, want_top3)
    return conf, label, _some(t3_conf, conf), _some(t3_super, conf)


def candidates(relation, splits, hier, row_ov, logsig, row_sub, row_obj, box_cat, pass_bitmap=None, super_rel=None, conf_sub=None,
               conf_obj=None, layout=0, want_top3=False):
    conf, label, t3_conf, t3_super = _call("candidates")(relation, list(splits), bool(hier), row_ov, logsig, row_sub, row_obj, box_cat,
                                                          pass_bitmap, super_rel, conf_sub, conf_obj, layout, bool(want_top3))
    return conf, label, (t3_conf if want_top3 else None), (t3_super if want_top3 else None)


# ------------------------------------------------------------------------------------------------ R10-R13 top-K + matching
@_op("topk_match", "(Tensor cand_offsets, Tensor cand_conf, Tensor? cand_label, int k_per_row, Tensor row_sub, Tensor row_obj, "
     "Tensor pred_cat, Tensor pred_box, Tensor gt_offsets, Tensor gt_label, Tensor gt_sub, Tensor gt_obj, Tensor gt_cat, "
     "Tensor gt_box, Tensor(a!) counters, Tensor? cand_row, Tensor? synonyms, Tensor? zs_bitmap, int mode, Tensor? t3_labels, "
     "Tensor? t3_super, int feature_size, float iou_thresh, int[] top_k, bool want_topk) -> Tensor")
def _topk_match(cand_offsets, cand_conf, cand_label, k_per_row, row_sub, row_obj, pred_cat, pred_box, gt_offsets, gt_label, gt_sub,
                gt_obj, gt_cat, gt_box, counters, cand_row, synonyms, zs_bitmap, mode, t3_labels, t3_super, feature_size, iou_thresh,
                top_k, want_topk):
    out = _A.topk_match(cand_offsets, cand_conf, cand_label, k_per_row, row_sub, row_obj, pred_cat, pred_box, gt_offsets, gt_label,
                        gt_sub, gt_obj, gt_cat, gt_box, counters, cand_row=cand_row, synonyms=synonyms, zs_bitmap=zs_bitmap, mode=mode,
                        t3_labels=t3_labels, t3_super=t3_super, feature_size=feature_size, iou_thresh=iou_thresh, top_k=tuple(top_k),
                        want_topk=want_topk)
    return _some(out, cand_conf)


def topk_match(cand_offsets, cand_conf, cand_label, k_per_row, row_sub, row_obj, pred_cat, pred_box, gt_offsets, gt_label, gt_sub,
               gt_obj, gt_cat, gt_box, counters, *, cand_row=None, synonyms=None, zs_bitmap=None, mode=0, t3_labels=None,
               t3_super=None, feature_size=32, iou_thresh=0.5, top_k=tables.TOP_K, want_topk=False):
    out = _call("topk_match")(cand_offsets, cand_conf, cand_label, k_per_row, row_sub, row_obj, pred_cat, pred_box, gt_offsets,
                              gt_label, gt_sub, gt_obj, gt_cat, gt_box, counters, cand_row, synonyms, zs_bitmap, mode, t3_labels,
                              t3_super, feature_size, float(iou_thresh), [int(k) for k in top_k], bool(want_topk))
    return out if want_topk else None


@_op("topk_select", "(Tensor cand_offsets, Tensor cand_conf, int top_max) -> Tensor")
def _topk_select(cand_offsets, cand_conf, top_max):
    return _A.topk_select(cand_offsets, cand_conf, top_max)


def topk_select(cand_offsets, cand_conf, top_max=128):
    return _call("topk_select")(cand_offsets, cand_conf, top_max)


@_op("connectivity_stats", "(Tensor connectivity, Tensor gt_directed, Tensor gt_undirected, Tensor(a!) stats) -> ()")
def _connectivity_stats(connectivity, gt_directed, gt_undirected, stats):
    _A.connectivity_stats(connectivity, gt_directed, gt_undirected, stats)


def connectivity_stats(connectivity, gt_directed, gt_undirected, stats):
    _call("connectivity_stats")(connectivity, gt_directed, gt_undirected, stats)


# ------------------------------------------------------------------------------------------------ R14 / N1 SGB twin
@_op("sgb_pair_gather", "(Tensor edge_rep, Tensor pair_idx, int hidden, bool split, bool f16) -> Tensor")
def _sgb_pair_gather(edge_rep, pair_idx, hidden, split, f16):
    return _A.sgb_pair_gather(edge_rep, pair_idx, hidden, split, f16)


def sgb_pair_gather(edge_rep, pair_idx, hidden, split=False, f16=False):
    return _call("sgb_pair_gather")(edge_rep, pair_idx, hidden, bool(split), bool(f16))


@_op("split_bf16x3", "(Tensor x) -> Tensor")
def _split_bf16x3(x):
    return _A.split_bf16x3(x)


def split_bf16x3(x):
    """f32 [n,k] -> bf16 [n,3k] = [hi | lo | hi] (A side of the bf16x3 scheme)."""
    return _call("split_bf16x3")(x)


@_op("sgb_hier_softmax", "(Tensor logits, int[] splits, Tensor? bias_table, int num_obj, Tensor? pair_pred, Tensor? label_ids) "
     "-> (Tensor, Tensor)")
def _sgb_hier_softmax(logits, splits, bias_table, num_obj, pair_pred, label_ids):
    return _A.sgb_hier_softmax(logits, tuple(splits), bias_table, num_obj, pair_pred, label_ids)


def sgb_hier_softmax(logits, splits, bias_table=None, num_obj=151, pair_pred=None, label_ids=None):
    return _call("sgb_hier_softmax")(logits, list(splits), bias_table, num_obj, pair_pred, label_ids)


@_op("sgb_candidates", "(Tensor rel, int[] splits, Tensor pair_offsets, Tensor pair_img, Tensor pair_idx, Tensor obj_scores, "
     "Tensor label_ids) -> (Tensor, Tensor, Tensor)")
def _sgb_candidates(rel, splits, pair_offsets, pair_img, pair_idx, obj_scores, label_ids):
    return _A.sgb_candidates(rel, tuple(splits), pair_offsets, pair_img, pair_idx, obj_scores, label_ids)


def sgb_candidates(rel, splits, pair_offsets, pair_img, pair_idx, obj_scores, label_ids):
    return _call("sgb_candidates")(rel, list(splits), pair_offsets, pair_img, pair_idx, obj_scores, label_ids)


@_op("sgb_rank_match", "(Tensor ranked, Tensor? reject, Tensor pair_offsets, Tensor cand_score, Tensor cand_label, Tensor cand_row, "
     "Tensor pair_idx, Tensor pred_cls, Tensor pred_box, Tensor gt_offsets, Tensor gt_rel, Tensor gt_cls, Tensor gt_box, "
     "float iou_thresh, int[] top_k) -> (Tensor, Tensor, Tensor, Tensor, Tensor)")
def _sgb_rank_match(ranked, reject, pair_offsets, cand_score, cand_label, cand_row, pair_idx, pred_cls, pred_box, gt_offsets, gt_rel,
                    gt_cls, gt_box, iou_thresh, top_k):
    return _A.sgb_rank_match(ranked, reject, pair_offsets, cand_score, cand_label, cand_row, pair_idx, pred_cls, pred_box, gt_offsets,
                             gt_rel, gt_cls, gt_box, iou_thresh, tuple(top_k))


def sgb_rank_match(ranked, reject, pair_offsets, cand_score, cand_label, cand_row, pair_idx, pred_cls, pred_box, gt_offsets, gt_rel,
                   gt_cls, gt_box, iou_thresh=0.5, top_k=(20, 50, 100)):
    return _call("sgb_rank_match")(ranked, reject, pair_offsets, cand_score, cand_label, cand_row, pair_idx, pred_cls, pred_box,
                                   gt_offsets, gt_rel, gt_cls, gt_box, float(iou_thresh), [int(k) for k in top_k])


# ------------------------------------------------------------------------------------------------ N2 proposal front-end
_PROP_KEYS = ("offsets", "cats", "conf", "box_f", "box_i", "box_img")


@_op("detr_proposals", "(Tensor pred_logits, Tensor pred_boxes, Tensor label_map, Tensor? sub2super, int num_classes, int topk_cat, "
     "int feature_size, float nms_thresh) -> Tensor[]")
def _detr_proposals(pred_logits, pred_boxes, label_map, sub2super, num_classes, topk_cat, feature_size, nms_thresh):
    d = _A.detr_proposals(pred_logits, pred_boxes, label_map, sub2super, num_classes, topk_cat, feature_size, nms_thresh)
    out = [torch.from_numpy(d["offsets_host"])] + [d[k] for k in _PROP_KEYS]
    return out + ([d["supers"]] if d["supers"] is not None else [])


def detr_proposals(pred_logits, pred_boxes, label_map, sub2super=None, num_classes=150, topk_cat=2, feature_size=32, nms_thresh=0.5):
    """evaluate.py:311-370.  Returns CSR device arrays of the surviving proposals (one [B+1]-int D2H read sizes them)."""
    out = _call("detr_proposals")(pred_logits, pred_boxes, label_map, sub2super, num_classes, topk_cat, feature_size, float(nms_thresh))
    off_host = out[0].numpy()
    d = dict(zip(_PROP_KEYS, out[1:1 + len(_PROP_KEYS)]))
    d.update(n=int(off_host[-1]), offsets_host=off_host, supers=out[-1] if sub2super is not None else None)
    return d


_MOC_KEYS = ("offsets", "cats", "conf", "box_i", "src", "box_img")


@_op("match_object_categories", "(Tensor prop_cats, Tensor prop_conf, Tensor prop_box, Tensor prop_offsets, Tensor gt_box, "
     "Tensor gt_offsets, Tensor? sub2super, int num_classes, int feature_size) -> Tensor[]")
def _match_object_categories(prop_cats, prop_conf, prop_box, prop_offsets, gt_box, gt_offsets, sub2super, num_classes, feature_size):
    d = _A.match_object_categories(prop_cats, prop_conf, prop_box, prop_offsets, gt_box, gt_offsets, sub2super, num_classes, feature_size)
    if d is None:
        return []
    out = [torch.from_numpy(d["offsets_host"])] + [d[k] for k in _MOC_KEYS]
    return out + ([d["supers"]] if d["supers"] is not None else [])


def match_object_categories(prop_cats, prop_conf, prop_box, prop_offsets, gt_box, gt_offsets, sub2super=None, num_classes=150,
                            feature_size=32):
    """utils.py:376-422 on CSR device arrays.  Returns None when the reference would return (None, None, None)."""
    out = _call("match_object_categories")(prop_cats, prop_conf, prop_box, prop_offsets, gt_box, gt_offsets, sub2super, num_classes,
                                           feature_size)
    if len(out) == 0:
        return None
    off_host = out[0].numpy()
    d = dict(zip(_MOC_KEYS, out[1:1 + len(_MOC_KEYS)]))
    d.update(n=int(off_host[-1]), offsets_host=off_host, supers=out[-1] if sub2super is not None else None)
    return d


@_op("targets_flat", "(Tensor dir_tri, Tensor rel_tri, Tensor tri_offsets, Tensor box_offsets) -> (Tensor, Tensor, Tensor, Tensor)")
def _targets_flat(dir_tri, rel_tri, tri_offsets, box_offsets):
    return _A.targets_flat(dir_tri, rel_tri, tri_offsets, box_offsets)


def targets_flat(dir_tri, rel_tri, tri_offsets, box_offsets):
    """utils.py:294-352 on the packed triangle arrays -> (gt_offsets [B+1], label, sub, obj) device arrays."""
    return _call("targets_flat")(dir_tri, rel_tri, tri_offsets, box_offsets)


# ------------------------------------------------------------------------------------------------ R12 across ranks
@_op("counts_allreduce", "(Tensor(a!) counters, int nccl_comm) -> ()")
def _counts_allreduce(counters, nccl_comm):
    from . import _lib
    if counters.dtype != torch.int64 or not counters.is_contiguous():
        raise RuntimeError("hiercom_b200: counters must be a contiguous int64 tensor")
    _lib.check(_lib.load().hc_counts_allreduce(nccl_comm, counters.data_ptr(), counters.numel(), _lib.stream_ptr()), "hc_counts_allreduce")
    _A._count()


def counts_allreduce(counters, nccl_comm):
    """In-place ncclAllReduce(sum, int64) of the counter vector on the current stream (hc_counts_allreduce); `nccl_comm` is the
    ncclComm_t as an integer (dist.CounterComm owns one per process)."""
    _call("counts_allreduce")(counters, int(nccl_comm))
    return counters
