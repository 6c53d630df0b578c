"""Tensor-level wrappers over the C ABI (one function per entry point of include/hiercom_b200.h).

PyTorch is plumbing here: it owns device memory and the stream; every computation below is a hand-written sm_100a
kernel reached through ctypes.  `LAUNCHES` counts kernel launches issued through this module (bench.py reports it).
"""
import contextlib
import ctypes as C
import time

import numpy as np
import torch

from . import _lib, tables
from ._lib import (ACT_NONE, ACT_RELU, ACT_TANH, EPI_BF16, EPI_F32, EPI_POOL_BF16, EPI_POOL_DIFF_BF16, EPI_SPLIT3_BF16, GEMM_CONV3,  # noqa: F401
                   GEMM_CONV3_BLOCKS, GEMM_PLAIN, check, ptr, require_cuda, stream_ptr)

LAUNCHES = {"n": 0}
PROFILE = {"on": False, "events": []}     # bench.py: CUDA-event timing of tagged launches on the launching stream


ACT_DTYPES = (torch.bfloat16, torch.float16)       # 16-bit operand formats of the tensor-core path (bf16 default; fp16 opt-in)


def _f16(*tensors):
    """1 when the 16-bit operands are fp16, 0 for bf16; all given tensors must share one of the two formats."""
    dts = {t.dtype for t in tensors if t is not None}
    if len(dts) != 1 or next(iter(dts)) not in ACT_DTYPES:
        raise RuntimeError("hiercom_b200: 16-bit operands must all be bf16 or all fp16, got %s" % sorted(str(d) for d in dts))
    return 1 if next(iter(dts)) == torch.float16 else 0


def _count(n=1):
    LAUNCHES["n"] += n


SYNC_WAIT = {"s": 0.0, "n": 0}       # host seconds spent blocked in device -> host reads on the step path (bench.py reports it)


@contextlib.contextmanager
def _sync_wait():
    t0 = time.perf_counter()
    try:
        yield
    finally:
        SYNC_WAIT["s"] += time.perf_counter() - t0
        SYNC_WAIT["n"] += 1


class _timed:
    """Brackets one launch with CUDA events on the current stream when PROFILE['on'] (no host sync)."""

    def __init__(self, tag):
        self.tag = tag

    def __enter__(self):
        if PROFILE["on"]:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *a):
        if PROFILE["on"]:
            self.e1.record()
            PROFILE["events"].append((self.tag, self.e0, self.e1))
        return False


def cs_bitmap_build(aligned_keys, violated_keys):
    """Host: packed key arrays -> uint32[BITMAP_WORDS] pass bitmap (numpy)."""
    al = np.ascontiguousarray(np.asarray(aligned_keys, dtype=np.int64))
    vi = np.ascontiguousarray(np.asarray(violated_keys, dtype=np.int64))
    out = np.zeros(tables.BITMAP_WORDS, dtype=np.uint32)
    check(_lib.load().hc_cs_bitmap_build(ptr(al) if al.size else None, al.size, ptr(vi) if vi.size else None, vi.size,
                                         out.ctypes.data), "hc_cs_bitmap_build")
    return out


def pairs_enumerate(boxes, box_offsets, tri_offsets, p_max, rel_tri=None, dir_tri=None, group_id=None, n_groups=0,
                    max_tri=0, feature_size=32, offsets_host=None):
    """R1/R2/R4.  `p_max` = sum_i N_i(N_i-1) (host int, known from the CSR the caller built).  Returns a dict of
    device arrays trimmed to the number of surviving directed pairs.  The per-image pair offsets size the later launches on the
    host: they are read back from the device (one small D2H sync) unless the caller already counted them (`offsets_host`, int32
    [B+1], pipeline.host_pair_offsets) - then nothing is read back and the call returns without synchronising."""
    require_cuda(boxes, box_offsets, tri_offsets, rel_tri, dir_tri, group_id)
    dev = boxes.device
    n_images = box_offsets.numel() - 1
    ws_ov = torch.empty(max(p_max // 2, 1), dtype=torch.uint8, device=dev)
    ws_any = torch.empty(max(n_groups * max_tri, 1), dtype=torch.uint8, device=dev) if group_id is not None else None
    ws_counts = torch.empty(n_images, dtype=torch.int32, device=dev)
    pair_offsets = torch.empty(n_images + 1, dtype=torch.int32, device=dev)
    i32 = lambda: torch.empty(max(p_max, 1), dtype=torch.int32, device=dev)
    pair_sub, pair_obj, pair_img, pair_gt, pair_rel = i32(), i32(), i32(), i32(), i32()
    pair_ov = torch.empty(max(p_max, 1), dtype=torch.uint8, device=dev)
    total = torch.zeros(1, dtype=torch.int32, device=dev)
    check(_lib.load().hc_pairs_enumerate(ptr(boxes), ptr(box_offsets), n_images, ptr(group_id), n_groups, max_tri,
                                         ptr(rel_tri), ptr(dir_tri), ptr(tri_offsets), feature_size, ptr(ws_ov),
                                         ptr(ws_any), ptr(ws_counts), ptr(pair_offsets), ptr(pair_sub), ptr(pair_obj),
                                         ptr(pair_img), ptr(pair_ov), ptr(pair_gt), ptr(pair_rel), ptr(total),
                                         stream_ptr()), "hc_pairs_enumerate")
    _count(4)
    if offsets_host is None:
        with _sync_wait():
            offsets_host = pair_offsets.cpu().numpy()     # [B+1] ints: the one D2H sync of the step (sizes the GEMM launches)
    else:
        offsets_host = np.asarray(offsets_host)
        if offsets_host.shape != (n_images + 1,):
            raise RuntimeError("hiercom_b200: pairs_enumerate: offsets_host must hold n_images + 1 entries")
    n = int(offsets_host[-1])
    return dict(n=n, offsets=pair_offsets, offsets_host=offsets_host, sub=pair_sub[:n], obj=pair_obj[:n], img=pair_img[:n], ov=pair_ov[:n],
                gt=pair_gt[:n], rel=pair_rel[:n])


def tc_gemm(a, b, out, m, n, k, *, bias=None, lda=0, ldc=None, c_off=0, mode=GEMM_PLAIN, epilogue=EPI_BF16,
            act=ACT_NONE, n_img=0, h=0, w=0, c_total=0, c_base=0, c_in=0, group_m=0, m_sub=0, tag="tc_gemm", mul=None,
            blocks=None, n_blocks=None, block_rows=0, block_cols=0, k_masks=None, k_cell=0, add_a=None, add_a_rows=None, add_b=None, add_b_rows=None,
            out_rows=None, diff_sub=None, diff_obj=None, diff_bg=None, pair_sub=None, pair_obj=None, pair_row=None, cta_pairs=0, scratch=None,
            m_order=None):
    """out = epilogue(A @ B^T) on tcgen05 (see include/hiercom_b200.h hc_tc_gemm)."""
    require_cuda(a, b, out, bias, mul, blocks, n_blocks, k_masks, add_a, add_a_rows, add_b, add_b_rows, out_rows, diff_sub, diff_obj, diff_bg,
                 pair_sub, pair_obj, pair_row, scratch)
    for t, dt in ((k_masks, torch.int64), (add_a, torch.float32), (add_b, torch.float32), (add_a_rows, torch.int32), (add_b_rows, torch.int32),
                  (out_rows, torch.int32), (diff_sub, a.dtype), (diff_obj, a.dtype), (diff_bg, a.dtype), (scratch, a.dtype),
                  (pair_sub, torch.int32), (pair_obj, torch.int32), (pair_row, torch.int32)):
        if t is not None and (t.dtype != dt or not t.is_contiguous()):
            raise RuntimeError("hiercom_b200: tc_gemm side operand must be a contiguous %s tensor" % dt)
    if add_a is not None and (add_b is None or add_a.stride(0) != add_b.stride(0)):
        raise RuntimeError("hiercom_b200: tc_gemm add_a / add_b must come together with the same row stride")
    d = _lib.GemmDesc()
    d.a, d.b, d.bias, d.out = ptr(a), ptr(b), ptr(bias), ptr(out)
    d.m, d.n, d.k = m, n, k
    d.lda, d.ldc, d.c_off = lda, (ldc if ldc is not None else n), c_off
    d.mode, d.epilogue, d.act = mode, epilogue, act
    d.n_img, d.h, d.w, d.c_total, d.c_base, d.c_in = n_img, h, w, c_total, c_base, c_in
    d.group_m, d.m_sub = group_m, m_sub
    d.mul, d.ld_mul = ptr(mul), (mul.stride(0) if mul is not None else 0)
    d.blocks, d.n_blocks, d.block_rows, d.block_cols = ptr(blocks), ptr(n_blocks), block_rows, block_cols
    d.k_masks, d.k_cell = ptr(k_masks), k_cell
    d.add_a, d.add_a_rows, d.add_b, d.add_b_rows = ptr(add_a), ptr(add_a_rows), ptr(add_b), ptr(add_b_rows)
    d.ld_add = add_a.stride(0) if add_a is not None else 0
    d.out_rows = ptr(out_rows)
    d.diff_sub, d.diff_obj, d.diff_bg = ptr(diff_sub), ptr(diff_obj), ptr(diff_bg)
    d.pair_sub, d.pair_obj, d.pair_row = ptr(pair_sub), ptr(pair_obj), ptr(pair_row)
    d.cta_pairs, d.scratch = int(cta_pairs), ptr(scratch)
    if m_order is not None:
        tile_rows = 128 * max(int(m_sub), 1) * (2 if cta_pairs else 1)
        if m_order.dtype != torch.int32 or not m_order.is_contiguous() or not m_order.is_cuda or m_order.numel() != -(-m // tile_rows):
            raise RuntimeError("hiercom_b200: tc_gemm m_order must be a contiguous CUDA int32 permutation of the %d-row M tiles" % tile_rows)
    d.m_order = ptr(m_order)
    d.operand_f16 = _f16(a, b)
    if d.operand_f16 and epilogue in (EPI_BF16, EPI_POOL_BF16, EPI_POOL_DIFF_BF16) and out.dtype != torch.float16:
        raise RuntimeError("hiercom_b200: fp16 operands write fp16 outputs")
    if scratch is not None and diff_sub is not None and scratch.numel() < n_img * (h // 2) * (w // 2) * d.ldc:
        raise RuntimeError("hiercom_b200: tc_gemm scratch must hold n_img pooled maps")
    with _timed(tag):
        check(_lib.load().hc_tc_gemm(C.byref(d), stream_ptr()), "hc_tc_gemm")
    _count()
    return out


def conv3_active_blocks(boxes, pair_sub, pair_obj, block_rows=8, fs=32, blocks=None, n_blocks=None, shared=False, block_cols=8):
    """Work list of the block-sparse conv3_1 for the given directed pairs (include/hiercom_b200.h hc_conv3_active_blocks;
    shared=True: hc_conv3_shared_blocks, only the cells both boxes reach).
    Returns (blocks int32 [n_pairs * 32 / block_rows], n_blocks int32 [1]); both stay on the device."""
    require_cuda(boxes, pair_sub, pair_obj, blocks, n_blocks)
    n = pair_sub.numel()
    cap = max(n * (256 // (block_rows * (block_cols or 8))), 1)
    if blocks is None:
        blocks = torch.empty(cap, dtype=torch.int32, device=boxes.device)
    if n_blocks is None:
        n_blocks = torch.empty(1, dtype=torch.int32, device=boxes.device)
    if blocks.numel() < cap:
        raise RuntimeError("hiercom_b200: conv3_active_blocks needs room for %d work-list entries" % cap)
    name = "hc_conv3_shared_blocks" if shared else "hc_conv3_active_blocks"
    check(getattr(_lib.load(), name)(ptr(boxes), ptr(pair_sub), ptr(pair_obj), n, fs, block_rows, block_cols, ptr(blocks), ptr(n_blocks), stream_ptr()),
          name)
    _count()
    return blocks, n_blocks


def conv2_box_blocks(boxes, block_rows=4, fs=32, blocks=None, n_blocks=None):
    """Work list of the conv2_1 halves restricted to each box's footprint (include/hiercom_b200.h hc_conv2_box_blocks)."""
    require_cuda(boxes, blocks, n_blocks)
    n_box = boxes.shape[0]
    cap = max(n_box * 4 * (32 // block_rows), 1)
    if blocks is None:
        blocks = torch.empty(cap, dtype=torch.int32, device=boxes.device)
    if n_blocks is None:
        n_blocks = torch.empty(1, dtype=torch.int32, device=boxes.device)
    if blocks.numel() < cap:
        raise RuntimeError("hiercom_b200: conv2_box_blocks needs room for %d work-list entries" % cap)
    check(_lib.load().hc_conv2_box_blocks(ptr(boxes), n_box, fs, block_rows, ptr(blocks), ptr(n_blocks), stream_ptr()), "hc_conv2_box_blocks")
    _count()
    return blocks, n_blocks


def pair_cell_keys(boxes, pair_sub, pair_obj, fs=32):
    """Sort key (int32 [n]) of the cell rectangle both boxes of each directed pair reach (include/hiercom_b200.h hc_pair_cell_keys)."""
    require_cuda(boxes, pair_sub, pair_obj)
    n = pair_sub.numel()
    keys = torch.empty(n, dtype=torch.int32, device=boxes.device)
    check(_lib.load().hc_pair_cell_keys(ptr(boxes), ptr(pair_sub), ptr(pair_obj), n, fs, ptr(keys), stream_ptr()), "hc_pair_cell_keys")
    _count()
    return keys


def tile_cell_masks(boxes, row_sub, row_obj, rows_per_tile, fs=32):
    """int64 [ceil(n / rows_per_tile)] bitmaps: union over each tile's rows of the cells both boxes reach (hc_tile_cell_masks)."""
    require_cuda(boxes, row_sub, row_obj)
    n = row_sub.numel()
    masks = torch.zeros(max(-(-n // rows_per_tile), 1), dtype=torch.int64, device=boxes.device)
    check(_lib.load().hc_tile_cell_masks(ptr(boxes), ptr(row_sub), ptr(row_obj), n, fs, rows_per_tile, ptr(masks), stream_ptr()),
          "hc_tile_cell_masks")
    _count()
    return masks


def cells_zero(masks, rows_per_tile, n_rows, out):
    """Zero, in every row of out [>= n_rows, n_cells, cell], the cells its tile's mask visits (hc_cells_zero)."""
    require_cuda(masks, out)
    if out.dim() != 3 or not out.is_contiguous() or out.shape[0] < n_rows or masks.dtype != torch.int64 or \
            masks.numel() * rows_per_tile < n_rows:
        raise RuntimeError("hiercom_b200: cells_zero needs a contiguous [rows, cells, cell] operand and one int64 mask per tile")
    with _timed("d_zero"):
        check(_lib.load().hc_cells_zero(ptr(masks), rows_per_tile, n_rows, out.shape[1], out.shape[2] * out.element_size(), ptr(out),
                                        stream_ptr()), "hc_cells_zero")
    _count()
    return out


def p3_assemble(background, sub_maps, obj_maps, boxes, pair_sub, pair_obj, out, fs=32):
    """Pooled conv3_1 output of every pair outside the cells both boxes reach (include/hiercom_b200.h hc_p3_assemble)."""
    require_cuda(background, sub_maps, obj_maps, boxes, pair_sub, pair_obj, out)
    n = pair_sub.numel()
    cell_map = 8 * 8 * 1024
    if (background.numel() != cell_map or sub_maps.numel() % cell_map or obj_maps.numel() != sub_maps.numel() or out.numel() < n * cell_map
            or any(t.dtype != background.dtype or t.dtype not in ACT_DTYPES or not t.is_contiguous() for t in (background, sub_maps, obj_maps, out))
            or sub_maps.numel() // cell_map < boxes.shape[0]):
        raise RuntimeError("hiercom_b200: p3_assemble needs contiguous bf16 [*,8,8,1024] maps (one per box) and room for n_pairs rows")
    with _timed("p3_fill"):
        check(_lib.load().hc_p3_assemble(ptr(background), ptr(sub_maps), ptr(obj_maps), ptr(boxes), ptr(pair_sub), ptr(pair_obj), n, fs,
                                         ptr(out), stream_ptr()), "hc_p3_assemble")
    _count()
    return out


def broadcast_rows(src, n_rows, out):
    """out[i] = src for i < n_rows (the background pre-fill of the pooled conv3_1 output)."""
    require_cuda(src, out)
    row_bytes = src.numel() * src.element_size()
    if out.numel() * out.element_size() < n_rows * row_bytes or not src.is_contiguous() or not out.is_contiguous():
        raise RuntimeError("hiercom_b200: broadcast_rows needs contiguous operands and room for n_rows copies")
    with _timed("p3_fill"):
        check(_lib.load().hc_broadcast_rows(ptr(src), row_bytes, n_rows, ptr(out), stream_ptr()), "hc_broadcast_rows")
    _count()
    return out


def pack_pixels(src0, src1, k_pad, out=None, dtype=torch.bfloat16):
    """[B,C0,H,W] (+[B,C1,H,W]) f32 -> [B*H*W, k_pad] in the 16-bit operand format `dtype` (bf16 / fp16)."""
    require_cuda(src0, src1)
    src0 = src0.contiguous()
    b, c0 = src0.shape[0], src0.shape[1]
    hw = src0.shape[2] * src0.shape[3]
    c1 = 0
    if src1 is not None:
        src1 = src1.contiguous()
        c1 = src1.shape[1]
    if out is None:
        out = torch.empty(b * hw, k_pad, dtype=dtype, device=src0.device)
    check(_lib.load().hc_pack_pixels(ptr(src0), c0, ptr(src1), c1, b, hw, k_pad, ptr(out), _f16(out), stream_ptr()), "hc_pack_pixels")
    _count()
    return out


def box_select(t_img, boxes, box_img, fill, fs=32, out=None):
    require_cuda(t_img, boxes, box_img, fill)
    n_box, ch = boxes.shape[0], t_img.shape[-1]
    if out is None:
        out = torch.empty(n_box, fs, fs, ch, dtype=t_img.dtype, device=t_img.device)
    _f16(t_img, fill, out)
    check(_lib.load().hc_box_select(ptr(t_img), ptr(boxes), ptr(box_img), n_box, fs, ch, ptr(fill), ptr(out), stream_ptr()),
          "hc_box_select")
    _count()
    return out


def _footprint(fp, u):
    """fp = (boxes int32 [n_box,4], u_bg, v_bg [1,fs,fs,C] in u's format, conv2 block rows) or None -> (struct kept alive, pointer or None)."""
    if fp is None:
        return None, None
    boxes, u_bg, v_bg, rows = fp
    require_cuda(boxes, u_bg, v_bg)
    if (boxes.dtype != torch.int32 or not boxes.is_contiguous() or boxes.shape[0] < u.shape[0] or u_bg.dtype != u.dtype or v_bg.dtype != u.dtype
            or not u_bg.is_contiguous() or not v_bg.is_contiguous() or u_bg.numel() != u[0].numel() or v_bg.numel() != u[0].numel()):
        raise RuntimeError("hiercom_b200: footprint maps need int32 boxes (one per map) and background maps of one map's shape and format")
    st = _lib.UVFootprint(ptr(boxes), ptr(u_bg), ptr(v_bg), int(rows))
    return st, C.byref(st)


def pair_relu_pool(u, v, bias, pair_sub, pair_obj, fs=32, out=None, cover=None, fp=None):
    require_cuda(u, v, bias, pair_sub, pair_obj, cover)
    if cover is not None and (cover.dtype != torch.int64 or not cover.is_contiguous() or cover.numel() < pair_sub.numel()):
        raise RuntimeError("hiercom_b200: pair_relu_pool cover must be a contiguous int64 tensor with one word per pair")
    n, ch = pair_sub.numel(), u.shape[-1]
    if out is None:
        out = torch.empty(n, fs // 2, fs // 2, ch, dtype=u.dtype, device=u.device)
    with _timed("pair_pool"):
        keep, fpp = _footprint(fp, u)
        check(_lib.load().hc_pair_relu_pool(ptr(u), ptr(v), ptr(bias), ptr(pair_sub), ptr(pair_obj), n, fs, ch, ptr(cover), fpp, ptr(out),
                                            _f16(u, v, out), stream_ptr()), "hc_pair_relu_pool")
    _count()
    return out


def pair_lut_build(pair_sub, pair_obj, pair_img, box_offsets, n_box, n_max):
    require_cuda(pair_sub, pair_obj, pair_img, box_offsets)
    lut = torch.empty(n_box, n_max, dtype=torch.int32, device=box_offsets.device)
    check(_lib.load().hc_pair_lut_build(ptr(pair_sub), ptr(pair_obj), ptr(pair_img), ptr(box_offsets), pair_sub.numel(), n_box, n_max,
                                        ptr(lut), stream_ptr()), "hc_pair_lut_build")
    _count(1)
    return lut


def pair_cover_masks(boxes, pair_sub, pair_obj, block_rows, block_cols, shared, fs=32, out=None):
    """int64 [n] bitmaps of the 8x8-grid cells the listed conv3_1 blocks of each pair cover (include/hiercom_b200.h hc_pair_cover_masks)."""
    require_cuda(boxes, pair_sub, pair_obj, out)
    n = pair_sub.numel()
    if out is None:
        out = torch.empty(max(n, 1), dtype=torch.int64, device=boxes.device)
    if out.dtype != torch.int64 or out.numel() < n or not out.is_contiguous():
        raise RuntimeError("hiercom_b200: pair_cover_masks needs a contiguous int64 output with one word per pair")
    check(_lib.load().hc_pair_cover_masks(ptr(boxes), ptr(pair_sub), ptr(pair_obj), n, fs, block_rows, block_cols, 1 if shared else 0,
                                          ptr(out), stream_ptr()), "hc_pair_cover_masks")
    _count()
    return out


def pair_relu_pool_tiled(u, v, bias, box_offsets, lut, img0, n_img, pair_base, chunk_pairs, fs=32, out=None, cover=None, fp=None):
    require_cuda(u, v, bias, box_offsets, lut, cover)
    if cover is not None and (cover.dtype != torch.int64 or cover.numel() < chunk_pairs or not cover.is_contiguous()):
        raise RuntimeError("hiercom_b200: pair_relu_pool_tiled cover must be a contiguous int64 tensor with one word per pair of the chunk")
    ch = u.shape[-1]
    if out is None:
        out = torch.empty(chunk_pairs, fs // 2, fs // 2, ch, dtype=u.dtype, device=u.device)
    with _timed("pair_pool"):
        keep, fpp = _footprint(fp, u)
        check(_lib.load().hc_pair_relu_pool_tiled(ptr(u), ptr(v), ptr(bias), ptr(box_offsets), ptr(lut), lut.shape[1], img0, n_img,
                                                  pair_base, chunk_pairs, fs, ch, ptr(cover), fpp, ptr(out), _f16(u, v, out), stream_ptr()),
              "hc_pair_relu_pool_tiled")
    _count()
    return out


def hier_head(fc2_raw, fc2_bias, emb, row_sub, row_obj, box_cat, box_super, w_heads, b_heads, splits, flat=False,
              temps=(1.0, 1.0, 1.0), num_obj=150, num_super=17, want_pred=False):
    require_cuda(fc2_raw, fc2_bias, emb, row_sub, row_obj, box_cat, box_super, w_heads, b_heads)
    n, hidden = fc2_raw.shape
    dev = fc2_raw.device
    r = sum(splits)
    relation = torch.empty(n, r, dtype=torch.float32, device=dev)
    sup = None if flat else torch.empty(n, 3, dtype=torch.float32, device=dev)
    conn = torch.empty(n, dtype=torch.float32, device=dev)
    logsig = torch.empty(n, dtype=torch.float32, device=dev)
    pred = torch.empty(n, hidden, dtype=torch.float32, device=dev) if want_pred else None
    box_emb = None
    if fc2_bias is not None:
        # label columns of fc2 summed once per box ([as subject | as object]); a pair then adds two rows instead of <= 10
        n_box = box_cat.numel()
        box_emb = torch.empty(n_box, 2 * hidden, dtype=torch.float32, device=dev)
        check(_lib.load().hc_box_label_embed(ptr(emb), num_obj, num_super, ptr(box_cat), ptr(box_super), n_box, hidden, ptr(box_emb),
                                             stream_ptr()), "hc_box_label_embed")
        _count()
    check(_lib.load().hc_hier_head(ptr(fc2_raw), fc2_raw.stride(0), n, hidden, ptr(fc2_bias), ptr(emb), num_obj, num_super,
                                   ptr(row_sub), ptr(row_obj), ptr(box_cat), ptr(box_super), ptr(w_heads), ptr(b_heads),
                                   splits[0], splits[1], splits[2], int(flat), temps[0], temps[1], temps[2], ptr(relation),
                                   ptr(sup), ptr(conn), ptr(logsig), ptr(pred), ptr(box_emb), stream_ptr()), "hc_hier_head")
    _count()
    return relation, sup, conn, logsig, pred


def candidates(relation, splits, hier, row_ov, logsig, row_sub, row_obj, box_cat, pass_bitmap=None, super_rel=None,
               conf_sub=None, conf_obj=None, layout=0, want_top3=False):
    require_cuda(relation, row_ov, logsig, row_sub, row_obj, box_cat, pass_bitmap, super_rel, conf_sub, conf_obj)
    n = relation.shape[0]
    dev = relation.device
    k = 3 if hier else 1
    cand_conf = torch.empty(n * k, dtype=torch.float32, device=dev)
    cand_label = torch.empty(n * k, dtype=torch.int32, device=dev)
    t3_conf = torch.empty(n, dtype=torch.float32, device=dev) if want_top3 else None
    t3_super = torch.empty(n, dtype=torch.uint8, device=dev) if want_top3 else None
    check(_lib.load().hc_candidates(ptr(relation), relation.stride(0), n, splits[0], splits[1], splits[2], int(hier),
                                    ptr(row_ov), ptr(logsig), ptr(conf_sub), ptr(conf_obj), ptr(row_sub), ptr(row_obj),
                                    ptr(box_cat), ptr(pass_bitmap), ptr(super_rel), ptr(cand_conf), ptr(cand_label),
                                    ptr(t3_conf), ptr(t3_super), layout, stream_ptr()), "hc_candidates")
    _count()
    return cand_conf, cand_label, t3_conf, t3_super


def topk_match(cand_offsets, cand_conf, cand_label, k_per_row, row_sub, row_obj, pred_cat, pred_box, gt_offsets, gt_label,
               gt_sub, gt_obj, gt_cat, gt_box, counters, *, cand_row=None, synonyms=None, zs_bitmap=None, mode=0,
               t3_labels=None, t3_super=None, feature_size=32, iou_thresh=0.5, top_k=tables.TOP_K, want_topk=False):
    require_cuda(cand_offsets, cand_conf, cand_label, row_sub, row_obj, pred_cat, pred_box, gt_offsets, gt_label, gt_sub,
                 gt_obj, gt_cat, gt_box, counters, cand_row, synonyms, zs_bitmap, t3_labels, t3_super)
    n_images = cand_offsets.numel() - 1
    top_max = int(top_k[-1])
    topk_out = torch.empty(n_images, top_max, dtype=torch.int32, device=cand_conf.device) if want_topk else None
    check(_lib.load().hc_topk_match(ptr(cand_offsets), n_images, ptr(cand_conf), ptr(cand_label), ptr(cand_row), k_per_row,
                                    ptr(row_sub), ptr(row_obj), ptr(pred_cat), ptr(pred_box), ptr(gt_offsets), ptr(gt_label),
                                    ptr(gt_sub), ptr(gt_obj), ptr(gt_cat), ptr(gt_box), ptr(synonyms), tables.NUM_OBJ,
                                    tables.NUM_PRED, ptr(zs_bitmap), feature_size, float(iou_thresh), top_max, int(top_k[0]),
                                    int(top_k[1]), int(top_k[2]), mode, ptr(t3_labels), ptr(t3_super), ptr(counters),
                                    ptr(topk_out), stream_ptr()), "hc_topk_match")
    _count()
    return topk_out


def topk_select(cand_offsets, cand_conf, top_max=128):
    require_cuda(cand_offsets, cand_conf)
    n_images = cand_offsets.numel() - 1
    out = torch.empty(n_images, top_max, dtype=torch.int32, device=cand_conf.device)
    check(_lib.load().hc_topk_select(ptr(cand_offsets), n_images, ptr(cand_conf), top_max, ptr(out), stream_ptr()), "hc_topk_select")
    _count()
    return out


def connectivity_stats(connectivity, gt_directed, gt_undirected, stats):
    require_cuda(connectivity, gt_directed, gt_undirected, stats)
    check(_lib.load().hc_connectivity_stats(ptr(connectivity), ptr(gt_directed), ptr(gt_undirected), connectivity.numel(),
                                            ptr(stats), stream_ptr()), "hc_connectivity_stats")
    _count()


# ------------------------------------------------------------------------------------------------------ SGB twin (R14/N1)
def sgb_pair_gather(edge_rep, pair_idx, hidden, split=False, f16=False):
    """split: the bf16x3 [hi | lo | hi] layout; f16: plain fp16 rows (the fp16 operand format of tc_gemm) instead of plain bf16."""
    require_cuda(edge_rep, pair_idx)
    n = pair_idx.shape[0]
    if split and f16:
        raise RuntimeError("hiercom_b200: sgb_pair_gather: the split layout is bf16-only")
    out = torch.empty(n, (3 if split else 1) * 2 * hidden, dtype=torch.float16 if f16 else torch.bfloat16, device=edge_rep.device)
    check(_lib.load().hc_sgb_pair_gather(ptr(edge_rep), ptr(pair_idx), n, hidden, 2 if f16 else int(split), ptr(out), stream_ptr()),
          "hc_sgb_pair_gather")
    _count()
    return out


def split_bf16x3(x):
    """f32 [n,k] -> bf16 [n,3k] = [hi | lo | hi] (A side of the bf16x3 scheme)."""
    require_cuda(x)
    n, k = x.shape
    out = torch.empty(n, 3 * k, dtype=torch.bfloat16, device=x.device)
    check(_lib.load().hc_split_bf16x3(ptr(x), x.stride(0), n, k, ptr(out), stream_ptr()), "hc_split_bf16x3")
    _count()
    return out


def pack_weight_bf16x3(w):
    """f32 [n,k] weight -> bf16 [n,3k] = [W_hi | W_hi | W_lo] (B side of the bf16x3 scheme); one-time packing (torch)."""
    w = w.detach().float()
    hi = w.to(torch.bfloat16)
    lo = (w - hi.float()).to(torch.bfloat16)
    return torch.cat((hi, hi, lo), dim=1).contiguous()


def sgb_hier_softmax(logits, splits, bias_table=None, num_obj=151, pair_pred=None, label_ids=None):
    require_cuda(logits, bias_table, pair_pred, label_ids)
    n, r = logits.shape[0], sum(splits)
    rel = torch.empty(n, r, dtype=torch.float32, device=logits.device)
    sup = torch.empty(n, 4, dtype=torch.float32, device=logits.device)
    check(_lib.load().hc_sgb_hier_softmax(ptr(logits), logits.stride(0), n, splits[0], splits[1], splits[2], ptr(bias_table), num_obj,
                                          ptr(pair_pred), ptr(label_ids), ptr(rel), ptr(sup), stream_ptr()), "hc_sgb_hier_softmax")
    _count()
    return rel, sup


def sgb_candidates(rel, splits, pair_offsets, pair_img, pair_idx, obj_scores, label_ids):
    require_cuda(rel, pair_offsets, pair_img, pair_idx, obj_scores, label_ids)
    n = rel.shape[0]
    dev = rel.device
    score = torch.empty(3 * n, dtype=torch.float32, device=dev)
    label = torch.empty(3 * n, dtype=torch.int32, device=dev)
    row = torch.empty(3 * n, dtype=torch.int32, device=dev)
    check(_lib.load().hc_sgb_candidates(ptr(rel), splits[0], splits[1], splits[2], ptr(pair_offsets), ptr(pair_img), ptr(pair_idx),
                                        ptr(obj_scores), ptr(label_ids), n, ptr(score), ptr(label), ptr(row), stream_ptr()),
          "hc_sgb_candidates")
    _count()
    return score, label, row


def sgb_rank_match(ranked, reject, pair_offsets, cand_score, cand_label, cand_row, pair_idx, pred_cls, pred_box, gt_offsets, gt_rel,
                   gt_cls, gt_box, iou_thresh=0.5, top_k=(20, 50, 100)):
    require_cuda(ranked, reject, pair_offsets, cand_score, cand_label, cand_row, pair_idx, pred_cls, pred_box, gt_offsets, gt_rel, gt_cls,
                 gt_box)
    n_img = pair_offsets.numel() - 1
    dev = cand_score.device
    top_max = int(top_k[-1])
    i32 = lambda *s: torch.zeros(*s, dtype=torch.int32, device=dev)
    final_rank, hits, ngt, hits_pc, cnt_pc = i32(n_img, top_max), i32(n_img, 3), i32(n_img), i32(n_img, 3, 51), i32(n_img, 51)
    check(_lib.load().hc_sgb_rank_match(ptr(ranked), ptr(reject), ptr(pair_offsets), n_img, ptr(cand_score), ptr(cand_label),
                                        ptr(cand_row), ptr(pair_idx), ptr(pred_cls), ptr(pred_box), ptr(gt_offsets), ptr(gt_rel),
                                        ptr(gt_cls), ptr(gt_box), float(iou_thresh), top_max, int(top_k[0]), int(top_k[1]),
                                        int(top_k[2]), ptr(final_rank), ptr(hits), ptr(ngt), ptr(hits_pc), ptr(cnt_pc), stream_ptr()),
          "hc_sgb_rank_match")
    _count()
    return final_rank, hits, ngt, hits_pc, cnt_pc


# ------------------------------------------------------------------------------------------ proposal front-end (N2)
def detr_proposals(pred_logits, pred_boxes, label_map, sub2super=None, num_classes=150, topk_cat=2, feature_size=32,
                   nms_thresh=0.5):
    """evaluate.py:311-370.  Returns CSR device arrays of the surviving proposals (one [B+1]-int D2H read sizes them)."""
    require_cuda(pred_logits, pred_boxes, label_map, sub2super)
    pred_logits, pred_boxes = pred_logits.contiguous().float(), pred_boxes.contiguous().float()
    b, q = pred_logits.shape[0], pred_logits.shape[1]
    if pred_logits.shape[2] != num_classes + 1 or tuple(pred_boxes.shape) != (b, q, 4):
        raise RuntimeError("hiercom_b200 detr_proposals: pred_logits must be [B,Q,num_classes+1] and pred_boxes [B,Q,4]")
    dev = pred_logits.device
    e = q * topk_cat
    i32 = lambda *s: torch.empty(*s, dtype=torch.int32, device=dev)
    f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    ws_label, ws_conf, ws_box, ws_valid = i32(b * e), f32(b * e), f32(b * e, 4), torch.empty(b * e, dtype=torch.uint8, device=dev)
    st_label, st_conf, st_box, st_count, offsets = i32(b * e), f32(b * e), f32(b * e, 4), i32(b), i32(b + 1)
    check(_lib.load().hc_detr_proposals(ptr(pred_logits), ptr(pred_boxes), b, q, num_classes, topk_cat, ptr(label_map), feature_size,
                                        float(nms_thresh), ptr(ws_label), ptr(ws_conf), ptr(ws_box), ptr(ws_valid), ptr(st_label),
                                        ptr(st_conf), ptr(st_box), ptr(st_count), ptr(offsets), stream_ptr()), "hc_detr_proposals")
    _count(3)
    offsets_host = offsets.cpu().numpy()
    n = int(offsets_host[-1])
    cats, conf, box_f, box_i = i32(max(n, 1)), f32(max(n, 1)), f32(max(n, 1), 4), i32(max(n, 1), 4)
    supers = torch.empty(max(n, 1), 4, dtype=torch.int8, device=dev) if sub2super is not None else None
    box_img = i32(max(n, 1))
    check(_lib.load().hc_proposals_pack(ptr(st_label), ptr(st_conf), ptr(st_box), ptr(offsets), b, e, ptr(sub2super), num_classes,
                                        ptr(cats), ptr(conf), ptr(box_f), ptr(box_i), ptr(supers), ptr(box_img), stream_ptr()),
          "hc_proposals_pack")
    _count()
    return dict(n=n, offsets=offsets, offsets_host=offsets_host, cats=cats[:n], conf=conf[:n], box_f=box_f[:n], box_i=box_i[:n],
                supers=None if supers is None else supers[:n], box_img=box_img[:n])


def match_object_categories(prop_cats, prop_conf, prop_box, prop_offsets, gt_box, gt_offsets, sub2super=None, num_classes=150,
                            feature_size=32):
    """utils.py:376-422 on CSR device arrays.  Returns None when the reference would return (None, None, None)."""
    require_cuda(prop_cats, prop_conf, prop_box, prop_offsets, gt_box, gt_offsets, sub2super)
    dev = prop_box.device
    b = prop_offsets.numel() - 1
    n_gt = gt_box.shape[0]
    i32 = lambda *s: torch.empty(*s, dtype=torch.int32, device=dev)
    ws_idx, ws_iou = i32(max(n_gt, 1), 2), torch.empty(max(n_gt, 1), 2, dtype=torch.float32, device=dev)
    ws_count, out_offsets, status = i32(b), i32(b + 1), i32(1)
    check(_lib.load().hc_match_object_categories(ptr(prop_box), ptr(prop_offsets), ptr(gt_box), ptr(gt_offsets), b, feature_size,
                                                 ptr(ws_idx), ptr(ws_iou), ptr(ws_count), ptr(out_offsets), ptr(status), stream_ptr()),
          "hc_match_object_categories")
    _count(2)
    host = torch.cat((out_offsets, status)).cpu().numpy()        # one D2H read: sizes + the reference's "None" condition
    if int(host[-1]) != 0:
        return None
    offsets_host = host[:-1]
    n = int(offsets_host[-1])
    cats, conf, box, src = i32(max(n, 1)), torch.empty(max(n, 1), dtype=torch.float32, device=dev), i32(max(n, 1), 4), i32(max(n, 1))
    supers = torch.empty(max(n, 1), 4, dtype=torch.int8, device=dev) if sub2super is not None else None
    img = i32(max(n, 1))
    check(_lib.load().hc_match_object_categories_fill(ptr(prop_cats), ptr(prop_conf), ptr(prop_offsets), ptr(gt_box), ptr(gt_offsets), b,
                                                      ptr(ws_idx), ptr(ws_iou), ptr(out_offsets), ptr(sub2super), num_classes,
                                                      ptr(cats), ptr(conf), ptr(box), ptr(src), ptr(supers), ptr(img), stream_ptr()),
          "hc_match_object_categories_fill")
    _count()
    return dict(n=n, offsets=out_offsets, offsets_host=offsets_host, cats=cats[:n], conf=conf[:n], box_i=box[:n], src=src[:n],
                supers=None if supers is None else supers[:n], box_img=img[:n])


def targets_flat(dir_tri, rel_tri, tri_offsets, box_offsets):
    """utils.py:294-352 on the packed triangle arrays -> (gt_offsets [B+1], label, sub, obj) device arrays; label/sub/obj are
    over-allocated to sum T_i entries and valid up to gt_offsets[-1] (no host sync here)."""
    require_cuda(dir_tri, rel_tri, tri_offsets, box_offsets)
    dev = dir_tri.device
    b = box_offsets.numel() - 1
    cap = max(dir_tri.numel(), 1)
    i32 = lambda *s: torch.empty(*s, dtype=torch.int32, device=dev)
    ws_count, gt_offsets, label, sub, obj = i32(b), i32(b + 1), i32(cap), i32(cap), i32(cap)
    check(_lib.load().hc_targets_flat(ptr(dir_tri), ptr(rel_tri), ptr(tri_offsets), ptr(box_offsets), b, ptr(ws_count), ptr(gt_offsets),
                                      ptr(label), ptr(sub), ptr(obj), stream_ptr()), "hc_targets_flat")
    _count(3)
    return gt_offsets, label, sub, obj


def hier_loss(relation, super_rel, connectivity, row_target, group_offsets, group_rows, group_weight, class_weight, splits, hier=True,
              temps=(1.0, 1.0, 1.0), aligned_bitmap=None, violated_bitmap=None, row_sub=None, row_obj=None, box_cat=None,
              lambdas=(0.1, 1.0, 1.0, 0.1, 10.0), want_grad=True):
    """N4: per-call training losses + gradient with respect to the head's logits (include/hiercom_b200.h hc_hier_loss).
    lambdas = (connectivity, not_connected, commonsense, cs_weak, cs_strong).  Returns (group_loss [M,3], total [4], d_logits)."""
    require_cuda(relation, super_rel, connectivity, row_target, group_offsets, group_rows, group_weight, class_weight, aligned_bitmap,
                 violated_bitmap, row_sub, row_obj, box_cat)
    n = relation.shape[0]
    dev = relation.device
    m = group_offsets.numel() - 1
    n_out = sum(splits) + (4 if hier else 1)
    group_loss = torch.empty(m, 3, dtype=torch.float32, device=dev)
    total = torch.empty(4, dtype=torch.float32, device=dev)
    d_logits = torch.empty(n, n_out, dtype=torch.float32, device=dev) if want_grad else None
    check(_lib.load().hc_hier_loss(ptr(relation), relation.stride(0), ptr(super_rel), ptr(connectivity), n, splits[0], splits[1], splits[2],
                                   int(hier), temps[0], temps[1], temps[2], ptr(row_target), ptr(group_offsets), ptr(group_rows), m,
                                   ptr(group_weight), ptr(class_weight), ptr(aligned_bitmap), ptr(violated_bitmap), ptr(row_sub),
                                   ptr(row_obj), ptr(box_cat), lambdas[0], lambdas[1], lambdas[2], lambdas[3], lambdas[4],
                                   ptr(group_loss), ptr(total), ptr(d_logits), n_out, stream_ptr()), "hc_hier_loss")
    _count(2)
    return group_loss, total, d_logits


def hier_head_bwd(d_logits, pred, w_heads, scale=None, want_pred=True, want_weights=True):
    """Backward of fc3_x / fc4 / fc5: (d_pred [n,512], d_w [n_out,512], d_b [n_out]) (hc_hier_head_bwd)."""
    require_cuda(d_logits, pred, w_heads, scale)
    n, n_out = d_logits.shape
    dev = d_logits.device
    d_pred = torch.empty(n, pred.shape[1], dtype=torch.float32, device=dev) if want_pred else None
    d_w = d_b = ws = None
    parts = 0
    if want_weights:
        parts = max(1, min(148, (n + 63) // 64))
        d_w = torch.empty(n_out, pred.shape[1], dtype=torch.float32, device=dev)
        d_b = torch.empty(n_out, dtype=torch.float32, device=dev)
        ws = torch.empty(parts * n_out * (pred.shape[1] + 1), dtype=torch.float32, device=dev)
    check(_lib.load().hc_hier_head_bwd(ptr(d_logits), d_logits.stride(0), ptr(pred), pred.stride(0), n, n_out, ptr(w_heads), ptr(scale),
                                       ptr(d_pred), ptr(d_w), ptr(d_b), ptr(ws), parts, stream_ptr()), "hc_hier_head_bwd")
    _count(int(want_pred) + 2 * int(want_weights))
    return d_pred, d_w, d_b
