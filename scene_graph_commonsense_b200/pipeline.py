"""Batched entry point of the relation path: whole images in, integer Recall counters out.

Replaces the reference's per-(graph_iter, edge_iter, direction) host loop (evaluate.py:132-217: ~40 tiny launches and
several host syncs per directed pair) with ~20 launches per *batch*:

  pairs_enumerate                       R1 R2 R4   evaluate.py:111-116,132-156
  pack_pixels -> conv1 GEMM (+tanh)     R5         model.py:139-140, once per IMAGE (1x1 conv commutes with the 0/1 mask)
  box_select                            R3         train_test.py:391,398 (feature*mask never materialised)
  conv2 subject/object half convs       R5         model.py:143, once per BOX (conv2_1 is linear before the ReLU)
  per pair chunk: pair_relu_pool -> conv3+ReLU+pool -> fc1+ReLU -> fc2      model.py:143-150,175
  hier_head                             R6 R7      model.py:152-168,176-184
  candidates                            R8 R9      evaluator.py:157-179,231-266
  topk_match (Evaluator, Evaluator_Top3)  R10-R13  evaluator.py:294-356,704-766
  connectivity_stats                               train_utils.py:169-183

Data layout in HBM: CSR over images (box_offsets, tri_offsets, pair offsets); activations NHWC bf16 so the channel
index is the GEMM K index and a TMA box row; all counters int64 in one 765-slot vector (tables.EV_* / T3_*).
"""
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import ops, tables
from ._lib import ACT_NONE, ACT_TANH, EPI_BF16, GEMM_CONV3
from .model import K1_PAD, PackedHead


@dataclass
class DeviceBatch:
    """One evaluation window on the device.  Built from host samples by `from_samples` (pinned staging + async H2D)."""
    feat: torch.Tensor            # f32 [B,256,32,32]
    depth: torch.Tensor           # f32 [B,1,32,32]
    boxes: torch.Tensor           # int32 [nbox,4] (xmin,xmax,ymin,ymax), truncated toward zero
    box_offsets: torch.Tensor     # int32 [B+1]
    box_img: torch.Tensor         # int32 [nbox]
    cats: torch.Tensor            # int32 [nbox]
    supers: torch.Tensor          # int8  [nbox,4]
    tri_offsets: torch.Tensor     # int32 [B+1]
    rel_tri: Optional[torch.Tensor]   # int32 [sum T_i]  relationships[g-1][e] at t = g(g-1)/2+e
    dir_tri: Optional[torch.Tensor]   # int8  [sum T_i]  subj_or_obj[g-1][e]
    group_id: Optional[torch.Tensor]  # int32 [B] lock-step batch of each image ("batch" skip mode) or None
    n_groups: int
    max_tri: int
    p_max: int                    # sum_i N_i(N_i-1)
    h2d_bytes: int
    # SGDET/SGCLS extras
    conf: Optional[torch.Tensor] = None        # f32 [nbox] object-label confidences
    box_offsets_host: Optional[np.ndarray] = None
    gt: Optional[dict] = None                  # flat GT triplet tables (targets.flat_targets_sgd)
    cover_fraction: Optional[float] = None     # host estimate: share of the conv3_1 pixels the shared-footprint work lists would visit
    pair_offsets_host: Optional[np.ndarray] = None   # int32 [B+1] directed-pair offsets counted on the host (host_pair_offsets): no D2H per step

    @property
    def n_images(self):
        return self.box_offsets.numel() - 1


class HostBatch:
    """Pinned host staging of one evaluation window (what a data loader hands over): numpy/torch CPU buffers only.
    `to_device` issues the async H2D copies; `h2d_bytes` is the exact number of bytes they move."""

    def __init__(self, arrays, meta, pinned=True):
        self.meta = meta
        self.t = {}
        for k, v in arrays.items():
            if v is None:
                self.t[k] = None
                continue
            t = v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))
            self.t[k] = t.pin_memory() if pinned else t
        self.pinned = pinned
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.t.values() if t is not None)

    def to_device(self, device):
        d = {k: (None if t is None else t.to(device, non_blocking=self.pinned)) for k, t in self.t.items()}
        m = self.meta
        gt = None
        if d.get("gt_offsets") is not None:
            gt = dict(offsets=d["gt_offsets"], label=d["gt_label"], sub=d["gt_sub"], obj=d["gt_obj"], cat=d["gt_cat"], box=d["gt_box"])
        return DeviceBatch(d.get("feat"), d.get("depth"), d["boxes"], d["box_offsets"], d["box_img"], d["cats"], d["supers"],
                           d["tri_offsets"], d.get("rel_tri"), d.get("dir_tri"), d.get("group_id"), m["n_groups"], m["max_tri"],
                           m["p_max"], self.h2d_bytes, conf=d.get("conf"), gt=gt,
                           box_offsets_host=self.t["box_offsets"].numpy(), cover_fraction=m.get("cover_fraction"),
                           pair_offsets_host=m.get("pair_offsets"))


def _cell_interval(lo, hi):
    """Host twin of `active_cells` (csrc/blocks.cu): the pooled conv3_1 cells (8 per axis) a box interval [lo, hi) of the 32-grid reaches."""
    qlo, qhi = np.maximum(0, (lo - 1) >> 1), np.minimum(15, hi >> 1)
    qlo, qhi = np.maximum(0, qlo - 1), np.minimum(15, qhi + 1)
    return qlo >> 1, (qhi >> 1) + 1


def footprint_cover_fraction(boxes, box_offsets, fs=32):
    """Host only (numpy, at batch-build time): the share of the dense conv3_1 pixels that the shared-footprint work lists of this
    window would visit - per ordered pair of an image, the cell rectangle BOTH boxes reach covered by 2 x 2-cell blocks
    (ceil(w/2) * ceil(h/2) blocks of 4 of the 64 cells).  Ignores the overlap skip rule (an upper bound on the pairs).  It only
    steers a scheduling decision (`RelationPipeline.dense_above`): both formulations give the same scores."""
    b = np.clip(np.asarray(boxes, dtype=np.int64), 0, fs)
    empty = (b[:, 1] <= b[:, 0]) | (b[:, 3] <= b[:, 2])
    xa, xb = _cell_interval(b[:, 0], b[:, 1])
    ya, yb = _cell_interval(b[:, 2], b[:, 3])
    xa, xb, ya, yb = (np.where(empty, 0, v) for v in (xa, xb, ya, yb))
    blocks = pairs = 0
    for i in range(len(box_offsets) - 1):
        s, e = int(box_offsets[i]), int(box_offsets[i + 1])
        n = e - s
        if n < 2:
            continue
        w = np.maximum(0, np.minimum(xb[s:e, None], xb[None, s:e]) - np.maximum(xa[s:e, None], xa[None, s:e]))
        h = np.maximum(0, np.minimum(yb[s:e, None], yb[None, s:e]) - np.maximum(ya[s:e, None], ya[None, s:e]))
        t = ((w + 1) >> 1) * ((h + 1) >> 1)
        blocks += int(t.sum() - np.trace(t))
        pairs += n * (n - 1)
    return blocks * 4.0 / (64.0 * pairs) if pairs else 0.0


def host_pair_offsets(boxes, box_offsets, group_id=None, fs=32):
    """Host twin (numpy, at batch-build time) of the COUNTING half of hc_pairs_enumerate (R1/R2/R4, evaluate.py:111-116,132-156):
    the directed-pair CSR offsets int32 [B+1] the device will produce, so the step needs no device -> host read to size its launches.
    A pair (g, e) of image i is overlapping iff the two rectangles - Python slice semantics of the reference's masks - intersect;
    it survives iff it overlaps (`group_id` None: per-image rule) or overlaps in ANY image of i's lock-step group that has it (batch rule)."""
    b = np.asarray(boxes, dtype=np.int64).reshape(-1, 4)

    def sb(v):
        v = np.where(v < 0, np.maximum(v + fs, 0), v)
        return np.minimum(v, fs)

    x0, x1, y0, y1 = sb(b[:, 0]), sb(b[:, 1]), sb(b[:, 2]), sb(b[:, 3])
    x1, y1 = np.maximum(x1, x0), np.maximum(y1, y0)
    n_img = len(box_offsets) - 1
    ovs = []
    for i in range(n_img):
        s, e = int(box_offsets[i]), int(box_offsets[i + 1])
        g, l = np.tril_indices(e - s, -1)                    # t = g(g-1)/2 + e in the reference's loop order (g outer, e < g)
        w = np.minimum(x1[s + g], x1[s + l]) - np.maximum(x0[s + g], x0[s + l])
        h = np.minimum(y1[s + g], y1[s + l]) - np.maximum(y0[s + g], y0[s + l])
        ovs.append((w > 0) & (h > 0))
    counts = np.zeros(n_img, dtype=np.int64)
    if group_id is None:
        for i, ov in enumerate(ovs):
            counts[i] = int(ov.sum())
    else:
        gid = np.asarray(group_id)
        for gval in np.unique(gid):
            members = np.nonzero(gid == gval)[0]
            any_t = np.zeros(max((len(ovs[i]) for i in members), default=0), dtype=bool)
            for i in members:
                any_t[:len(ovs[i])] |= ovs[i]
            for i in members:
                counts[i] = int(any_t[:len(ovs[i])].sum())
    return np.concatenate(([0], np.cumsum(2 * counts))).astype(np.int32)


def host_batch_from_samples(samples, skip_mode="batch", group_size=None, sgdet=False, pinned=True, with_maps=True):
    """Host lists (reference dataloader tuple shape, dataloader.py:159-165) -> HostBatch (CSR packing, no device work)."""
    n_img = len(samples)
    boxes_l = [(s.bbox_pred if sgdet else s.bbox) for s in samples]
    counts = np.array([b.shape[0] for b in boxes_l], dtype=np.int64)
    tri = counts * (counts - 1) // 2
    arrays = {}
    arrays["box_offsets"] = np.concatenate(([0], np.cumsum(counts))).astype(np.int32)
    arrays["tri_offsets"] = np.concatenate(([0], np.cumsum(tri))).astype(np.int32)
    arrays["boxes"] = np.concatenate([b.numpy() for b in boxes_l]).astype(np.int32)   # float -> int32 truncates toward zero == int()
    cats_l = [(s.categories_pred if sgdet else s.categories) for s in samples]
    arrays["cats"] = np.concatenate([c.numpy() for c in cats_l]).astype(np.int32)
    sup_l = [(s.super_categories_pred if sgdet else s.super_categories) for s in samples]
    supers = -np.ones((int(counts.sum()), 4), dtype=np.int8)
    r = 0
    for lst in sup_l:
        for sc in lst:
            v = np.asarray(sc, dtype=np.int64)
            v = v if len(v) <= 4 else v[:1]      # raw list; the kernels sum first + last entry like utils.py:136-149
            supers[r, :len(v)] = v
            r += 1
    arrays["supers"] = supers
    arrays["box_img"] = np.repeat(np.arange(n_img, dtype=np.int32), counts)
    if not sgdet:
        arrays["rel_tri"] = np.concatenate([np.concatenate([r_.numpy() for r_ in s.relationships]) if len(s.relationships)
                                            else np.zeros(0, np.int64) for s in samples]).astype(np.int32)
        arrays["dir_tri"] = np.concatenate([np.concatenate([r_.numpy() for r_ in s.subj_or_obj]) if len(s.subj_or_obj)
                                            else np.zeros(0, np.float32) for s in samples]).astype(np.int8)
    n_groups = 0
    if skip_mode == "batch":
        gs = group_size or n_img
        gid = (np.arange(n_img) // gs).astype(np.int32)
        n_groups = int(gid.max()) + 1
        arrays["group_id"] = gid
    elif skip_mode != "per_image":
        raise ValueError("skip_mode must be 'batch' or 'per_image'")
    if with_maps:
        arrays["feat"] = torch.stack([s.feat for s in samples])
        arrays["depth"] = torch.stack([s.depth for s in samples])
    if sgdet:
        arrays["conf"] = np.concatenate([s.cat_conf_pred.numpy() for s in samples]).astype(np.float32)
        from .targets import flat_targets_sgd
        for k, v in flat_targets_sgd(samples).items():
            arrays["gt_" + k] = v
    meta = dict(n_groups=n_groups, max_tri=int(tri.max()) if len(tri) else 0, p_max=int((counts * (counts - 1)).sum()),
                cover_fraction=footprint_cover_fraction(arrays["boxes"], arrays["box_offsets"]),
                pair_offsets=host_pair_offsets(arrays["boxes"], arrays["box_offsets"], arrays.get("group_id")))
    return HostBatch(arrays, meta, pinned)


def batch_from_samples(samples, device, skip_mode="batch", group_size=None, sgdet=False, pinned=True, with_maps=True):
    return host_batch_from_samples(samples, skip_mode, group_size, sgdet, pinned, with_maps).to_device(device)


class RelationPipeline:
    """pairs -> head -> candidates -> top-K -> counters for whole batches (the batched entry point of SURVEY §8b)."""

    def __init__(self, packed: Optional[PackedHead], device, commonsense=True, aligned_keys=None, violated_keys=None,
                 top_k=tables.TOP_K, iou_thresh=0.5, feature_size=32, chunk_pairs=16384, predcls=True, conv3_m_sub=2,
                 hier=None, splits=None, overlap=True, conv2_m_sub=1, chunk_policy="waves", conv3_block_rows=4, conv3_shared=True,
                 fc1_shared=True, conv3_block_cols=4, dense_above=0.85):
        self.packed = packed
        # A window whose boxes are so large that the shared-footprint work lists would visit more than this share of the dense
        # conv3_1 pixels (`DeviceBatch.cover_fraction`, a host estimate made when the batch is built) takes the DENSE kernels: there
        # is nothing to skip and the per-box maps / bookkeeping only cost (measured crossover 0.88: full-grid boxes 251 ms shared
        # vs 226 ms dense, profiles/bench_r02g_*).  Same scores either way (bit-identical conv3_1, fc1 to fp32 rounding order).
        self.dense_above = float(dense_above)
        self.last_path = None               # "shared" / "blocks" / "dense": the formulation the last forward_pairs took
        self.host_offsets = os.environ.get("HC_HOST_OFFSETS", "1") != "0"    # use DeviceBatch.pair_offsets_host (no D2H read per step)
        # U / V without the background pre-fill (5.4 GB of writes per cfg2 window): the pooling kernels read the background maps outside
        # a box's conv2_1 footprint rectangle themselves - the same bits (shared-footprint path, box-footprint conv2 only)
        self.uv_select = os.environ.get("HC_UV_SELECT", "1") != "0"
        # first window's sort / masks / zero fill on the pooling stream: measured 42.78 vs 42.92 ms (r02t) - not worth a second allocator
        # pool holding the 13 GB operand; off by default
        self.early_prep = os.environ.get("HC_EARLY_PREP", "0") != "0"
        # per-box fc1 rows as a K-cell-sparse GEMM over each box's own cells (needs the CTA-pair conv3_1 kernel); 0 = dense rows
        self.fc1_box_sparse = os.environ.get("HC_FC1_BOX_SPARSE", "1") != "0"
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("hiercom_b200: RelationPipeline needs a CUDA device (no CPU fallback)")
        self.top_k = tuple(int(k) for k in top_k)
        self.iou_thresh = float(iou_thresh)
        self.fs = feature_size
        self.chunk_pairs = int(chunk_pairs)
        self.chunk_policy = chunk_policy      # "waves": sized for fc1's wave quantisation; "greedy": fill every chunk to the cap
        self.predcls = predcls
        self.conv3_m_sub = conv3_m_sub
        self.conv2_m_sub = conv2_m_sub      # short K (1152): 128-row tiles keep two TMEM stages, so the bf16 epilogue overlaps the next tile
        self.overlap = overlap
        # 0: dense conv3_1.  8 / 4: block-sparse conv3_1 - only 8 x {8,4}-pixel blocks that meet the dilated footprint of the
        # pair's two boxes are computed, the rest of the pooled output is a broadcast of the weights-only background
        if conv3_block_rows not in (0, 2, 4, 8):
            raise ValueError("conv3_block_rows must be 0 (dense), 8, 4 or 2")
        self.conv3_block_rows = int(conv3_block_rows)
        # block width in conv3 pixels: 8, or 4 (with 4 rows: 2 x 2-cell blocks hug the cell rectangles more tightly; with 2 rows: blocks
        # one pooled cell tall and two wide - 11.6 % of the conv3_1 pixels at cfg2 where 2 x 2-cell blocks visit 15.2 %; CTA-pair kernel only)
        self.conv3_block_cols = 4 if (int(conv3_block_cols) == 4 and self.conv3_block_rows in (2, 4)) else 8
        if self.conv3_block_rows == 2 and self.conv3_block_cols != 4:
            raise ValueError("2-row blocks are 4 pixels wide (conv3_block_cols=4)")
        # block-sparse only.  True: a cell of the pooled conv3_1 output that only ONE box of the pair reaches is taken from that
        # box's own map ((box, empty) / (empty, box), computed once per box of the window), so a pair computes only the cells
        # BOTH boxes reach.  False: every cell either box reaches is computed per pair.  Same bits either way.
        self.conv3_shared = bool(conv3_shared) and self.conv3_block_rows > 0
        # shared-footprint fc1 (needs conv3_shared): fc1 is linear, so fc1(pair) = fc1(subject map) + fc1(object map) - fc1(background)
        # + fc1(d), d = the pair's conv3_1 output minus those maps, non-zero only in the cells BOTH boxes reach.  Rows are sorted by
        # that cell rectangle and the GEMM visits, per 256-row tile, only the K cells some row of the tile uses (exact: the skipped
        # operand is zero).  Same sums up to fp32 rounding order and one bf16 rounding of d - not bit-identical to the dense fc1.
        self.fc1_shared = bool(fc1_shared) and self.conv3_shared
        self.fc1_window_pairs = 262144       # pairs per shared-fc1 window: its operand is 128 KB per pair (32 GB at the cap)
        # shared-fc1 path: the tiled pooling writes a pooled conv2 pixel of a pair only if one of the pair's listed conv3_1 blocks
        # reads it (block + 1-pixel halo, `ops.pair_cover_masks`) - the rest of the buffer is never read
        self.pool_footprint = os.environ.get("HC_POOL_FOOTPRINT", "1") != "0"
        self.debug_poison = False
        # conv2_1 halves only within one pixel of each box, background elsewhere: bit-identical to the dense halves
        # (tests/test_gpu_sparse.py), cfg2 step 55.5 -> 51.7 ms on one box (profiles/bench_r01N_*)
        self.conv2_sparse = os.environ.get("HC_CONV2_SPARSE", "1") != "0"
        self.early_pool = os.environ.get("HC_EARLY_POOL", "1") != "0"     # first chunks' pooling starts under the per-box stages
        # conv3_1 on tcgen05 cta_group::2 CTA pairs (4x4-pixel blocks only): the pair shares the tile's pixel operand, so each SM
        # stages half of the small TMA boxes; bit-identical to the single-CTA kernel (tests/test_gpu_sparse.py)
        self.conv3_pairs = int(os.environ.get("HC_CONV3_PAIRS", "1")) if (self.conv3_block_rows in (2, 4) and self.conv3_block_cols == 4) else 0
        if self.conv3_block_rows == 2 and not (self.conv3_pairs and bool(conv3_shared) and bool(fc1_shared)):
            raise ValueError("4x2-pixel blocks need the CTA-pair kernel on the shared-footprint path (conv3_shared, fc1_shared, HC_CONV3_PAIRS=1)")
        self.last_n_blocks = None            # int32 [n_chunks] device tensor: work-list lengths of the last forward_pairs
        self.last_k_masks = None             # int64 [n_tiles] device tensor: K-cell masks of the last shared fc1
        self.last_box_k_masks = None         # the same for the per-box fc1 rows (None: dense rows)
        self.splits = tuple(splits) if splits is not None else (packed.splits if packed is not None and not packed.flat else (15, 11, 24))
        self.hier = (not packed.flat if packed is not None else True) if hier is None else bool(hier)
        self.pass_bitmap = None
        if commonsense:
            bm = ops.cs_bitmap_build(tables.commonsense_aligned_keys() if aligned_keys is None else aligned_keys,
                                     tables.commonsense_violated_keys() if violated_keys is None else violated_keys)
            self.pass_bitmap = torch.from_numpy(bm.view(np.int32)).to(self.device)
        self.zs_bitmap = torch.from_numpy(tables.keys_to_bitmap(tables.zero_shot_keys()).view(np.int32)).to(self.device)
        self.synonyms = None if predcls else torch.from_numpy(tables.object_synonym_matrix()).to(self.device)
        self.n_sm = torch.cuda.get_device_properties(self.device).multi_processor_count
        self.counters = torch.zeros(tables.COUNTER_SIZE, dtype=torch.int64, device=self.device)
        self.stats = torch.zeros(5, dtype=torch.int64, device=self.device)

    # ------------------------------------------------------------------------------------------------ stages
    def enumerate_pairs(self, b: DeviceBatch):
        """R1/R2/R4.  With `b.pair_offsets_host` (counted when the batch was built) nothing is read back from the device: the step is
        enqueued without a host sync, so the host runs ahead of the GPU instead of idling it at every window boundary."""
        known = b.pair_offsets_host if self.host_offsets else None
        return ops.pairs_enumerate(b.boxes, b.box_offsets, b.tri_offsets, b.p_max, b.rel_tri, b.dir_tri, b.group_id, b.n_groups,
                                   b.max_tri, self.fs, offsets_host=known)

    def box_features(self, b: DeviceBatch, boxes=None, box_img=None, prefill=True):
        """Per-image conv1 (+tanh), per-box mask select, per-box conv2 halves -> U, V [nbox,32,32,512] (16-bit).
        prefill=False (box-footprint conv2 only): U / V are defined only inside each box's footprint rectangle; the pooling ops then
        take `self.packed.uv_footprint(boxes)` and read the background maps elsewhere."""
        pk, fs = self.packed, self.fs
        n_img = b.n_images
        boxes = b.boxes if boxes is None else boxes
        box_img = b.box_img if box_img is None else box_img
        x = ops.pack_pixels(b.feat, b.depth, K1_PAD, dtype=pk.act_dtype)
        t = torch.empty(n_img * fs * fs, 256, dtype=self.packed.act_dtype, device=self.device)
        ops.tc_gemm(x, pk.w1, t, n_img * fs * fs, 256, K1_PAD, bias=pk.b1, lda=K1_PAD, ldc=256, epilogue=EPI_BF16, act=ACT_TANH,
                    group_m=8, tag="conv1")
        abox = ops.box_select(t, boxes, box_img, pk.fill, fs)
        if self.conv2_sparse:
            return pk.conv2_halves_sparse(abox, boxes, m_sub=self.conv2_m_sub, prefill=prefill)
        return pk.conv2_halves(abox, m_sub=self.conv2_m_sub)

    def box_maps(self, boxes_x, u, v, with_background_row=False, fp=None):
        """Pooled conv3_1 output of every box of the window paired with the EMPTY box (the last row of boxes_x / u / v):
        -> sub_maps = (box, empty), obj_maps = (empty, box), each [n_box,8,8,1024] bf16, and the work-list lengths.  A real
        pair's output equals sub_maps[s] in the cells only its subject's box reaches and obj_maps[o] in those only its object's
        box reaches (bit for bit: same kernels, same operands inside the receptive field)."""
        pk, br, bc = self.packed, self.conv3_block_rows, self.conv3_block_cols
        n_box = boxes_x.shape[0] - 1
        idx = torch.arange(n_box, dtype=torch.int32, device=self.device)
        empty = torch.full((n_box,), n_box, dtype=torch.int32, device=self.device)
        sub, obj = torch.cat((idx, empty)), torch.cat((empty, idx))
        maps = torch.empty(2 * n_box + (1 if with_background_row else 0), 8, 8, 1024, dtype=self.packed.act_dtype, device=self.device)
        if with_background_row:             # row 2*n_box = the background itself (its fc1 row is the "- fc1(background)" term)
            maps[2 * n_box:].copy_(pk.p3_background())
        starts = list(range(0, 2 * n_box, self.chunk_pairs))
        nblk = torch.zeros(max(len(starts), 1), dtype=torch.int32, device=self.device)
        for k, s in enumerate(starts):
            e = min(2 * n_box, s + self.chunk_pairs)
            p2 = ops.pair_relu_pool(u, v, None, sub[s:e], obj[s:e], self.fs, fp=fp)
            blocks, _ = ops.conv3_active_blocks(boxes_x, sub[s:e], obj[s:e], br, self.fs, n_blocks=nblk[k:k + 1], block_cols=bc)
            ops.broadcast_rows(pk.p3_background(), e - s, maps[s:e])
            pk.conv3_blocks(p2, maps[s:e], e - s, blocks, nblk[k:k + 1], br, m_sub=self.conv3_m_sub, tag="conv3_box", block_cols=bc,
                            cta_pairs=self.conv3_pairs)
            del p2
        if with_background_row:
            return maps, nblk
        return maps[:n_box], maps[n_box:], nblk

    def box_maps_sparse(self, boxes_x, u, v, fp=None):
        """`box_maps` + the per-box fc1 rows in one pass, with the fc1 rows K-cell-sparse: a box's map differs from the background only
        in the cells the box itself reaches, so the 2*n_box maps are produced in SORTED order (by the box's cell rectangle) and the
        fc1 GEMM visits, per 256-row tile, only the union of its rows' cells; the background's share of the skipped cells is a
        per-tile constant (`PackedHead.fc1_rows_sparse`).  No extra rounding: the operand is the maps themselves.
        -> maps [2*n_box + 1, 8,8,1024] in sorted order (last row = background), map_row int32 [2*n_box] (row of (box i, empty) at i,
        of (empty, box i) at n_box + i), f_rows f32 [2*n_box, 4096] = fc1(map) in BOX order, work-list lengths."""
        pk, br, bc, fs, dev = self.packed, self.conv3_block_rows, self.conv3_block_cols, self.fs, self.device
        n_box = boxes_x.shape[0] - 1
        n = 2 * n_box
        idx = torch.arange(n_box, dtype=torch.int32, device=dev)
        empty = torch.full((n_box,), n_box, dtype=torch.int32, device=dev)
        own = torch.cat((idx, idx))                                   # the box whose own cells a row can differ from the background in
        keys = ops.pair_cell_keys(boxes_x, own, own, fs)             # box & box = the box's own cell rectangle
        perm64 = torch.sort(keys, stable=True)[1]                    # sorted row -> (role, box)
        perm = perm64.to(torch.int32)
        map_row = torch.empty_like(perm)
        map_row[perm64] = torch.arange(n, dtype=torch.int32, device=dev)
        sub = torch.cat((idx, empty))[perm64].contiguous()            # the (box, empty) / (empty, box) pairs, in sorted order
        obj = torch.cat((empty, idx))[perm64].contiguous()
        own_sorted = own[perm64].contiguous()
        masks = ops.tile_cell_masks(boxes_x, own_sorted, own_sorted, 256, fs)
        bg = pk.p3_background()
        maps = torch.empty(n + 1, 8, 8, 1024, dtype=pk.act_dtype, device=dev)
        maps[n:].copy_(bg)
        starts = list(range(0, n, self.chunk_pairs))
        nblk = torch.zeros(max(len(starts), 1), dtype=torch.int32, device=dev)
        for k, s in enumerate(starts):
            e = min(n, s + self.chunk_pairs)
            cover = ops.pair_cover_masks(boxes_x, sub[s:e], obj[s:e], br, bc, False, fs) if self.pool_footprint else None
            p2 = ops.pair_relu_pool(u, v, None, sub[s:e], obj[s:e], fs, cover=cover, fp=fp)   # only the pixels the listed blocks read
            blocks, _ = ops.conv3_active_blocks(boxes_x, sub[s:e], obj[s:e], br, fs, n_blocks=nblk[k:k + 1], block_cols=bc)
            ops.broadcast_rows(bg, e - s, maps[s:e])
            pk.conv3_blocks(p2, maps[s:e], e - s, blocks, nblk[k:k + 1], br, m_sub=self.conv3_m_sub, tag="conv3_box", block_cols=bc,
                            cta_pairs=self.conv3_pairs)
            del p2
        f_rows = pk.fc1_rows_sparse(maps[:n], n, masks, perm)
        self.last_box_k_masks = masks
        return maps, map_row, f_rows, nblk

    @staticmethod
    def _greedy_chunks(offsets_host, cap):
        chunks, i, n_img = [], 0, len(offsets_host) - 1
        while i < n_img:
            j = i + 1
            while j < n_img and offsets_host[j + 1] - offsets_host[i] <= cap:
                j += 1
            if offsets_host[j] > offsets_host[i]:
                chunks.append((i, j - i, int(offsets_host[i]), int(offsets_host[j] - offsets_host[i])))
            i = j
        return chunks

    @staticmethod
    def _chunk_cost(chunks, n_sm=148):
        """Estimated milliseconds a chunking adds beyond the tile work itself (see `_image_chunks`)."""
        if not chunks:
            return 0.0
        rounds = sum(-(-(-(-c[3] // 256) * 16) // n_sm) for c in chunks)
        return 0.78 * rounds + 77e-6 * min(c[3] for c in chunks) + 0.03 * len(chunks)

    def _image_chunks(self, offsets_host, n_sm=None):
        """Image-aligned chunks of at most `chunk_pairs` directed pairs: (img0, n_img, pair_base, n_pairs).
        The chunk size is picked for the fc1 GEMM's wave quantisation: a chunk of n pairs is ceil(n/256) x 16 tiles of
        256 x 256 on `n_sm` persistent CTAs, i.e. ceil(tiles / n_sm) rounds, and a badly sized chunk idles most of the last
        round (cfg2 at 10 images per chunk: 976 tiles = 6.6 rounds -> 7, 94 % busy; at 12 images: 1 184 tiles = 8.0 rounds).
        Every greedy chunking whose capacity is one of the reachable prefix sizes is scored in milliseconds: fc1 rounds
        (0.78 ms each, measured), plus the pooling of the FIRST chunk (77 ns per pair - the only pooling not hidden under a
        previous chunk's GEMMs; the shortest chunk is moved to the front), plus 30 us of launch / pipeline fill per chunk."""
        n_img = len(offsets_host) - 1
        if getattr(self, "chunk_policy", "waves") == "greedy":
            return self._greedy_chunks(offsets_host, self.chunk_pairs)
        if n_sm is None:
            n_sm = getattr(self, "n_sm", None) or 148
        caps = {int(self.chunk_pairs)}
        for j in range(1, n_img + 1):
            c = int(offsets_host[j] - offsets_host[0])
            if 0 < c <= self.chunk_pairs:
                caps.add(c)
        best = None
        for cap in sorted(caps):
            ch = self._greedy_chunks(offsets_host, cap)
            key = self._chunk_cost(ch, n_sm)
            if best is None or key < best[0]:
                best = (key, ch)
        chunks = list(best[1]) if best else []
        if len(chunks) > 1:
            k = min(range(len(chunks)), key=lambda t: chunks[t][3])
            chunks.insert(0, chunks.pop(k))
        return chunks

    def forward_pairs_tensors(self, feat, depth, boxes, cats, supercats, pair_index, box_img=None):
        """The batched entry point with the argument list of SURVEY §8(b): plain tensors instead of a DeviceBatch.
          feat f32 [B,256,32,32], depth f32 [B,1,32,32]      DETR features / depth of the window's images (evaluate.py:103-109)
          boxes [nbox,4] (xmin,xmax,ymin,ymax) on the 32-grid  (any numeric dtype; truncated toward zero like the reference's int())
          cats int [nbox], supercats int8 [nbox,4] (-1 padded raw lists) or a list of id lists (utils.py:136-149 input format)
          pair_index int [P,2] = (subject box row, object box row); box_img int [nbox] image of each box (default: one image)
        -> (relation [P,R], super [P,3], connectivity [P], log sigmoid(connectivity) [P]) exactly as `forward_pairs`."""
        from .model import supers_to_table
        dev = self.device
        boxes = torch.as_tensor(boxes).to(dev).to(torch.int32).contiguous()
        n_box = boxes.shape[0]
        box_img = torch.zeros(n_box, dtype=torch.int32, device=dev) if box_img is None else torch.as_tensor(box_img).to(dev, torch.int32).contiguous()
        n_img = int(feat.shape[0])
        counts = torch.bincount(box_img.long(), minlength=n_img)
        box_offsets = torch.cat((counts.new_zeros(1), counts.cumsum(0))).to(torch.int32)
        if not isinstance(supercats, torch.Tensor):
            supercats = supers_to_table(supercats, dev)
        pair_index = torch.as_tensor(pair_index).to(dev)
        b = DeviceBatch(feat.to(dev, torch.float32).contiguous(), depth.to(dev, torch.float32).contiguous(), boxes, box_offsets, box_img,
                        torch.as_tensor(cats).to(dev, torch.int32).contiguous(), supercats.to(dev, torch.int8).contiguous(),
                        torch.zeros(n_img + 1, dtype=torch.int32, device=dev), None, None, None, 0, 0, 0, 0)
        pairs = dict(n=int(pair_index.shape[0]), sub=pair_index[:, 0].to(torch.int32).contiguous(), obj=pair_index[:, 1].to(torch.int32).contiguous())
        if pairs["n"] == 0:
            r = sum(self.splits)
            z = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)
            return z(0, r), z(0, 3), z(0), z(0)
        return self.forward_pairs(b, pairs)

    def forward_pairs(self, b, pairs=None, *tensors, **kw):
        """R3,R5-R7 for every directed pair of the batch -> (relation [P,R], super [P,3], connectivity [P], logsig [P]).
        Called with tensors, `forward_pairs(feat, depth, boxes, cats, supercats, pair_index[, box_img])`, it is the SURVEY §8(b)
        signature (see `forward_pairs_tensors`).
        Pair lists from `enumerate_pairs` take the tiled outer-sum pooling kernel on image-aligned chunks, with the
        pooling of chunk k+1 overlapped (second stream) with the tensor-core GEMMs of chunk k; arbitrary pair lists
        (no `offsets_host`) take the generic gather kernel."""
        if not isinstance(b, DeviceBatch):
            return self.forward_pairs_tensors(b, pairs, *tensors, **kw)
        pk = self.packed
        n = pairs["n"]
        br, bc, shared = self.conv3_block_rows, self.conv3_block_cols, self.conv3_shared
        if br and b.cover_fraction is not None and b.cover_fraction > self.dense_above:
            br, bc, shared = 0, 8, False                         # nothing to skip in this window: the dense kernels
            self.last_n_blocks = self.last_k_masks = None
        self.last_path = "dense" if not br else ("shared" if shared else "blocks")
        if self.fc1_shared and shared and n > 0:
            return self._forward_pairs_fc1_shared(b, pairs)
        if shared:      # one more box per window: the empty one (all background), partner of every box in `box_maps`
            boxes_x = torch.cat((b.boxes, b.boxes.new_zeros(1, 4)))
            u, v = self.box_features(b, boxes_x, torch.cat((b.box_img, b.box_img.new_zeros(1))))
            p3_bg = pk.p3_background()
            sub_maps, obj_maps, nblk_box = self.box_maps(boxes_x, u, v)
            list_blocks = ops.conv3_shared_blocks
        else:
            u, v = self.box_features(b)
            nblk_box = None
            list_blocks = ops.conv3_active_blocks

        def prefill(p3, sub, obj, cnt):
            """everything of the pooled conv3_1 output that the work list will not write"""
            if shared:
                ops.p3_assemble(p3_bg, sub_maps, obj_maps, b.boxes, sub, obj, p3)
            else:
                ops.broadcast_rows(p3_bg, cnt, p3)

        raw = torch.empty(n, 512, dtype=torch.float32, device=self.device)
        if "offsets_host" not in pairs:
            starts = list(range(0, n, self.chunk_pairs))
            nblk = torch.zeros(max(len(starts), 1), dtype=torch.int32, device=self.device) if br else None
            for k, s in enumerate(starts):
                e = min(n, s + self.chunk_pairs)
                p2 = ops.pair_relu_pool(u, v, None, pairs["sub"][s:e], pairs["obj"][s:e], self.fs)
                if br:
                    blocks, _ = list_blocks(b.boxes, pairs["sub"][s:e], pairs["obj"][s:e], br, self.fs, n_blocks=nblk[k:k + 1], block_cols=bc)
                    p3 = None
                    if shared:
                        p3 = torch.empty(e - s, 8, 8, 1024, dtype=self.packed.act_dtype, device=self.device)
                        prefill(p3, pairs["sub"][s:e], pairs["obj"][s:e], e - s)
                    pk.conv3_fc(p2, m_sub=self.conv3_m_sub, raw=raw[s:e], blocks=blocks, n_blocks=nblk[k:k + 1], block_rows=br, p3=p3,
                                block_cols=bc)
                    del p3
                else:
                    pk.conv3_fc(p2, m_sub=self.conv3_m_sub, raw=raw[s:e])
                del p2
            self.last_n_blocks = nblk if nblk_box is None else torch.cat((nblk, nblk_box))
        else:
            n_box = b.boxes.shape[0]
            n_max = int(np.max(np.diff(b.box_offsets_host))) if b.box_offsets_host is not None else int(
                (b.box_offsets[1:] - b.box_offsets[:-1]).max().item())
            lut = ops.pair_lut_build(pairs["sub"], pairs["obj"], pairs["img"], b.box_offsets, n_box, n_max)
            chunks = self._image_chunks(pairs["offsets_host"])
            cap = max(c[3] for c in chunks)
            bufs = [torch.empty(cap, self.fs // 2, self.fs // 2, 512, dtype=self.packed.act_dtype, device=self.device)
                    for _ in range(2 if self.overlap and len(chunks) > 1 else 1)]
            if br:      # per buffer: the work list and the background-filled pooled conv3_1 output
                blk_bufs = [torch.empty(cap * (256 // (br * bc)), dtype=torch.int32, device=self.device) for _ in bufs]
                p3_bufs = [torch.empty(cap, 8, 8, 1024, dtype=self.packed.act_dtype, device=self.device) for _ in bufs]
                nblk = torch.zeros(len(chunks), dtype=torch.int32, device=self.device)
                p3_bg = pk.p3_background()
            main = torch.cuda.current_stream()
            side = self._side_stream() if len(bufs) == 2 else main
            ready = torch.cuda.Event()
            ready.record(main)
            gemm_done = []
            for k, (img0, n_img, base, cnt) in enumerate(chunks):
                buf = bufs[k % len(bufs)]
                with torch.cuda.stream(side):
                    if side is not main:
                        side.wait_event(ready)
                        if k >= 2:
                            side.wait_event(gemm_done[k - 2])          # buffer free again
                    ops.pair_relu_pool_tiled(u, v, None, b.box_offsets, lut, img0, n_img, base, cnt, self.fs, out=buf)
                    if br:
                        list_blocks(b.boxes, pairs["sub"][base:base + cnt], pairs["obj"][base:base + cnt], br, self.fs,
                                    blocks=blk_bufs[k % len(bufs)], n_blocks=nblk[k:k + 1], block_cols=bc)
                        prefill(p3_bufs[k % len(bufs)], pairs["sub"][base:base + cnt], pairs["obj"][base:base + cnt], cnt)
                    pooled = torch.cuda.Event()
                    pooled.record(side)
                if side is not main:
                    main.wait_event(pooled)
                if br:
                    pk.conv3_fc(buf, m_sub=self.conv3_m_sub, raw=raw[base:base + cnt], n=cnt, blocks=blk_bufs[k % len(bufs)],
                                n_blocks=nblk[k:k + 1], block_rows=br, p3=p3_bufs[k % len(bufs)], block_cols=bc)
                else:
                    pk.conv3_fc(buf, m_sub=self.conv3_m_sub, raw=raw[base:base + cnt], n=cnt)
                ev = torch.cuda.Event()
                ev.record(main)
                gemm_done.append(ev)
            if side is not main:
                main.wait_stream(side)
            if br:
                self.last_n_blocks = nblk if nblk_box is None else torch.cat((nblk, nblk_box))
        relation, sup, conn, logsig, _ = pk.heads(raw, pairs["sub"], pairs["obj"], b.cats, b.supers)
        return relation, sup, conn, logsig

    def _forward_pairs_fc1_shared(self, b: DeviceBatch, pairs):
        """`forward_pairs` with the shared-footprint fc1: per-box conv3_1 maps and their fc1 rows once per box; per pair only the
        conv3_1 blocks covering the cells both boxes reach, written as differences into the sorted operand d; ONE K-cell-sparse fc1
        + fc2 over all pairs of a window; raw comes back in pair order through the fc2 epilogue's row map.  The operand is 128 KB
        per pair (only the visited cells are touched), so batches beyond `fc1_window_pairs` are cut into image-aligned windows."""
        pk, fs = self.packed, self.fs
        n, n_box = pairs["n"], b.boxes.shape[0]
        dev = self.device
        lut = None
        if "offsets_host" in pairs:
            n_max = int(np.max(np.diff(b.box_offsets_host))) if b.box_offsets_host is not None else int(
                (b.box_offsets[1:] - b.box_offsets[:-1]).max().item())
            lut = ops.pair_lut_build(pairs["sub"], pairs["obj"], pairs["img"], b.box_offsets, n_box, n_max)
        windows = self._fc1_windows(pairs)
        # The first window's row order, K-cell masks and zero-filled operand need only the boxes and the pair list: they are made on
        # the pooling stream NOW, under the per-image / per-box stages (HC_EARLY_PREP; everything the allocator may recycle into these
        # buffers was last used by work already queued on the compute stream, hence the event)
        prep0 = None
        if self.early_prep and self.overlap:
            main, side = torch.cuda.current_stream(), self._side_stream()
            queued = torch.cuda.Event()
            queued.record(main)
            with torch.cuda.stream(side):
                side.wait_event(queued)
                prep0 = self._window_prep(b, pairs, windows[0][0], windows[0][1])
                prep0["done"] = torch.cuda.Event()
                prep0["done"].record(side)
            for t in prep0.values():
                if isinstance(t, torch.Tensor):
                    t.record_stream(main)                            # allocated on the pooling stream, consumed on the compute stream
        boxes_x = torch.cat((b.boxes, b.boxes.new_zeros(1, 4)))      # + the empty box (all background), partner of every box
        select = self.uv_select and self.conv2_sparse                 # no background pre-fill of U / V: the pooling kernels select
        u, v = self.box_features(b, boxes_x, torch.cat((b.box_img, b.box_img.new_zeros(1))), prefill=not select)
        fp = pk.uv_footprint(boxes_x) if select else None
        # The pooling of the first chunks needs only U and V: its buffers are taken NOW (anything the allocator recycles into them
        # was last used by work already queued on this stream) and an event lets the pooling stream start under the per-box stages
        early = self._pool_buffers(windows[0][2])
        uv_ready = torch.cuda.Event()
        uv_ready.record(torch.cuda.current_stream())
        if self.fc1_box_sparse:
            # maps in sorted order + their fc1 rows from a K-cell-sparse GEMM over each box's own cells
            maps, map_row, f_box, nblk_box = self.box_maps_sparse(boxes_x, u, v, fp=fp)
            bias_eff = (pk.b_fc1 - pk.fc1_background()).contiguous()
            sub_rows = map_row[pairs["sub"].long()].contiguous()                     # row of (subject, empty) / (empty, object) in `maps`
            obj_rows = map_row[(pairs["obj"] + n_box).long()].contiguous()
        else:
            maps, nblk_box = self.box_maps(boxes_x, u, v, with_background_row=True, fp=fp)
            f_box = pk.fc1_rows(maps, 2 * n_box + 1)                 # fc1 (no bias) of (box, empty), (empty, box), background
            bias_eff = (pk.b_fc1 - f_box[2 * n_box]).contiguous()
            sub_rows, obj_rows = pairs["sub"], pairs["obj"] + n_box
        raw = torch.empty(n, 512, dtype=torch.float32, device=dev)
        nblks, masks_all = [], []
        for i, (w0, w1, chunks) in enumerate(windows):
            nblk, masks = self._fc1_shared_window(b, pairs, w0, w1, chunks, u, v, lut, maps, f_box, bias_eff, raw,
                                                  early if i == 0 else None, uv_ready if i == 0 and self.early_pool else None, fp=fp,
                                                  prep=prep0 if i == 0 else None, map_rows=(sub_rows, obj_rows))
            nblks.append(nblk)
            masks_all.append(masks)
        self.last_n_blocks = torch.cat(nblks + [nblk_box])
        self.last_k_masks = torch.cat(masks_all)
        relation, sup, conn, logsig, _ = pk.heads(raw, pairs["sub"], pairs["obj"], b.cats, b.supers)
        return relation, sup, conn, logsig

    def _fc1_windows(self, pairs):
        """Host only: cut the pair list into shared-fc1 windows of at most `fc1_window_pairs` pairs (image-aligned when the list
        came from `enumerate_pairs`) and each window into pooling / conv3_1 chunks: [(w0, w1, [(img0, n_img, pair_base, n_pairs)])]."""
        n = pairs["n"]
        windows = []
        if "offsets_host" in pairs:
            off = pairs["offsets_host"]
            for img0, n_img, base, cnt in self._greedy_chunks(off, self.fc1_window_pairs):
                chunks = [(img0 + c[0], c[1], c[2], c[3]) for c in self._image_chunks(off[img0:img0 + n_img + 1])]
                windows.append((base, base + cnt, chunks))
        else:
            for w0 in range(0, n, self.fc1_window_pairs):
                w1 = min(n, w0 + self.fc1_window_pairs)
                windows.append((w0, w1, [(0, 0, s, min(w1, s + self.chunk_pairs) - s) for s in range(w0, w1, self.chunk_pairs)]))
        return windows

    def _pool_buffers(self, chunks):
        """Everything the pooling stream writes for a window: pooled conv2 buffers (double-buffered), work lists, cover words, counts."""
        fs, br, bc, dev = self.fs, self.conv3_block_rows, self.conv3_block_cols, self.device
        cap = max(c[3] for c in chunks)
        n_buf = 2 if self.overlap and len(chunks) > 1 else 1
        return dict(bufs=[torch.empty(cap, fs // 2, fs // 2, 512, dtype=self.packed.act_dtype, device=dev) for _ in range(n_buf)],
                    blk=[torch.empty(cap * (256 // (br * bc)), dtype=torch.int32, device=dev) for _ in range(n_buf)],
                    cov=[torch.empty(cap, dtype=torch.int64, device=dev) for _ in range(n_buf)],
                    nblk=torch.zeros(len(chunks), dtype=torch.int32, device=dev))

    def _window_prep(self, b, pairs, w0, w1):
        """Row order of the fc1 operand of pairs [w0, w1): sorted by the cell rectangle both boxes reach (pairs with none last), the
        per-tile K-cell masks, and the operand itself with the visited cells zero-filled.  Needs only the boxes and the pair list."""
        fs, dev, n = self.fs, self.device, w1 - w0
        sub_w, obj_w = pairs["sub"][w0:w1], pairs["obj"][w0:w1]
        keys = ops.pair_cell_keys(b.boxes, sub_w, obj_w, fs)
        perm64 = torch.sort(keys, stable=True)[1]                    # sorted row -> pair (window-local)
        perm = perm64.to(torch.int32)
        row_of = torch.empty_like(perm)
        row_of[perm64] = torch.arange(n, dtype=torch.int32, device=dev)
        row_sub, row_obj = sub_w[perm64].contiguous(), obj_w[perm64].contiguous()
        masks = ops.tile_cell_masks(b.boxes, row_sub, row_obj, 256, fs)
        d = torch.empty(n, 64, 1024, dtype=self.packed.act_dtype, device=dev)
        ops.cells_zero(masks, 256, n, d)
        return dict(perm=perm, row_of=row_of, row_sub=row_sub, row_obj=row_obj, masks=masks, d=d, done=None)

    def _fc1_shared_window(self, b, pairs, w0, w1, chunks, u, v, lut, maps, f_box, bias_eff, raw, pool=None, pool_ready=None, fp=None,
                           prep=None, map_rows=None):
        """Pairs [w0, w1) of the batch: sort, conv3_1 differences chunk by chunk (pooling of chunk k+1 under the GEMM of chunk k),
        then one K-cell-sparse fc1 + fc2 into raw[w0:w1]."""
        pk, fs, br, bc = self.packed, self.fs, self.conv3_block_rows, self.conv3_block_cols
        n, n_box = w1 - w0, b.boxes.shape[0]
        if map_rows is None:                 # maps = [(box, empty) maps | (empty, box) maps | background], in box order
            map_rows = (pairs["sub"], pairs["obj"] + n_box)
        main = torch.cuda.current_stream()
        if prep is None:
            prep = self._window_prep(b, pairs, w0, w1)
        elif prep.get("done") is not None:
            main.wait_event(prep["done"])                            # prepared on the side stream under the per-box stages
        perm, row_of, row_sub, row_obj, masks, d = (prep[k] for k in ("perm", "row_of", "row_sub", "row_obj", "masks", "d"))
        if pool is None:
            pool = self._pool_buffers(chunks)
        bufs, blk_bufs, cov_bufs, nblk = pool["bufs"], pool["blk"], pool["cov"], pool["nblk"]
        two = len(bufs) == 2
        side = self._side_stream() if two else main
        ready = pool_ready
        if ready is None:
            ready = torch.cuda.Event()
            ready.record(main)
        gemm_done = []
        for k, (img0, n_img, base, cnt) in enumerate(chunks):
            buf, blk = bufs[k % len(bufs)], blk_bufs[k % len(bufs)]
            sub_k, obj_k = pairs["sub"][base:base + cnt], pairs["obj"][base:base + cnt]
            with torch.cuda.stream(side):
                if side is not main:
                    side.wait_event(ready)
                    if k >= 2:
                        side.wait_event(gemm_done[k - 2])              # buffer free again
                if self.debug_poison:            # tests: anything conv3_1 reads that the pooling did not write shows up as NaN
                    buf.fill_(float("nan"))
                if lut is not None:
                    cover = None
                    if self.pool_footprint:      # only the pooled pixels a listed block reads (block + halo)
                        cover = ops.pair_cover_masks(b.boxes, sub_k, obj_k, br, bc, True, fs, out=cov_bufs[k % len(bufs)])
                    ops.pair_relu_pool_tiled(u, v, None, b.box_offsets, lut, img0, n_img, base, cnt, fs, out=buf, cover=cover, fp=fp)
                else:
                    ops.pair_relu_pool(u, v, None, sub_k, obj_k, fs, out=buf, fp=fp)
                ops.conv3_shared_blocks(b.boxes, sub_k, obj_k, br, fs, blocks=blk, n_blocks=nblk[k:k + 1], block_cols=bc)
                pooled = torch.cuda.Event()
                pooled.record(side)
            if side is not main:
                main.wait_event(pooled)
            pk.conv3_diff(buf, d, cnt, blk, nblk[k:k + 1], br, maps, maps, map_rows[0][base:base + cnt], map_rows[1][base:base + cnt],
                          row_of[base - w0:base - w0 + cnt], m_sub=self.conv3_m_sub, block_cols=bc, cta_pairs=self.conv3_pairs)
            ev = torch.cuda.Event()
            ev.record(main)
            gemm_done.append(ev)
        if side is not main:
            main.wait_stream(side)
        pk.fc1_shared_fc2(d, n, masks, f_box[:n_box], f_box[n_box:], row_sub, row_obj, bias_eff, perm, raw[w0:w1])
        return nblk, masks

    def _side_stream(self):
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def evaluate(self, b: DeviceBatch, pairs, relation, sup, conn_logsig, connectivity=None, want_topk=False):
        """R8-R13 on given scores (bit-exact stage: identical scores in -> identical counters out)."""
        n = pairs["n"]
        if n == 0:
            return None
        hier = self.hier
        k = 3 if hier else 1
        want_t3 = hier and self.predcls
        cand_conf, cand_label, t3_conf, t3_super = ops.candidates(
            relation, self.splits, hier, pairs["ov"], conn_logsig, pairs["sub"], pairs["obj"], b.cats, self.pass_bitmap, sup,
            conf_sub=None if b.conf is None else b.conf[pairs["sub"].long()],
            conf_obj=None if b.conf is None else b.conf[pairs["obj"].long()], want_top3=want_t3)
        if self.predcls:
            gt = dict(offsets=pairs["offsets"], label=pairs["gt"], sub=pairs["sub"], obj=pairs["obj"], cat=b.cats, box=b.boxes)
        else:
            gt = b.gt
        cand_offsets = (pairs["offsets"] * k).contiguous()
        ev = self.counters[:tables.EV_SIZE]
        topk = ops.topk_match(cand_offsets, cand_conf, cand_label, k, pairs["sub"], pairs["obj"], b.cats, b.boxes, gt["offsets"],
                              gt["label"], gt["sub"], gt["obj"], gt["cat"], gt["box"], ev, synonyms=self.synonyms,
                              zs_bitmap=self.zs_bitmap, mode=0, feature_size=self.fs, iou_thresh=self.iou_thresh, top_k=self.top_k,
                              want_topk=want_topk)
        topk3 = None
        if want_t3:
            t3 = self.counters[tables.EV_SIZE:]
            topk3 = ops.topk_match(pairs["offsets"], t3_conf, None, 1, pairs["sub"], pairs["obj"], b.cats, b.boxes, gt["offsets"],
                                   gt["label"], gt["sub"], gt["obj"], gt["cat"], gt["box"], t3, mode=1, t3_labels=cand_label,
                                   t3_super=t3_super, feature_size=self.fs, iou_thresh=self.iou_thresh, top_k=self.top_k,
                                   want_topk=want_topk)
        if connectivity is not None and self.predcls:
            ops.connectivity_stats(connectivity, pairs["gt"], pairs["rel"], self.stats)
        return dict(cand_conf=cand_conf, cand_label=cand_label, t3_conf=t3_conf, topk=topk, topk3=topk3)

    def step(self, b: DeviceBatch):
        """One pass of the hot path over one batch; returns the number of directed pairs processed."""
        pairs = self.enumerate_pairs(b)
        if pairs["n"] == 0:
            return 0
        relation, sup, conn, logsig = self.forward_pairs(b, pairs)
        self.evaluate(b, pairs, relation, sup, logsig, connectivity=conn)
        return pairs["n"]

    def run(self, host_batches, before_step=None, after_step=None, pipelined=True):
        """Streams evaluation windows from pinned host staging: yields `(n_pairs, counters_host)` per window, in order.  The H2D copy
        of window k+1 is issued on a copy stream before window k's kernels are enqueued, so it rides under window k's compute;
        every window's inputs cross PCIe exactly once and every window's counters are read back (D2H into pinned memory).
        `pipelined`: the read-back of window k is an asynchronous copy behind its kernels and is handed out AFTER window k+1 has
        been enqueued, so the GPU never waits for the host between windows (with `DeviceBatch.pair_offsets_host` the step itself
        has no host sync either); `pipelined=False` reads window k back before window k+1 is enqueued.
        `before_step(self)` / `after_step(self)` hook in a counter reset / the cross-rank all-reduce (a tensor returned by
        `after_step` - the reduced copy - is what gets read back instead of the rank-local counters)."""
        main = torch.cuda.current_stream(self.device)
        if getattr(self, "_copy", None) is None:
            self._copy = torch.cuda.Stream(device=self.device)
        copy = self._copy
        it = iter(host_batches)

        def fetch():
            hb = next(it, None)
            if hb is None:
                return None
            with torch.cuda.stream(copy):
                b = hb.to_device(self.device)
                ready = torch.cuda.Event()
                ready.record(copy)
            return b, ready

        pending = None                      # (n_pairs, pinned result, event) of the previous window
        nxt = fetch()
        while nxt is not None:
            b, ready = nxt
            main.wait_event(ready)
            for t in list(vars(b).values()) + list((b.gt or {}).values()):
                if isinstance(t, torch.Tensor) and t.is_cuda:
                    t.record_stream(main)           # allocated on the copy stream, consumed on the compute stream
            nxt = fetch()
            if before_step is not None:
                before_step(self)
            n = self.step(b)
            result = self.counters
            if after_step is not None:
                r = after_step(self)
                if isinstance(r, torch.Tensor):        # e.g. dist.allreduce_counters(self.counters): the GLOBAL sums are read back
                    result = r
            if not pipelined:
                yield n, result.cpu()
                continue
            # pinned staging from a small ring owned by the pipeline (no pinned allocation - which can synchronise the device - on
            # the step path); a slot is reused two windows later, after its copy was handed out as an ordinary host tensor
            ring = getattr(self, "_result_ring", None)
            if ring is None or ring[0].shape != result.shape or ring[0].dtype != result.dtype:
                ring = self._result_ring = [torch.empty(result.shape, dtype=result.dtype, pin_memory=True) for _ in range(3)]
                self._ring_pos = 0
            host = ring[self._ring_pos]
            self._ring_pos = (self._ring_pos + 1) % len(ring)
            host.copy_(result, non_blocking=True)      # stream-ordered behind this window's kernels, before the next reset
            done = torch.cuda.Event()
            done.record(main)
            if pending is not None:
                pending[2].synchronize()
                yield pending[0], pending[1].clone()
            pending = (n, host, done)
        if pending is not None:
            pending[2].synchronize()
            yield pending[0], pending[1].clone()

    # ------------------------------------------------------------------------------------------------ results
    def reset(self):
        self.counters.zero_()
        self.stats.zero_()

    def metrics(self, counters=None):
        c = (self.counters if counters is None else counters).cpu().numpy()
        return metrics_from_counters(c, self.top_k)


def _recall_block(c, top_k):
    """evaluator.py:358-365: R@k in Python floats, mR@k = float32 nanmean over predicates."""
    nk = len(top_k)
    hits = c[:nk].astype(np.float64)
    hits_pc = c[nk:nk + nk * tables.NUM_PRED].reshape(nk, tables.NUM_PRED)
    n = float(c[nk + nk * tables.NUM_PRED])
    n_pc = c[nk + nk * tables.NUM_PRED + 1:nk + nk * tables.NUM_PRED + 1 + tables.NUM_PRED]
    recall = [float(hits[i]) / max(n, 1e-3) for i in range(nk)]
    per_class = [torch.as_tensor(hits_pc[i], dtype=torch.float32) / torch.as_tensor(n_pc, dtype=torch.float32) for i in range(nk)]
    mean_recall = [torch.nanmean(r) for r in per_class]
    return recall, per_class, mean_recall


def metrics_from_counters(c, top_k=tables.TOP_K):
    """765-slot int64 counter vector -> the reference's return tuples:
    Evaluator.compute 6-tuple (evaluator.py:367) and Evaluator_Top3.compute 3-tuple (:773)."""
    c = np.asarray(c, dtype=np.int64)
    ev = _recall_block(c[:tables.EV_BLOCK], top_k) + _recall_block(c[tables.EV_BLOCK:tables.EV_SIZE], top_k)
    t3c = c[tables.EV_SIZE:]
    nk = len(top_k)
    t3_block = np.concatenate((t3c[tables.T3_HITS:tables.T3_TOP1], t3c[tables.T3_NGT:tables.T3_SIZE]))
    t3 = _recall_block(t3_block, top_k)
    t3_top1 = np.concatenate((t3c[tables.T3_TOP1:tables.T3_NGT], t3c[tables.T3_NGT:tables.T3_SIZE]))
    return dict(evaluator=ev, top3=t3, top3_top1=_recall_block(t3_top1, top_k))
