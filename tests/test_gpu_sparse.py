"""GPU tests of the block-sparse conv3_1 (HC_GEMM_CONV3_BLOCKS): the work list covers the dilated footprint of each pair's
two boxes (or, shared-footprint mode, only the cells BOTH boxes reach - the rest is assembled from per-box maps), and the sparse
path (pre-fill + listed blocks) reproduces the dense path BIT FOR BIT - the reference
(model.py:138-150 on `feature*mask`, train_test.py:391,398) is dense, so equality with the dense kernels is the parity bar."""
import numpy as np
import pytest
import torch

from scene_graph_commonsense_b200 import synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda"

# (xmin, xmax, ymin, ymax) on the 32-grid: corners, borders, 1-pixel, empty, inverted, whole map, negative (Python slice wrap)
EDGE_BOXES = [(0, 32, 0, 32), (0, 1, 0, 1), (31, 32, 31, 32), (0, 4, 28, 32), (5, 5, 3, 9), (9, 3, 2, 7), (14, 18, 14, 18),
              (0, 32, 15, 17), (15, 17, 0, 32), (-4, 32, 3, 12), (1, 3, 1, 3), (2, 30, 2, 30), (7, 9, 20, 31), (30, 32, 0, 2)]


def _cells_1d(lo, hi):
    """include/hiercom_b200.h hc_conv3_active_blocks: box interval [lo,hi) -> 8-grid cells that can differ from the background."""
    if hi <= lo:
        return 0, 0
    qlo, qhi = max(0, (lo - 1) // 2), min(15, hi // 2)
    qlo, qhi = max(0, qlo - 1), min(15, qhi + 1)
    return qlo // 2, qhi // 2 + 1


def _slice_bound(v, size=32):
    if v < 0:
        v = max(v + size, 0)
    return min(v, size)


def _cell_mask(box):
    x0, x1, y0, y1 = (_slice_bound(int(v)) for v in box)
    x1, y1 = max(x1, x0), max(y1, y0)
    m = np.zeros((8, 8), bool)
    xa, xb = _cells_1d(x0, x1)
    ya, yb = _cells_1d(y0, y1)
    if xb > xa and yb > ya:
        m[ya:yb, xa:xb] = True
    return m


def _random_boxes(n, seed):
    g = torch.Generator().manual_seed(seed)
    b = synthetic.make_boxes(g, n).to(torch.int32)
    k = min(len(EDGE_BOXES), n)
    b[:k] = torch.tensor(EDGE_BOXES[:k], dtype=torch.int32)
    return b


SHAPES = [(8, 8), (4, 8), (4, 4), (2, 4)]   # (block_rows, block_cols) in conv3 pixels; 2-row blocks run on the CTA-pair kernel only


@pytest.mark.parametrize("block_rows,block_cols", SHAPES)
def test_work_list_covers_active_cells(block_rows, block_cols):
    from scene_graph_commonsense_b200 import ops
    boxes = _random_boxes(40, 5)
    n_box = boxes.shape[0]
    sub, obj = np.nonzero(~np.eye(n_box, dtype=bool))
    blocks, n_blocks = ops.conv3_active_blocks(boxes.to(DEV), torch.from_numpy(sub.astype(np.int32)).to(DEV),
                                               torch.from_numpy(obj.astype(np.int32)).to(DEV), block_rows, block_cols=block_cols)
    nb = int(n_blocks.item())
    e = blocks[:nb].cpu().numpy()
    pair, cy, cx = e >> 8, (e >> 4) & 15, e & 15
    hc, wc = block_rows // 2, block_cols // 2
    assert (np.diff(pair) >= 0).all() and pair.max() < len(sub)                  # pairs in order
    assert (cx <= 8 - wc).all() and (cy <= 8 - hc).all()                         # blocks stay inside the 16 x 16 map
    counts = np.bincount(pair, minlength=len(sub))
    assert counts.max() <= 256 // (block_rows * block_cols)                      # never more than the dense tiling
    cover = np.zeros((len(sub), 8, 8), bool)
    for p, y, x in zip(pair, cy, cx):
        cover[p, y:y + hc, x:x + wc] = True
    masks = np.stack([_cell_mask(b) for b in boxes.numpy()])
    want = masks[sub] | masks[obj]
    assert not (want & ~cover).any()
    assert (counts[~want.reshape(len(sub), -1).any(1)] == 0).all()               # two empty boxes: nothing to compute
    # an empty pair list is a no-op with n_blocks == 0
    z = torch.zeros(0, dtype=torch.int32, device=DEV)
    _, n0 = ops.conv3_active_blocks(boxes.to(DEV), z, z, block_rows, block_cols=block_cols)
    assert int(n0.item()) == 0


def _packed(seed=0, gain=1.0):
    from scene_graph_commonsense_b200 import model
    # kernel-level tests below build their operands as explicit bf16 tensors: pin the packed weights to the same 16-bit format
    return model.PackedHead(synthetic.head_state_dict(seed=seed, logit_gain=gain), DEV, operand_dtype=torch.bfloat16)


@pytest.mark.parametrize("block_rows,m_sub,block_cols,cta_pairs", [(8, 2, 8, 0), (4, 2, 8, 0), (8, 1, 8, 0), (4, 1, 8, 0), (4, 2, 4, 0),
                                                                   (4, 1, 4, 0), (4, 2, 4, 1), (2, 2, 4, 1)])
def test_sparse_conv3_equals_dense_bit_for_bit(block_rows, m_sub, block_cols, cta_pairs):
    """conv3_1 + ReLU + pool on the listed blocks over a background pre-fill == the dense kernel, every bf16 bit
    (cta_pairs=1: the tcgen05 cta_group::2 pair kernel, weights on the M side, the tile's pixels shared by the two CTAs)."""
    from scene_graph_commonsense_b200 import ops
    from scene_graph_commonsense_b200._lib import EPI_POOL_BF16, GEMM_CONV3, GEMM_CONV3_BLOCKS
    pk = _packed()
    boxes = _random_boxes(20, 11).to(DEV)
    n_box = boxes.shape[0]
    g = torch.Generator().manual_seed(3)
    t_img = torch.tanh(torch.randn(1, 32 * 32, 256, generator=g)).to(torch.bfloat16).to(DEV)
    abox = ops.box_select(t_img, boxes, torch.zeros(n_box, dtype=torch.int32, device=DEV), pk.fill, 32)
    u, v = pk.conv2_halves(abox)
    sub, obj = np.nonzero(~np.eye(n_box, dtype=bool))
    keep = np.random.default_rng(0).permutation(len(sub))[:150]
    keep[:14] = np.arange(14)                                                   # (0,1) .. : the whole-map box against every edge box
    sub_t = torch.from_numpy(sub[keep].astype(np.int32)).to(DEV)
    obj_t = torch.from_numpy(obj[keep].astype(np.int32)).to(DEV)
    n = sub_t.numel()
    p2 = ops.pair_relu_pool(u, v, None, sub_t, obj_t, 32)
    dense = torch.empty(n, 8, 8, 1024, dtype=torch.bfloat16, device=DEV)
    ops.tc_gemm(p2, pk.w3, dense, n * 256, 1024, 9 * 512, bias=pk.b3, ldc=1024, mode=GEMM_CONV3, epilogue=EPI_POOL_BF16, n_img=n, h=16,
                w=16, c_total=512, c_base=0, c_in=512, group_m=1, m_sub=2)
    blocks, n_blocks = ops.conv3_active_blocks(boxes, sub_t, obj_t, block_rows, block_cols=block_cols)
    sparse = ops.broadcast_rows(pk.p3_background(), n, torch.empty_like(dense))
    assert torch.equal(sparse[n - 1], pk.p3_background()[0])
    ops.tc_gemm(p2, pk.w3, sparse, n * 256, 1024, 9 * 512, bias=pk.b3, ldc=1024, mode=GEMM_CONV3_BLOCKS, epilogue=EPI_POOL_BF16, n_img=n,
                h=16, w=16, c_total=512, c_base=0, c_in=512, m_sub=m_sub, blocks=blocks, n_blocks=n_blocks, block_rows=block_rows,
                block_cols=block_cols, cta_pairs=cta_pairs)
    torch.cuda.synchronize()
    nb = int(n_blocks.item())
    assert 0 < nb < n * (256 // (block_rows * block_cols))                                       # the list is really sparse on these boxes
    bad = (dense.view(torch.int16) != sparse.view(torch.int16)).flatten(1).any(1).nonzero().flatten().tolist()
    assert not bad, "pairs %s differ (boxes %s)" % (bad[:5], [(int(sub[keep][i]), int(obj[keep][i])) for i in bad[:5]])
    # the background really is what a box-free pair produces, and it is not trivially zero
    assert float(pk.p3_background().float().abs().max()) > 0


@pytest.mark.parametrize("block_rows,block_cols", SHAPES)
def test_shared_list_and_assembly_equal_dense_bit_for_bit(block_rows, block_cols):
    """Shared-footprint path: per-box maps ((box, empty) / (empty, box)) + background assembled per pair, conv3_1 only on the
    cover of the cells BOTH boxes reach == the dense kernel on every pair, every bf16 bit; the list covers the intersection."""
    from scene_graph_commonsense_b200 import ops
    from scene_graph_commonsense_b200._lib import EPI_POOL_BF16, GEMM_CONV3
    pk = _packed()
    boxes = _random_boxes(24, 17)
    n_box = boxes.shape[0]
    boxes_x = torch.cat((boxes, boxes.new_zeros(1, 4))).to(DEV)
    g = torch.Generator().manual_seed(4)
    t_img = torch.tanh(torch.randn(1, 32 * 32, 256, generator=g)).to(torch.bfloat16).to(DEV)
    abox = ops.box_select(t_img, boxes_x, torch.zeros(n_box + 1, dtype=torch.int32, device=DEV), pk.fill, 32)
    u, v = pk.conv2_halves(abox)
    sub, obj = np.nonzero(~np.eye(n_box, dtype=bool))
    sub_t = torch.from_numpy(sub.astype(np.int32)).to(DEV)
    obj_t = torch.from_numpy(obj.astype(np.int32)).to(DEV)
    n = sub_t.numel()
    p2 = ops.pair_relu_pool(u, v, None, sub_t, obj_t, 32)
    dense = torch.empty(n, 8, 8, 1024, dtype=torch.bfloat16, device=DEV)
    ops.tc_gemm(p2, pk.w3, dense, n * 256, 1024, 9 * 512, bias=pk.b3, ldc=1024, mode=GEMM_CONV3, epilogue=EPI_POOL_BF16, n_img=n, h=16,
                w=16, c_total=512, c_base=0, c_in=512, group_m=1, m_sub=2)
    # per-box maps through the same kernels
    idx = torch.arange(n_box, dtype=torch.int32, device=DEV)
    emp = torch.full((n_box,), n_box, dtype=torch.int32, device=DEV)
    s1, o1 = torch.cat((idx, emp)), torch.cat((emp, idx))
    p2b = ops.pair_relu_pool(u, v, None, s1, o1, 32)
    blk1, nb1 = ops.conv3_active_blocks(boxes_x, s1, o1, block_rows, block_cols=block_cols)
    maps = ops.broadcast_rows(pk.p3_background(), 2 * n_box, torch.empty(2 * n_box, 8, 8, 1024, dtype=torch.bfloat16, device=DEV))
    pk.conv3_blocks(p2b, maps, 2 * n_box, blk1, nb1, block_rows, block_cols=block_cols, cta_pairs=int(block_rows == 2))
    # work list = cover of the intersection
    blocks, n_blocks = ops.conv3_shared_blocks(boxes_x, sub_t, obj_t, block_rows, block_cols=block_cols)
    nb = int(n_blocks.item())
    e = blocks[:nb].cpu().numpy()
    pair, cy, cx = e >> 8, (e >> 4) & 15, e & 15
    hc, wc = block_rows // 2, block_cols // 2
    cover = np.zeros((n, 8, 8), bool)
    for p, y, x in zip(pair, cy, cx):
        cover[p, y:y + hc, x:x + wc] = True
    masks = np.stack([_cell_mask(b) for b in boxes.numpy()])
    want = masks[sub] & masks[obj]
    assert not (want & ~cover).any()
    assert (np.bincount(pair, minlength=n)[~want.reshape(n, -1).any(1)] == 0).all()    # disjoint reach: nothing per pair
    _, nb_union = ops.conv3_active_blocks(boxes_x, sub_t, obj_t, block_rows, block_cols=block_cols)
    assert 0 < nb < int(nb_union.item())
    # assembly (poisoned buffer: every cell must be written by the assembly or by the listed blocks)
    out = torch.full((n, 8, 8, 1024), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.p3_assemble(pk.p3_background(), maps[:n_box], maps[n_box:], boxes_x[:n_box], sub_t, obj_t, out)
    written = ~torch.isnan(out.float()).any(3).cpu().numpy()          # [n, 8, 8] cells the assembly wrote
    assert (written == ~want).all()
    pk.conv3_blocks(p2, out, n, blocks, n_blocks, block_rows, block_cols=block_cols, cta_pairs=int(block_rows == 2))
    torch.cuda.synchronize()
    bad = (dense.view(torch.int16) != out.view(torch.int16)).flatten(1).any(1).nonzero().flatten().tolist()
    assert not bad, "pairs %s differ (boxes %s)" % (bad[:5], [(int(sub[i]), int(obj[i])) for i in bad[:5]])


@pytest.mark.parametrize("block_rows,shared,block_cols", [(8, False, 8), (4, False, 8), (8, True, 8), (4, True, 8), (4, False, 4), (4, True, 4)])
def test_pipeline_sparse_equals_dense(block_rows, shared, block_cols):
    """Whole forward (chunked + overlapped, and the generic pair-list path): identical raw head outputs and counters."""
    from scene_graph_commonsense_b200 import pipeline
    pk = _packed(gain=40.0)
    samples = synthetic.make_batch([70, 71, 72, 73, 74], [9, 1, 12, 7, 10], p_rel=0.5)
    samples[0].bbox[:6] = torch.tensor(EDGE_BOXES[:6], dtype=samples[0].bbox.dtype)
    outs = []
    for br in (0, block_rows):
        pipe = pipeline.RelationPipeline(pk, DEV, commonsense=True, chunk_pairs=120, conv3_block_rows=br, conv3_shared=shared, fc1_shared=False,
                                         conv3_block_cols=block_cols)
        b = pipeline.batch_from_samples(samples, DEV, skip_mode="per_image")
        pairs = pipe.enumerate_pairs(b)
        rel, sup, conn, logsig = pipe.forward_pairs(b, pairs)
        pipe.evaluate(b, pairs, rel, sup, logsig, connectivity=conn)
        generic = {k: val for k, val in pairs.items() if k != "offsets_host"}
        rel_g = pipe.forward_pairs(b, generic)[0]
        torch.cuda.synchronize()
        outs.append((rel.clone(), sup.clone(), conn.clone(), pipe.counters.clone(), rel_g.clone()))
        if br:
            assert pipe.last_n_blocks is not None and int(pipe.last_n_blocks.sum()) > 0
    for a, c in zip(outs[0], outs[1]):
        assert torch.equal(a, c)
    assert torch.equal(outs[0][0], outs[0][4])


@pytest.mark.parametrize("block_rows,m_sub", [(4, 1), (8, 1), (4, 2)])
def test_conv2_halves_on_the_box_footprint_equal_dense_bit_for_bit(block_rows, m_sub):
    """conv2_1 halves computed only within one pixel of each box over a background pre-fill == the dense halves, every bf16 bit;
    the work list covers the dilated box rectangle with blocks inside the map and lists nothing for an empty box."""
    from scene_graph_commonsense_b200 import ops
    pk = _packed()
    boxes = _random_boxes(30, 23)
    boxes_dev = boxes.to(DEV)
    n_box = boxes.shape[0]
    g = torch.Generator().manual_seed(9)
    t_img = torch.tanh(torch.randn(2, 32 * 32, 256, generator=g)).to(torch.bfloat16).to(DEV)
    box_img = (torch.arange(n_box, dtype=torch.int32) % 2).to(DEV)
    abox = ops.box_select(t_img, boxes_dev, box_img, pk.fill, 32)
    u0, v0 = pk.conv2_halves(abox, m_sub=m_sub)
    u1, v1 = pk.conv2_halves_sparse(abox, boxes_dev, m_sub=m_sub, block_rows=block_rows)
    torch.cuda.synchronize()
    assert torch.equal(u0.view(torch.int16), u1.view(torch.int16))
    assert torch.equal(v0.view(torch.int16), v1.view(torch.int16))
    blocks, n_blocks = ops.conv2_box_blocks(boxes_dev, block_rows)
    e = blocks[:int(n_blocks.item())].cpu().numpy()
    box, oy, ox = e >> 8, 2 * ((e >> 4) & 15), 2 * (e & 15)
    assert (ox + 8 <= 32).all() and (oy + block_rows <= 32).all() and (np.diff(box) >= 0).all()
    cover = np.zeros((n_box, 32, 32), bool)
    for b_, y, x in zip(box, oy, ox):
        cover[b_, y:y + block_rows, x:x + 8] = True
    for i, bx in enumerate(boxes.numpy()):
        x0, x1, y0, y1 = (_slice_bound(int(v)) for v in bx)
        want = np.zeros((32, 32), bool)
        if x1 > x0 and y1 > y0:
            want[max(y0 - 1, 0):y1 + 1, max(x0 - 1, 0):x1 + 1] = True
        assert not (want & ~cover[i]).any()
        if not want.any():
            assert not cover[i].any()
    assert 0 < len(e) < n_box * 4 * (32 // block_rows)


@pytest.mark.parametrize("fmt", ["bf16", "fp16"])
def test_pooling_on_footprint_only_conv2_halves_equals_prefilled(fmt):
    """`conv2_halves_sparse(prefill=False)` writes U / V only on each box's conv2_1 footprint; the pooling kernels (pair-list and
    tiled, with and without the conv3_1 cover) given `uv_footprint` read the background maps elsewhere and produce the same bits as
    on the pre-filled maps - the un-written part is NaN-poisoned here, so any stray read would show."""
    from scene_graph_commonsense_b200 import model, ops, pipeline
    dt = torch.float16 if fmt == "fp16" else torch.bfloat16
    pk = model.PackedHead(synthetic.head_state_dict(seed=0), DEV, operand_dtype=dt)
    samples = synthetic.make_batch([95, 96, 97], [14, 9, 3], with_maps=True)
    samples[0].bbox[:14] = torch.tensor(EDGE_BOXES[:14], dtype=samples[0].bbox.dtype)
    pipe = pipeline.RelationPipeline(pk, DEV, commonsense=False)
    b = pipeline.batch_from_samples(samples, DEV, skip_mode="batch")
    pairs = pipe.enumerate_pairs(b)
    n = pairs["n"]
    boxes_x = torch.cat((b.boxes, b.boxes.new_zeros(1, 4)))
    box_img_x = torch.cat((b.box_img, b.box_img.new_zeros(1)))
    u1, v1 = pipe.box_features(b, boxes_x, box_img_x, prefill=True)
    x = ops.pack_pixels(b.feat, b.depth, model.K1_PAD, dtype=dt)
    t = torch.empty(b.n_images * 1024, 256, dtype=dt, device=DEV)
    ops.tc_gemm(x, pk.w1, t, b.n_images * 1024, 256, model.K1_PAD, bias=pk.b1, lda=model.K1_PAD, ldc=256, epilogue=ops.EPI_BF16, act=ops.ACT_TANH, group_m=8)
    abox = ops.box_select(t, boxes_x, box_img_x, pk.fill, 32)
    u2, v2 = pk.conv2_halves_sparse(abox, boxes_x, prefill=False, poison=True)
    fp = pk.uv_footprint(boxes_x)
    assert torch.isnan(u2.float()).any() and torch.isnan(v2.float()).any()          # the poison really is there
    inside = ~torch.isnan(u2.float()).any(3)
    assert torch.equal(u2[inside].view(torch.int16), u1[inside].view(torch.int16))  # what WAS written equals the complete maps
    # pair-list kernel, plus pairs with the empty box (row n_box) as in `box_maps`
    n_box = b.boxes.shape[0]
    sub = torch.cat((pairs["sub"], torch.arange(n_box, dtype=torch.int32, device=DEV)))
    obj = torch.cat((pairs["obj"], torch.full((n_box,), n_box, dtype=torch.int32, device=DEV)))
    want = ops.pair_relu_pool(u1, v1, None, sub, obj)
    got = ops.pair_relu_pool(u2, v2, None, sub, obj, fp=fp)
    assert not torch.isnan(got.float()).any() and torch.equal(got.view(torch.int16), want.view(torch.int16))
    # tiled kernel, without and with the conv3_1 cover
    n_max = int(np.max(np.diff(b.box_offsets_host)))
    lut = ops.pair_lut_build(pairs["sub"], pairs["obj"], pairs["img"], b.box_offsets, n_box, n_max)
    want_t = ops.pair_relu_pool_tiled(u1, v1, None, b.box_offsets, lut, 0, b.n_images, 0, n)
    got_t = ops.pair_relu_pool_tiled(u2, v2, None, b.box_offsets, lut, 0, b.n_images, 0, n, fp=fp)
    assert not torch.isnan(got_t.float()).any() and torch.equal(got_t.view(torch.int16), want_t.view(torch.int16))
    assert torch.equal(want_t.view(torch.int16), want[:n].view(torch.int16))
    cover = ops.pair_cover_masks(b.boxes, pairs["sub"], pairs["obj"], 4, 4, True)
    a = torch.full_like(want_t, 7.0)
    c = torch.full_like(want_t, 7.0)
    ops.pair_relu_pool_tiled(u1, v1, None, b.box_offsets, lut, 0, b.n_images, 0, n, out=a, cover=cover)
    ops.pair_relu_pool_tiled(u2, v2, None, b.box_offsets, lut, 0, b.n_images, 0, n, out=c, cover=cover, fp=fp)
    torch.cuda.synchronize()
    assert not torch.isnan(c.float()).any() and torch.equal(a.view(torch.int16), c.view(torch.int16))
    # and the whole forward with / without the select agrees bit for bit
    outs = []
    for sel in (True, False):
        p2 = pipeline.RelationPipeline(pk, DEV, commonsense=False, chunk_pairs=150)
        p2.uv_select = sel
        outs.append([z.clone() for z in p2.forward_pairs(b, p2.enumerate_pairs(b))])
    for x1, x2 in zip(*outs):
        assert torch.equal(x1, x2)
