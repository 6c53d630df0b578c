"""GPU tests of the shared-footprint fc1: fc1 (model.py:149) is linear, so a pair's fc1 = fc1(subject map) + fc1(object map)
- fc1(background) + fc1(d), d = the pair's pooled conv3_1 output minus those maps (zero outside the cells both boxes reach), and
the GEMM over d visits only the K cells a tile's rows use.  Bars: the K-cell-sparse GEMM equals the dense GEMM on the same operand
bit for bit; the difference epilogue equals the fp32 formula on the dense kernel's output bit for bit; the whole path stays within
north_star's 2e-3 on joint probabilities of the dense path, whose parity with the fp32 reference tests/test_gpu_model.py holds."""
import numpy as np
import pytest
import torch

from scene_graph_commonsense_b200 import synthetic
from tests.test_gpu_sparse import EDGE_BOXES, _cell_mask, _packed, _random_boxes

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("m,with_tables", [(700, False), (700, True), (100, True), (1024, True)])
def test_k_cell_sparse_gemm_equals_dense_bit_for_bit(m, with_tables):
    from scene_graph_commonsense_b200 import ops
    from scene_graph_commonsense_b200._lib import ACT_RELU, EPI_BF16, EPI_F32
    g = torch.Generator().manual_seed(m)
    n_cells, cell, n = 16, 128, 512
    k = n_cells * cell
    a = torch.randn(m, n_cells, cell, generator=g)
    use = torch.rand(m, n_cells, generator=g) < 0.25
    use[256:512] = False                                   # a whole 256-row tile with nothing to do
    if m > 600:
        use[600:, :] = False
        use[600:, 3] = True
    a = (a * use[:, :, None]).to(torch.bfloat16).to(DEV)
    w = (torch.randn(n, k, generator=g) / 16).to(torch.bfloat16).to(DEV)
    bias = torch.randn(n, generator=g).to(DEV)
    n_tiles = -(-m // 256)
    masks = torch.zeros(n_tiles, dtype=torch.int64)
    for t in range(n_tiles):
        bits = use[256 * t:256 * (t + 1)].any(0)
        masks[t] = int(sum(1 << c for c in range(n_cells) if bits[c]))
    assert int(masks[1]) == 0 if n_tiles > 1 else True
    masks = masks.to(DEV)
    kw = {}
    if with_tables:
        fa, fb = torch.randn(37, n, generator=g).to(DEV), torch.randn(41, n, generator=g).to(DEV)
        ia = torch.randint(0, 37, (m,), generator=g, dtype=torch.int32).to(DEV)
        ib = torch.randint(0, 41, (m,), generator=g, dtype=torch.int32).to(DEV)
        kw = dict(add_a=fa, add_a_rows=ia, add_b=fb, add_b_rows=ib)
    outs = []
    for km in (None, masks):
        o = torch.full((m, n), float("nan"), dtype=torch.bfloat16, device=DEV)
        ops.tc_gemm(a, w, o, m, n, k, bias=bias, lda=k, ldc=n, epilogue=EPI_BF16, act=ACT_RELU, group_m=3, m_sub=2,
                    k_masks=km, k_cell=cell if km is not None else 0, **kw)
        outs.append(o)
    torch.cuda.synchronize()
    assert torch.equal(outs[0].view(torch.int16), outs[1].view(torch.int16))
    ref = a.float().view(m, k) @ w.float().t() + bias
    if with_tables:
        ref = ref + fa[ia.long()] + fb[ib.long()]
    ref = torch.relu(ref)
    assert float((outs[1].float() - ref).abs().max()) <= 0.02 * float(ref.abs().max()) + 1e-2
    # row map of the f32 epilogue (fc2 writes raw back in pair order)
    perm = torch.randperm(m, generator=g).to(torch.int32).to(DEV)
    o1 = torch.empty(m, n, dtype=torch.float32, device=DEV)
    o2 = torch.full((m, n), float("nan"), dtype=torch.float32, device=DEV)
    ops.tc_gemm(a, w, o1, m, n, k, lda=k, ldc=n, epilogue=EPI_F32, m_sub=2)
    ops.tc_gemm(a, w, o2, m, n, k, lda=k, ldc=n, epilogue=EPI_F32, m_sub=2, out_rows=perm, k_masks=masks, k_cell=cell)
    torch.cuda.synchronize()
    assert torch.equal(o2[perm.long()], o1)


@pytest.mark.parametrize("block_rows,block_cols,cta_pairs", [(8, 8, 0), (4, 8, 0), (4, 4, 0), (4, 4, 1), (2, 4, 1)])
def test_difference_epilogue_and_keys(block_rows, block_cols, cta_pairs):
    """HC_EPI_POOL_DIFF_BF16 == (x - sub_map) - (obj_map - background) on the dense kernel's x, bit for bit, at row pair_row[i]; zero in
    every covered cell that only one box reaches; keys / tile masks describe the cell rectangles both boxes reach."""
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
    idx = torch.arange(n_box, dtype=torch.int32, device=DEV)
    emp = torch.full((n_box,), n_box, dtype=torch.int32, device=DEV)
    s1, o1 = torch.cat((idx, emp)), torch.cat((emp, idx))
    p2b = ops.pair_relu_pool(u, v, None, s1, o1, 32)
    blk1, nb1 = ops.conv3_active_blocks(boxes_x, s1, o1, block_rows, block_cols=block_cols)
    maps = ops.broadcast_rows(pk.p3_background(), 2 * n_box, torch.empty(2 * n_box, 8, 8, 1024, dtype=torch.bfloat16, device=DEV))
    pk.conv3_blocks(p2b, maps, 2 * n_box, blk1, nb1, block_rows, block_cols=block_cols, cta_pairs=cta_pairs)
    sub_maps, obj_maps, bg = maps[:n_box], maps[n_box:], pk.p3_background()
    # keys and sorted order
    keys = ops.pair_cell_keys(boxes_x, sub_t, obj_t).cpu().numpy()
    masks_np = np.stack([_cell_mask(b) for b in boxes.numpy()])
    want = masks_np[sub] & masks_np[obj]                                       # [n, 8, 8]
    for p in range(n):
        if not want[p].any():
            assert keys[p] == 4096
        else:
            ys, xs = np.nonzero(want[p].any(1))[0], np.nonzero(want[p].any(0))[0]
            assert keys[p] == ((ys[0] * 8 + ys[-1]) * 8 + xs[0]) * 8 + xs[-1]
    perm = torch.sort(torch.from_numpy(keys).to(DEV), stable=True)[1]
    row_of = torch.empty(n, dtype=torch.int32, device=DEV)
    row_of[perm] = torch.arange(n, dtype=torch.int32, device=DEV)
    tm = ops.tile_cell_masks(boxes_x, sub_t[perm].contiguous(), obj_t[perm].contiguous(), 256).cpu().numpy()
    order = perm.cpu().numpy()
    for t in range(len(tm)):
        bits = want[order[256 * t:256 * (t + 1)]].any(0).reshape(-1)
        assert int(tm[t]) == int(sum(1 << c for c in range(64) if bits[c])) - (1 << 64 if bits[63] else 0)
    # difference epilogue into a poisoned, then tile-zeroed buffer
    d = torch.full((n, 64, 1024), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.cells_zero(torch.from_numpy(tm).to(DEV), 256, n, d)
    blocks, n_blocks = ops.conv3_shared_blocks(boxes_x, sub_t, obj_t, block_rows, block_cols=block_cols)
    pk.conv3_diff(p2, d.view(n, 8, 8, 1024), n, blocks, n_blocks, block_rows, sub_maps, obj_maps, sub_t, obj_t, row_of,
                  block_cols=block_cols, cta_pairs=cta_pairs)
    torch.cuda.synchronize()
    ref = ((dense.float() - sub_maps[sub_t.long()].float()) - (obj_maps[obj_t.long()].float() - bg.float())).to(torch.bfloat16)
    got = d.view(n, 8, 8, 1024)[row_of.long()]                                 # back in pair order
    want_t = torch.from_numpy(want).to(DEV)
    tile_bits = torch.from_numpy(np.stack([[(int(tm[t]) >> c) & 1 for c in range(64)] for t in range(len(tm))]).astype(bool)).to(DEV)
    visited = tile_bits[(row_of.long() // 256)].view(n, 8, 8)                   # cells the GEMM will read for this pair's row
    assert not torch.isnan(got.float())[visited].any()                         # every visited cell was written (zero fill or epilogue)
    both = want_t[..., None].expand_as(ref)
    assert torch.equal(got[both].view(torch.int16), ref[both].view(torch.int16))
    only_visited = (visited & ~want_t)[..., None].expand_as(ref)
    assert float(got[only_visited].float().abs().max()) == 0.0                 # one-box / background cells: exactly zero
    assert float(ref[~both].float().abs().max()) == 0.0                        # and the dense formula agrees they are zero


@pytest.mark.parametrize("tiled,block_cols,block_rows", [(True, 4, 4), (False, 4, 4), (True, 8, 4), (True, 4, 2), (False, 4, 2)])
def test_pipeline_fc1_shared_within_tolerance_of_dense(tiled, block_cols, block_rows):
    """Whole forward, chunked and overlapped: joint probabilities within 2e-3 of the dense path (north_star's fp tolerance)."""
    from scene_graph_commonsense_b200 import pipeline
    pk = _packed(gain=40.0)
    samples = synthetic.make_batch([70, 71, 72, 73, 74, 75], [9, 1, 12, 7, 10, 40], p_rel=0.5)
    samples[0].bbox[:6] = torch.tensor(EDGE_BOXES[:6], dtype=samples[0].bbox.dtype)
    outs = []
    for fc1_shared in (False, True):
        pipe = pipeline.RelationPipeline(pk, DEV, commonsense=True, chunk_pairs=700, conv3_block_rows=block_rows if fc1_shared else 4,
                                         conv3_shared=True, fc1_shared=fc1_shared, conv3_block_cols=block_cols)
        pipe.debug_poison = True                  # footprint-aware pooling: a pixel conv3_1 reads but nobody wrote would be NaN
        b = pipeline.batch_from_samples(samples, DEV, skip_mode="per_image")
        pairs = pipe.enumerate_pairs(b)
        if not tiled:
            pairs = {k: val for k, val in pairs.items() if k != "offsets_host"}
        rel, sup, conn, logsig = pipe.forward_pairs(b, pairs)
        torch.cuda.synchronize()
        outs.append((rel.clone(), sup.clone(), conn.clone()))
        if fc1_shared:
            assert pipe.last_k_masks is not None and pipe.last_k_masks.numel() == -(-pairs["n"] // 256)
    (rel0, sup0, conn0), (rel1, sup1, conn1) = outs
    assert not torch.isnan(rel1).any() and not torch.isnan(conn1).any()
    assert float((rel0.exp() - rel1.exp()).abs().max()) <= 2e-3
    assert float((sup0.exp() - sup1.exp()).abs().max()) <= 2e-3
    assert float((torch.sigmoid(conn0) - torch.sigmoid(conn1)).abs().max()) <= 2e-3


@pytest.mark.parametrize("tiled", [True, False])
def test_fc1_windows_are_bit_identical_to_one_window(tiled):
    """A batch cut into several shared-fc1 windows (operand bounded at 128 KB per pair) gives the same bits as one window: a row's
    GEMM result does not depend on the tile, the sort order or the window it lands in."""
    from scene_graph_commonsense_b200 import pipeline
    pk = _packed(gain=40.0)
    samples = synthetic.make_batch([80, 81, 82, 83, 84], [11, 14, 3, 16, 9], p_rel=0.5)
    outs = []
    for cap in (262144, 40):
        pipe = pipeline.RelationPipeline(pk, DEV, commonsense=True, chunk_pairs=130)
        pipe.fc1_window_pairs = cap
        b = pipeline.batch_from_samples(samples, DEV, skip_mode="per_image")
        pairs = pipe.enumerate_pairs(b)
        if not tiled:
            pairs = {k: val for k, val in pairs.items() if k != "offsets_host"}
        rel, sup, conn, logsig = pipe.forward_pairs(b, pairs)
        pipe.evaluate(b, pairs, rel, sup, logsig, connectivity=conn)
        torch.cuda.synchronize()
        outs.append((rel.clone(), sup.clone(), conn.clone(), pipe.counters.clone(), int(pipe.last_k_masks.numel())))
    assert outs[1][4] > outs[0][4]                       # really several windows
    for a, c in zip(outs[0][:4], outs[1][:4]):
        assert torch.equal(a, c)


@pytest.mark.parametrize("block_rows,block_cols,shared", [(4, 4, True), (4, 8, True), (8, 8, False), (2, 4, True)])
def test_footprint_pooling_writes_exactly_what_the_blocks_read(block_rows, block_cols, shared):
    """`pair_cover_masks` == the union of the listed blocks of each pair; the tiled pooling with `cover` writes exactly the pooled
    pixels within one pixel of a covered cell, with the values of the full pooling."""
    from scene_graph_commonsense_b200 import ops, pipeline
    pk = _packed()
    samples = synthetic.make_batch([90, 91], [12, 9], p_rel=0.5)
    samples[0].bbox[:6] = torch.tensor(EDGE_BOXES[:6], dtype=samples[0].bbox.dtype)
    pipe = pipeline.RelationPipeline(pk, DEV, commonsense=False)
    b = pipeline.batch_from_samples(samples, DEV, skip_mode="batch")
    pairs = pipe.enumerate_pairs(b)
    n = pairs["n"]
    u, v = pipe.box_features(b)
    lister = ops.conv3_shared_blocks if shared else ops.conv3_active_blocks
    blocks, n_blocks = lister(b.boxes, pairs["sub"], pairs["obj"], block_rows, block_cols=block_cols)
    e = blocks[:int(n_blocks.item())].cpu().numpy()
    want = np.zeros((n, 8, 8), bool)
    for p, y, x in zip(e >> 8, (e >> 4) & 15, e & 15):
        want[p, y:y + block_rows // 2, x:x + block_cols // 2] = True
    cover = ops.pair_cover_masks(b.boxes, pairs["sub"], pairs["obj"], block_rows, block_cols, shared)
    got = cover[:n].cpu().numpy()
    bits = ((got[:, None].astype(np.uint64) >> np.arange(64, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(bool).reshape(n, 8, 8)
    assert (bits == want).all()
    n_box = b.boxes.shape[0]
    n_max = int(np.max(np.diff(b.box_offsets_host)))
    lut = ops.pair_lut_build(pairs["sub"], pairs["obj"], pairs["img"], b.box_offsets, n_box, n_max)
    full = ops.pair_relu_pool_tiled(u, v, None, b.box_offsets, lut, 0, 2, 0, n)
    part = torch.full_like(full, float("nan"))
    ops.pair_relu_pool_tiled(u, v, None, b.box_offsets, lut, 0, 2, 0, n, out=part, cover=cover)
    torch.cuda.synchronize()
    need = np.zeros((n, 16, 16), bool)              # pooled pixels within one pixel of a covered cell
    for p, cy, cx in zip(*np.nonzero(want)):
        need[p, max(2 * cy - 1, 0):2 * cy + 3, max(2 * cx - 1, 0):2 * cx + 3] = True
    written = ~torch.isnan(part.float()).any(3).cpu().numpy()
    assert (written == need).all()
    m = torch.from_numpy(need).to(DEV)
    assert torch.equal(part[m].view(torch.int16), full[m].view(torch.int16))
    # the generic (pair-list) pooling kernel takes the same cover
    part2 = torch.full_like(full, float("nan"))
    ops.pair_relu_pool(u, v, None, pairs["sub"], pairs["obj"], out=part2, cover=cover)
    torch.cuda.synchronize()
    assert ((~torch.isnan(part2.float()).any(3)).cpu().numpy() == need).all()
    assert torch.equal(part2[m].view(torch.int16), full[m].view(torch.int16))


@pytest.mark.parametrize("m,k_sparse", [(700, True), (1024, True), (1300, True), (300, False), (5121, False)])
def test_plain_gemm_on_cta_pairs_equals_single_cta_bit_for_bit(m, k_sparse):
    """PLAIN GEMM on tcgen05 cta_group::2 pairs (a pair owns one 256 x 256 tile: 128 rows of A and 128 columns of B per CTA and K
    step): same K order per output element, so every bit equals the single-CTA kernel - dense, K-cell-sparse with row gathers and an
    output row map (empty-mask tiles included), f32 and 16-bit epilogues, ragged M."""
    from scene_graph_commonsense_b200 import ops
    g = torch.Generator().manual_seed(5)
    n, k, cell = 512, 16 * 256, 256
    a = (torch.randn(m, k, generator=g) * 0.3).to(torch.bfloat16).to(DEV)
    w = (torch.randn(n, k, generator=g) * 0.05).to(torch.bfloat16).to(DEV)
    bias = torch.randn(n, generator=g).to(DEV)
    kw = {}
    if k_sparse:
        tiles = -(-m // 256)
        masks = torch.randint(0, 1 << 16, (tiles,), generator=g, dtype=torch.int64)
        masks[0] = 0                                                        # an empty mask: no MMA, the epilogue substitutes zeros
        for t in range(tiles):                                              # operand is zero wherever the tile's mask skips
            for c in range(16):
                if not (int(masks[t]) >> c) & 1:
                    a[256 * t:256 * (t + 1), cell * c:cell * (c + 1)] = 0
        fa, fb = torch.randn(9, n, generator=g).to(DEV), torch.randn(7, n, generator=g).to(DEV)
        ra = torch.randint(0, 9, (m,), generator=g, dtype=torch.int32).to(DEV)
        rb = torch.randint(0, 7, (m,), generator=g, dtype=torch.int32).to(DEV)
        kw = dict(k_masks=masks.to(DEV), k_cell=cell, add_a=fa, add_a_rows=ra, add_b=fb, add_b_rows=rb)
    outs = []
    # (k_masks are per 256 rows: the single-CTA tile of m_sub 2, or one unit of a pair tile - m_sub 2 pairs walk the union of two units)
    for pairs, m_sub in (((0, 2), (1, 2), (1, 1)) if k_sparse else ((0, 2), (0, 1), (1, 2), (1, 1))):
        o16 = torch.full((m, n), float("nan"), dtype=torch.bfloat16, device=DEV)
        ops.tc_gemm(a, w, o16, m, n, k, bias=bias, lda=k, ldc=n, epilogue=ops.EPI_BF16, act=ops.ACT_RELU, group_m=3, m_sub=m_sub, cta_pairs=pairs, **kw)
        o32 = torch.full((m, n), float("nan"), device=DEV)
        perm = torch.randperm(m, generator=g).to(torch.int32).to(DEV)
        kw32 = {q: v for q, v in kw.items() if q in ("k_masks", "k_cell")}
        ops.tc_gemm(a, w, o32, m, n, k, lda=k, ldc=n, epilogue=ops.EPI_F32, group_m=2, m_sub=m_sub, cta_pairs=pairs, out_rows=perm, **kw32)
        torch.cuda.synchronize()
        assert not torch.isnan(o16.float()).any() and not torch.isnan(o32).any()
        outs.append((o16.clone(), o32[perm.long()].clone()))
    # any visiting order of the M tiles gives the same bits (the pipeline passes the K-cell-sparse tiles longest-first)
    order = torch.randperm(-(-m // 256), generator=g).to(torch.int32).to(DEV)
    o16 = torch.full((m, n), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.tc_gemm(a, w, o16, m, n, k, bias=bias, lda=k, ldc=n, epilogue=ops.EPI_BF16, act=ops.ACT_RELU, group_m=3, m_sub=1, cta_pairs=1, m_order=order, **kw)
    torch.cuda.synchronize()
    assert torch.equal(o16.view(torch.int16), outs[-1][0].view(torch.int16))
    with pytest.raises(RuntimeError, match="m_order"):
        ops.tc_gemm(a, w, o16, m, n, k, bias=bias, lda=k, ldc=n, epilogue=ops.EPI_BF16, m_sub=1, cta_pairs=1, m_order=order[:-1].contiguous() if order.numel() > 1 else order.to(torch.int64))
    for o16, o32 in outs[:-1]:
        assert torch.equal(o16.view(torch.int16), outs[-1][0].view(torch.int16))
        assert torch.equal(o32, outs[-1][1])
    ref = a.float() @ w.float().t()
    assert float((outs[-1][1] - ref).abs().max()) <= 2e-3 * float(ref.abs().max())


def test_sparse_box_maps_and_fc1_rows_equal_the_dense_per_box_path():
    """`box_maps_sparse`: the per-box maps (produced in sorted order) are bit-identical to `box_maps`, and their fc1 rows from the
    K-cell-sparse GEMM + the per-tile background constant equal the dense per-box fc1 rows up to fp32 summation order; the whole
    forward with and without it agrees within 2e-4 on joint probabilities."""
    from scene_graph_commonsense_b200 import pipeline
    pk = _packed(gain=40.0)
    samples = synthetic.make_batch([80, 81, 82], [11, 6, 14], p_rel=0.5)
    samples[0].bbox[:6] = torch.tensor(EDGE_BOXES[:6], dtype=samples[0].bbox.dtype)
    b = pipeline.batch_from_samples(samples, DEV, skip_mode="per_image")
    pipe = pipeline.RelationPipeline(pk, DEV, commonsense=True, chunk_pairs=700)
    assert pipe.fc1_box_sparse
    n_box = b.boxes.shape[0]
    boxes_x = torch.cat((b.boxes, b.boxes.new_zeros(1, 4)))
    u, v = pipe.box_features(b, boxes_x, torch.cat((b.box_img, b.box_img.new_zeros(1))))
    maps_d, _ = pipe.box_maps(boxes_x, u, v, with_background_row=True)
    f_dense = pk.fc1_rows(maps_d, 2 * n_box + 1)
    maps_s, map_row, f_sparse, _ = pipe.box_maps_sparse(boxes_x, u, v)
    torch.cuda.synchronize()
    assert sorted(map_row.tolist()) == list(range(2 * n_box))
    assert torch.equal(maps_s[map_row.long()].view(torch.int16), maps_d[:2 * n_box].view(torch.int16))
    assert torch.equal(maps_s[2 * n_box].view(torch.int16), maps_d[2 * n_box].view(torch.int16))
    scale = float(f_dense.abs().max())
    assert float((f_sparse - f_dense[:2 * n_box]).abs().max()) <= 2e-5 * scale        # same products, fp32 summation order only
    assert float((pk.fc1_background() - f_dense[2 * n_box]).abs().max()) <= 1e-5 * scale
    assert float((pk.fc1_background_cells().sum(0) - f_dense[2 * n_box]).abs().max()) <= 1e-5 * scale
    outs = []
    for sparse in (True, False):
        pipe.fc1_box_sparse = sparse
        pairs = pipe.enumerate_pairs(b)
        outs.append([t.clone() for t in pipe.forward_pairs(b, pairs)])
    assert float((outs[0][0].exp() - outs[1][0].exp()).abs().max()) <= 2e-4
    assert float((outs[0][1].exp() - outs[1][1].exp()).abs().max()) <= 2e-4
