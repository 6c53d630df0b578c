"""CPU-only checks: the C-ABI library loads and exports every symbol include/hiercom_b200.h declares, the product never
touches the oracle or a CPU fallback, and the host-side logic (tables, targets, metrics, sharding, gloo all-reduce)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import hiercom_oracle as O
from scene_graph_commonsense_b200 import _lib, build, synthetic, tables, targets
from scene_graph_commonsense_b200 import dist as hdist
from tests import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "hiercom_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hc_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    build.build()
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), "missing export " + n
        assert n in _lib.SIGNATURES, "ctypes binding missing for " + n
    assert lib.hc_abi_version() == _lib.ABI_VERSION == 5
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.library_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (hc_[a-z0-9_]+)", out))
    assert exported == set(names)


def test_library_is_sm100a_tcgen05_code():
    build.build()
    sass = subprocess.run(["cuobjdump", "-sass", _lib.library_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic + " missing: the dense kernel is not on the tcgen05/TMA path"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_entry_points_fail_loudly_without_a_gpu():
    lib = _lib.load()
    assert lib.hc_device_check() == -4
    assert b"no CPU fallback" in lib.hc_last_error()
    from scene_graph_commonsense_b200 import ops, pipeline
    with pytest.raises(RuntimeError):
        ops.tc_gemm(torch.zeros(128, 64, dtype=torch.bfloat16), torch.zeros(128, 64, dtype=torch.bfloat16), torch.zeros(128, 128), 128, 128, 64, lda=64)
    with pytest.raises(RuntimeError):
        pipeline.RelationPipeline(None, "cpu")


def test_torch_custom_op_layer_registers_every_kernel_entry_point_for_cuda_only():
    """north_star: the Python modules reach the kernels through `torch.ops.hiercom.*`; no operator has a CPU kernel."""
    from scene_graph_commonsense_b200 import ops
    host_only = {"last_error", "abi_version", "device_check", "cs_bitmap_build", "nccl_unique_id", "nccl_comm_create", "nccl_comm_destroy",
                 "pairs_enumerate_workspace_bytes", "conv3_blocks_capacity", "conv2_box_blocks_capacity", "relation_workspace_bytes"}
    compute = {n[3:] for n in _lib.SIGNATURES} - host_only
    folded = {"box_label_embed": "hier_head", "proposals_pack": "detr_proposals", "match_object_categories_fill": "match_object_categories"}
    for name in compute:
        op = folded.get(name, name)
        assert op in ops.SCHEMAS, name
        qual = "hiercom::" + op
        assert torch._C._dispatch_has_kernel_for_dispatch_key(qual, "CUDA"), qual
        assert not torch._C._dispatch_has_kernel_for_dispatch_key(qual, "CPU"), qual
        assert not torch._C._dispatch_has_kernel_for_dispatch_key(qual, "CompositeImplicitAutograd"), qual
    with pytest.raises(NotImplementedError):
        torch.ops.hiercom.topk_select(torch.zeros(2, dtype=torch.int32), torch.zeros(4), 128)
    # caller-owned output buffers are declared mutable in the schema
    assert "Tensor(a!) out" in ops.SCHEMAS["tc_gemm"] and "Tensor(a!) counters" in ops.SCHEMAS["topk_match"]


def test_product_never_imports_the_oracle_or_reference():
    pkg = os.path.join(ROOT, "scene_graph_commonsense_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "hiercom_oracle" not in src, f
                assert "/root/reference" not in src or f == "convert_reference_data.py", f


def test_bitmap_build_matches_numpy_and_dict_semantics():
    from scene_graph_commonsense_b200 import ops
    al, vi = tables.commonsense_aligned_keys(), tables.commonsense_violated_keys()
    assert len(al) == 20884 and len(vi) == 1524 and len(np.intersect1d(al, vi)) == 403
    bm = ops.cs_bitmap_build(al, vi)
    np.testing.assert_array_equal(bm, tables.commonsense_pass_bitmap())
    passing = set(al.tolist()) - set(vi.tolist())
    bits = np.unpackbits(bm.view(np.uint8), bitorder="little")
    assert int(bits.sum()) == len(passing)
    rng = np.random.default_rng(0)
    for k in rng.integers(0, tables.TRIPLET_SPACE, 2000):
        assert bool(bits[k]) == (int(k) in passing)
    with pytest.raises(RuntimeError, match="HC_E_SHAPE"):
        ops.cs_bitmap_build(np.array([tables.TRIPLET_SPACE]), np.zeros(0, np.int64))
    assert len(tables.zero_shot_keys()) == 4314
    assert len(np.intersect1d(tables.zero_shot_keys(), tables.train_triplet_keys())) == 0   # evaluator.py:342 assert


def test_flat_targets_sgd_matches_oracle_restatement():
    batch = [synthetic.make_sgdet_image(i, a, b, p_rel=pr, with_maps=False) for i, a, b, pr in ((1, 6, 9, 0.5), (2, 2, 5, 0.0), (3, 9, 12, 0.7))]
    got = targets.flat_targets_sgd(batch)
    rel, cs, co, bs_, bo_ = O.match_target_sgd(batch)
    base = np.concatenate(([0], np.cumsum([len(s.categories) for s in batch])))
    for i in range(3):
        seg = slice(got["offsets"][i], got["offsets"][i + 1])
        if rel[i] is None:
            assert seg.start == seg.stop
            continue
        np.testing.assert_array_equal(got["label"][seg], rel[i])
        np.testing.assert_array_equal(got["cat"][got["sub"][seg]], cs[i])
        np.testing.assert_array_equal(got["cat"][got["obj"][seg]], co[i])
        np.testing.assert_array_equal(got["box"][got["sub"][seg]], bs_[i])
        np.testing.assert_array_equal(got["box"][got["obj"][seg]], bo_[i])
        assert (got["sub"][seg] >= base[i]).all() and (got["sub"][seg] < base[i + 1]).all()


def test_metrics_from_counters_matches_oracle_float_ops():
    from scene_graph_commonsense_b200 import pipeline
    case = dict(ids=[10, 11, 12], n=[9, 12, 7], run_mode="eval", hierar=True, kw=dict(gain=3.0))
    ev, t3, m, m3, _, _ = helpers.replay_predcls_case(case)
    c = np.concatenate((ev.counters(), t3.counters()))
    got = pipeline.metrics_from_counters(c)
    np.testing.assert_array_equal(helpers.flat_metrics(got["evaluator"]), helpers.flat_metrics(m))
    np.testing.assert_array_equal(helpers.flat_metrics(got["top3"]), helpers.flat_metrics(m3))


def test_synthetic_inputs_are_deterministic_per_image_id():
    a = synthetic.make_batch([5, 6, 7], [6, 4, 9], base_seed=3)
    b = synthetic.make_batch([7, 5], [9, 6], base_seed=3)
    assert torch.equal(a[2].feat, b[0].feat) and torch.equal(a[0].bbox, b[1].bbox)
    assert all(torch.equal(x, y) for x, y in zip(a[2].relationships, b[0].relationships))
    assert a[0].bbox.dtype == torch.int32 and int(a[0].bbox.max()) <= 32 and int(a[0].bbox.min()) >= 0
    area = (a[2].bbox[:, 1] - a[2].bbox[:, 0]) * (a[2].bbox[:, 3] - a[2].bbox[:, 2])
    assert (area[:-1] >= area[1:]).all()                            # area-descending like dataset_utils.py:117
    sd1, sd2 = synthetic.head_state_dict(seed=1, input_dim=16, feature_size=8), synthetic.head_state_dict(seed=1, input_dim=16, feature_size=8)
    assert all(torch.equal(sd1[k], sd2[k]) for k in sd1)


def test_sharding_is_round_robin_and_covers_every_image():
    ids = list(range(23))
    shards = [hdist.shard_image_ids(ids, r, 4) for r in range(4)]
    assert sorted(sum(shards, [])) == ids and shards[1][:3] == [1, 5, 9]


_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from scene_graph_commonsense_b200 import dist as hdist, synthetic, tables
from tests import helpers
rank, local, world = hdist.init_from_env(backend="gloo")
ids = list(range(400, 410)); ns = [6, 9, 4, 12, 7, 10, 3, 8, 11, 5]
mine = hdist.shard_image_ids(list(zip(ids, ns)), rank, world)
case = dict(ids=[i for i, _ in mine], n=[n for _, n in mine], run_mode="eval_cs", hierar=True, kw=dict(gain=3.0), cs=(2, 0.5, 0.1),
            windows=[[j] for j in range(len(mine))])
ev, t3, *_ = helpers.replay_predcls_case(case)
c = torch.from_numpy(np.concatenate((ev.counters(), t3.counters())))
local_copy = c.clone()
g = hdist.allreduce_counters(c)
g2 = hdist.allreduce_counters(c)          # idempotent: the rank-local vector is never modified, a second call gives the same sums
assert torch.equal(c, local_copy) and torch.equal(g, g2)
c = g
t = hdist.max_over_ranks(float(rank + 1), "cpu")
if rank == 0:
    np.save(os.environ["OUT"], c.numpy()); assert t == float(world)
dist.destroy_process_group()
'''


def test_two_rank_gloo_sharded_counts_equal_single_process(tmp_path):
    """World-size-2 `gloo` run of the N>1 host path: shard images round-robin, per-image skip mode, one integer
    all-reduce -> identical counters (hence identical R@K / mR@K) to the single-process run."""
    out = str(tmp_path / "c.npy")
    script = tmp_path / "w.py"
    script.write_text(_WORKER % dict(root=ROOT))
    env = dict(os.environ, OUT=out, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29631", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    ids = list(range(400, 410)); ns = [6, 9, 4, 12, 7, 10, 3, 8, 11, 5]
    case = dict(ids=ids, n=ns, run_mode="eval_cs", hierar=True, kw=dict(gain=3.0), cs=(2, 0.5, 0.1), windows=[[j] for j in range(10)])
    ev, t3, *_ = helpers.replay_predcls_case(case)
    single = np.concatenate((ev.counters(), t3.counters()))
    np.testing.assert_array_equal(np.load(out), single)
    assert single[tables.EV_NGT] > 0


def test_formats_annotation_pkl_to_packed_window_roundtrip(tmp_path):
    """N3: reference-format annotation file -> dataloader transforms -> packed CSR window -> HostBatch arrays."""
    from scene_graph_commonsense_b200 import formats, pipeline
    s = synthetic.make_image(3, 6, p_rel=0.7)
    inv = {v: k for k, v in enumerate(formats.RELATION_FREQ2SCAT[:-1])}           # scat id -> frequency id
    freq_rels = [torch.as_tensor([inv[int(v)] if int(v) >= 0 else -1 for v in r]) for r in s.relationships]
    freq_rels[0][0] = 12                                                            # 'wears' must merge into 'wearing' (4)
    annot = dict(image_depth=s.depth, categories=s.categories, super_categories=s.super_categories, bbox=s.bbox.float() + 0.4,
                 relationships=freq_rels, subj_or_obj=s.subj_or_obj)
    path = str(tmp_path / "1_annotations.pkl")
    torch.save(annot, path)
    got = formats.load_annotation_file(path, image_id=3)
    assert torch.equal(got.bbox, s.bbox)                                            # dataloader.py:129 bbox.int() truncation
    assert int(got.relationships[0][0]) == formats.RELATION_FREQ2SCAT[4]
    for a, b in zip(got.relationships[1:], s.relationships[1:]):
        assert torch.equal(a, b)
    too_many = dict(annot, categories=torch.zeros(21, dtype=torch.int64))
    assert formats.sample_from_annotation(too_many) is None and formats.sample_from_annotation(dict(annot, categories=s.categories[:1])) is None
    got.relationships[0] = s.relationships[0]
    packed = formats.pack_window([got, got])
    formats.save_window(str(tmp_path / "w.npz"), packed)
    hb = formats.host_batch_from_packed(formats.load_window(str(tmp_path / "w.npz")), pinned=False)
    ref = pipeline.host_batch_from_samples([s, s], pinned=False, with_maps=False)
    for k in ("box_offsets", "tri_offsets", "boxes", "cats", "supers", "box_img", "rel_tri", "dir_tri", "group_id"):
        assert torch.equal(hb.t[k], ref.t[k]), k
    assert hb.meta.keys() == ref.meta.keys() and all(np.array_equal(hb.meta[k], ref.meta[k]) for k in ref.meta)
    d = formats.keys_to_commonsense_dict(tables.commonsense_violated_keys()[:50])
    np.testing.assert_array_equal(np.sort(tables.dict_to_keys(d)), np.sort(tables.commonsense_violated_keys()[:50]))


def test_checkpoint_names_cover_both_reference_spellings(tmp_path):
    from scene_graph_commonsense_b200 import formats
    args = synthetic.reference_args(run_mode="eval_cs")
    args["training"]["checkpoint_path"] = str(tmp_path) + "/"
    names = formats.checkpoint_candidates(args, 2)
    assert names[0].endswith("HierRelationModel_CS_motif_2_0.pth") and names[1].endswith("HierRelationModel_CS_motif2_0.pth")
    lin = torch.nn.Linear(4, 3)
    torch.save({"module." + k: v for k, v in lin.state_dict().items()}, names[1])     # trainer's spelling + DDP prefix
    lin2 = torch.nn.Linear(4, 3)
    assert formats.load_checkpoint(lin2, args, 2) == names[1]
    assert torch.equal(lin2.weight, lin.weight)


def test_chunk_picker_partitions_images_and_minimises_fc1_rounds():
    """pipeline._image_chunks (host logic): image-aligned chunks partition the window, respect the pair cap (single
    over-sized images excepted) and never cost more (fc1 rounds + exposed first pooling + launches) than plain greedy chunking at the cap."""
    from scene_graph_commonsense_b200 import pipeline

    class _P:
        _greedy_chunks = staticmethod(pipeline.RelationPipeline._greedy_chunks)
        _chunk_cost = staticmethod(pipeline.RelationPipeline._chunk_cost)
        n_sm = 148

    rounds = lambda ch: sum(-(-(-(-c[3] // 256) * 16) // 148) for c in ch)
    rng = np.random.default_rng(0)
    cases = [(16384, np.full(64, 1560)), (40960, np.full(8, 9240)), (16384, np.full(1, 380)), (100, np.zeros(3, np.int64)),
             (5000, np.array([2, 240, 42, 6, 132, 72, 0, 380, 6000, 90]))]
    cases += [(int(rng.integers(500, 30000)), rng.integers(0, 4000, int(rng.integers(1, 40)))) for _ in range(20)]
    for cap, per_img in cases:
        off = np.concatenate(([0], np.cumsum(per_img)))
        p = _P()
        p.chunk_pairs = cap
        ch = pipeline.RelationPipeline._image_chunks(p, off)
        assert sum(c[3] for c in ch) == off[-1]
        covered = sorted((c[0], c[0] + c[1]) for c in ch)
        for (a0, a1), (b0, b1) in zip(covered, covered[1:]):
            assert a1 <= b0
        for img0, n_img, base, cnt in ch:
            assert base == off[img0] and cnt == off[img0 + n_img] - off[img0] and cnt > 0
            assert cnt <= cap or n_img == 1 or (off[img0 + 1:img0 + n_img + 1] - off[img0:img0 + n_img] > 0).sum() == 1
        cost = pipeline.RelationPipeline._chunk_cost
        assert cost(ch) <= cost(pipeline.RelationPipeline._greedy_chunks(off, cap)) + 1e-9
    p = _P()
    p.chunk_pairs = 16384
    ch = pipeline.RelationPipeline._image_chunks(p, np.arange(65) * 1560)
    assert rounds(ch) == 43 and ch[0][3] == min(c[3] for c in ch)          # cfg2: 43 rounds (10-image chunks needed 45)


def test_training_rows_packing_matches_loop_layout():
    """losses.training_rows_host (vectorised) == the loop restatement of train_test.py:189-258's call structure."""
    from oracle import train_oracle as TO
    from scene_graph_commonsense_b200 import losses
    from tests.golden_cases import TRAIN_CASES
    from tests.helpers import train_case_inputs
    for name in ("tr_hier_plain", "tr_hier_sparse"):
        inp = train_case_inputs(TRAIN_CASES[name])
        s = inp["samples"]
        cat = lambda lst: np.concatenate([np.concatenate([np.asarray(r) for r in x]) for x in lst if len(x)])
        h = losses.training_rows_host(inp["counts"], cat([x.relationships for x in s]), cat([x.subj_or_obj for x in s]))
        assert np.array_equal(h["row_sub"], inp["row_sub"]) and np.array_equal(h["row_obj"], inp["row_obj"])
        assert np.array_equal(h["row_target"], inp["target"])
        groups = inp["groups"]
        assert len(h["group_weight"]) == len(groups)
        for m, rows in enumerate(groups):
            assert h["group_rows"][h["group_offsets"][m]:h["group_offsets"][m + 1]].tolist() == rows.tolist()
        assert h["group_weight"].tolist() == list(range(len(groups), 0, -1))
    # two lock-step batches in one window: weights restart per batch
    h = losses.training_rows_host([3, 2, 4, 2], np.zeros(3 + 1 + 6 + 1, np.int64), -np.ones(11, np.int64), group_size=2)
    assert len(h["group_weight"]) == 6 + 12 and h["group_weight"][:6].tolist() == [6, 5, 4, 3, 2, 1] and h["group_weight"][6] == 12
    g, e = losses.tri_decode(np.arange(0, 5000))
    assert np.array_equal(g * (g - 1) // 2 + e, np.arange(5000)) and (e < g).all() and (e >= 0).all()


def test_fc1_windows_partition_the_pair_list_image_aligned():
    """Host logic of the shared-footprint fc1: windows tile [0, n) in order, chunks tile their window, both on image boundaries."""
    from scene_graph_commonsense_b200.pipeline import RelationPipeline
    pipe = RelationPipeline.__new__(RelationPipeline)
    pipe.chunk_pairs, pipe.chunk_policy, pipe.n_sm = 5000, "waves", 148
    rng = np.random.default_rng(3)
    boxes = rng.integers(0, 60, size=37)
    boxes[[4, 9]] = 0                                           # images without pairs
    off = np.concatenate(([0], np.cumsum(boxes * (boxes - 1).clip(0)))).astype(np.int64)
    n = int(off[-1])
    for cap in (1 << 30, 20000, 3000):
        pipe.fc1_window_pairs = cap
        wins = pipe._fc1_windows({"n": n, "offsets_host": off})
        assert wins[0][0] == 0 and wins[-1][1] == n
        for (a0, a1, _), (b0, b1, _) in zip(wins, wins[1:]):
            assert a1 == b0
        for w0, w1, chunks in wins:
            assert w0 in off and w1 in off
            per_image = np.diff(off)[(off[:-1] >= w0) & (off[1:] <= w1)]
            assert w1 - w0 <= cap or int((per_image > 0).sum()) == 1          # only a single image may exceed the cap
            spans = sorted((c[2], c[2] + c[3]) for c in chunks)
            assert spans[0][0] == w0 and spans[-1][1] == w1
            for (x0, x1), (y0, y1) in zip(spans, spans[1:]):
                assert x1 == y0
            for img0, n_img, base, cnt in chunks:
                assert off[img0] == base and off[img0 + n_img] == base + cnt and cnt > 0
        # generic pair lists (no image structure): fixed-size ranges
        wins = pipe._fc1_windows({"n": n})
        assert [w[:2] for w in wins] == [(s, min(n, s + cap)) for s in range(0, n, cap)]
        for w0, w1, chunks in wins:
            assert [c[2] for c in chunks] == list(range(w0, w1, pipe.chunk_pairs)) and sum(c[3] for c in chunks) == w1 - w0


def test_footprint_cells_contain_the_true_dependency_set_for_every_box_interval():
    """DESIGN 3a rests on one geometric claim: outside `active_cells(lo, hi)` the pooled conv3_1 output cannot depend on the box.
    Brute force per axis (box masks are rectangles and every stage is separable in its support): a pixel "differs from the
    background" inside the box; a 3x3 conv spreads that by one pixel, a 2x2 pool to the cell holding it - model.py:143-146."""
    from tests.test_gpu_sparse import _cells_1d

    def dilate(d):
        out = d.copy()
        out[1:] |= d[:-1]
        out[:-1] |= d[1:]
        return out

    for lo in range(0, 33):
        for hi in range(lo, 33):
            differs = np.zeros(32, bool)
            differs[lo:hi] = True                           # train_test.py:391,398: feature * mask; conv1 is 1x1 (no spread)
            d = dilate(differs)                             # conv2_1 3x3
            d = d.reshape(16, 2).any(1)                     # 2x2 max-pool
            d = dilate(d)                                   # conv3_1 3x3
            d = d.reshape(8, 2).any(1)                      # 2x2 max-pool -> the 8 cells of this axis
            a, b = _cells_1d(lo, hi)
            got = np.zeros(8, bool)
            got[a:b] = True
            assert not (d & ~got).any(), (lo, hi)
            if hi > lo:
                assert (d == got).all(), (lo, hi)           # and it is tight: nothing computed that could not differ


def test_shared_footprint_decomposition_holds_for_the_reference_formulation_in_fp32():
    """The algebra of DESIGN 3a on the REFERENCE formulation itself (train_test.py:391,398 masking + model.py:139-149 in torch fp32 on
    the CPU, no kernel of ours involved): outside the cells both boxes reach, the pooled conv3_1 output of a pair equals the
    (subject, empty) map, the (empty, object) map or the background - exactly - and fc1 of the pair is the sum of the three
    per-box terms plus fc1 of a difference that is zero outside those cells."""
    import torch.nn.functional as F
    from tests.test_gpu_sparse import _cell_mask
    g = torch.Generator().manual_seed(5)
    rnd = lambda *s: torch.randn(*s, generator=g)
    w11, b11, w12, b12 = rnd(128, 257, 1, 1) / 16, rnd(128) / 4, rnd(128, 257, 1, 1) / 16, rnd(128) / 4
    w2, b2 = rnd(512, 256, 3, 3) / 48, rnd(512) / 4
    w3, b3 = rnd(1024, 512, 3, 3) / 68, rnd(1024) / 4
    feat = torch.cat((rnd(256, 32, 32), torch.rand(1, 32, 32, generator=g)))

    def mask_of(box):
        m = torch.zeros(32, 32)
        if box is not None:
            x0, x1, y0, y1 = box
            m[y0:y1, x0:x1] = 1.0
        return m

    def p3(box_s, box_o):
        hs, ho = (feat * mask_of(box_s))[None], (feat * mask_of(box_o))[None]
        a = torch.cat((torch.tanh(F.conv2d(hs, w11, b11)), torch.tanh(F.conv2d(ho, w12, b12))), 1)
        x = F.max_pool2d(F.relu(F.conv2d(a, w2, b2, padding=1)), 2)
        return F.max_pool2d(F.relu(F.conv2d(x, w3, b3, padding=1)), 2)[0]            # [1024, 8, 8]

    bg = p3(None, None)
    w_fc = torch.randn(24, 1024 * 64, generator=g, dtype=torch.float64) / 256
    cases = [((3, 12, 5, 14), (9, 20, 8, 19)), ((0, 6, 0, 6), (20, 32, 22, 32)), ((2, 30, 2, 30), (14, 18, 14, 18)),
             ((0, 32, 15, 17), (15, 17, 0, 32)), ((5, 9, 20, 31), (6, 8, 3, 12))]
    for bs, bo in cases:
        pair, sub, obj = p3(bs, bo), p3(bs, None), p3(None, bo)
        ms, mo = torch.from_numpy(_cell_mask(bs)), torch.from_numpy(_cell_mask(bo))
        assert torch.equal(pair[:, ~ms & ~mo], bg[:, ~ms & ~mo])
        assert torch.equal(pair[:, ms & ~mo], sub[:, ms & ~mo])
        assert torch.equal(pair[:, ~ms & mo], obj[:, ~ms & mo])
        assert torch.equal(sub[:, ~ms], bg[:, ~ms]) and torch.equal(obj[:, ~mo], bg[:, ~mo])
        d = (pair - sub) - (obj - bg)
        assert float(d[:, ~(ms & mo)].abs().max()) == 0.0 if (~(ms & mo)).any() else True
        flat = lambda t: t.double().reshape(-1)
        lhs = w_fc @ flat(pair)
        rhs = w_fc @ flat(sub) + w_fc @ flat(obj) - w_fc @ flat(bg) + w_fc @ flat(d)
        assert float((lhs - rhs).abs().max()) <= 1e-5 * max(1.0, float(lhs.abs().max()))     # d itself is rounded to fp32


def test_cover_fraction_estimate_matches_brute_force_block_count_and_steers_the_dense_fallback():
    """`pipeline.footprint_cover_fraction` (host, numpy) = the share of conv3_1 pixels the shared-footprint work lists visit,
    counted here pair by pair with 2 x 2-cell blocks over the cell rectangle both boxes reach (oracle.parity's cell geometry)."""
    from oracle import parity as PA
    from scene_graph_commonsense_b200 import pipeline, synthetic
    for mode, lo, hi in (("small", 0.10, 0.25), ("vg", 0.35, 0.80), ("full", 1.0, 1.0)):
        samples = synthetic.make_batch([3, 4, 5], [9, 12, 1], with_maps=False, box_mode=mode)
        hb = pipeline.host_batch_from_samples(samples, with_maps=False, pinned=False)
        blocks = pairs = 0
        for s in samples:
            bx = s.bbox.numpy()
            for i in range(len(bx)):
                for j in range(len(bx)):
                    if i == j:
                        continue
                    pairs += 1
                    xs, ys = PA._clip_box(bx[i]), PA._clip_box(bx[j])
                    (ax, bx_), (cx, dx) = PA._cells(xs[0], xs[1]), PA._cells(ys[0], ys[1])
                    (ay, by), (cy, dy) = PA._cells(xs[2], xs[3]), PA._cells(ys[2], ys[3])
                    w, h = max(0, min(bx_, dx) - max(ax, cx)), max(0, min(by, dy) - max(ay, cy))
                    blocks += -(-w // 2) * -(-h // 2)
        want = blocks * 4.0 / (64.0 * pairs)
        got = hb.meta["cover_fraction"]
        assert abs(got - want) < 1e-12, (mode, got, want)
        assert lo <= got <= hi, (mode, got)
    assert pipeline.footprint_cover_fraction(np.zeros((0, 4)), np.array([0])) == 0.0


def test_gt_from_ranking_places_ranked_triplets_once_per_unordered_pair():
    from scene_graph_commonsense_b200 import synthetic
    s = synthetic.make_image(11, 12, with_maps=False)
    n_before = sum(int((r >= 0).sum()) for r in s.relationships)
    ranked = [(5, 2, 7), (2, 5, 9), (0, 3, 1), (11, 10, 49), (4, 4, 3), (1, 0, 20)]
    s = synthetic.assign_gt_from_ranking(s, ranked, p_keep=1.0, thin=0.0)
    got = {}
    for gi in range(1, 12):
        for e in range(gi):
            if int(s.relationships[gi - 1][e]) >= 0:
                d = int(s.subj_or_obj[gi - 1][e])
                got[(gi, e)] = ((gi, e) if d == 1 else (e, gi)) + (int(s.relationships[gi - 1][e]),)
    # (2,5,9) lost to the better-ranked (5,2,7) on the same unordered pair; (4,4,.) is not a pair
    assert got == {(5, 2): (5, 2, 7), (3, 0): (0, 3, 1), (11, 10): (11, 10, 49), (1, 0): (1, 0, 20)}
    assert all(int(d) == -1 for dd, rr in zip(s.subj_or_obj, s.relationships) for d, r in zip(dd, rr) if int(r) < 0)
    s2 = synthetic.assign_gt_from_ranking(synthetic.make_image(11, 12, with_maps=False), ranked)          # defaults: thinned + 60 %
    n_after = sum(int((r >= 0).sum()) for r in s2.relationships)
    assert 0 < n_after < n_before
    s3 = synthetic.assign_gt_from_ranking(synthetic.make_image(11, 12, with_maps=False), ranked)
    assert all(torch.equal(a, b) for a, b in zip(s2.relationships, s3.relationships))                     # deterministic


def test_host_pair_offsets_equal_the_oracle_enumeration_in_both_skip_modes():
    """`pipeline.host_pair_offsets` (what lets a step run without a device -> host read) == the per-image directed-pair counts of
    the oracle's replay of evaluate.py:132-156, for the per-image and the whole-batch skip rule, ragged images, degenerate,
    negative and out-of-range boxes."""
    from oracle import hiercom_oracle as O
    from scene_graph_commonsense_b200 import pipeline, synthetic
    samples = synthetic.make_batch([21, 22, 23, 24, 25], [9, 1, 14, 2, 7], with_maps=False)
    samples[0].bbox[:5] = torch.tensor([(0, 32, 0, 32), (5, 5, 3, 9), (9, 3, 2, 7), (-4, 32, 3, 12), (30, 40, -2, 2)], dtype=samples[0].bbox.dtype)
    for mode, gs in (("per_image", None), ("batch", None), ("batch", 2)):
        hb = pipeline.host_batch_from_samples(samples, skip_mode=mode, group_size=gs, with_maps=False, pinned=False)
        got = np.diff(hb.meta["pair_offsets"])
        groups = [list(range(len(samples)))] if gs is None else [list(range(i, min(i + gs, len(samples)))) for i in range(0, len(samples), gs)]
        want = np.zeros(len(samples), dtype=np.int64)
        for grp in (groups if mode == "batch" else [[i] for i in range(len(samples))]):
            masks = [[O.box_mask(b) for b in samples[i].bbox] for i in grp]
            n_max = max(len(m) for m in masks)
            for g in range(1, n_max):
                for e in range(g):
                    has = [k for k, m in enumerate(masks) if len(m) > g]
                    if any(bool((masks[k][g] & masks[k][e]).any()) for k in has):
                        for k in has:
                            want[grp[k]] += 2
        np.testing.assert_array_equal(got, want)


def test_sparse_box_rows_decomposition_and_tile_order_on_the_cpu():
    """Host-side checks of the K-cell-sparse per-box fc1 rows (no kernel involved, torch fp64 on the CPU): for maps that equal the
    background outside a tile's cell mask, fc1(map) == sum over the mask's cells of W_c.map_c + sum over the other cells of W_c.bg_c -
    the per-tile constant `PackedHead.fc1_rows_sparse` adds - and `PackedHead.longest_first` orders tiles by cell count, stable."""
    from scene_graph_commonsense_b200.model import PackedHead
    g = torch.Generator().manual_seed(3)
    cells, kc, n_out, rows = 64, 8, 16, 6
    w = torch.randn(n_out, cells, kc, generator=g, dtype=torch.float64)
    bg = torch.randn(cells, kc, generator=g, dtype=torch.float64)
    mask_bits = torch.zeros(cells, dtype=torch.bool)
    mask_bits[[3, 4, 11, 12, 63]] = True
    maps = bg.repeat(rows, 1, 1)
    maps[:, mask_bits] = torch.randn(rows, int(mask_bits.sum()), kc, generator=g, dtype=torch.float64)      # differ only inside the mask
    dense = torch.einsum("nck,rck->rn", w, maps)
    per_cell_bg = torch.einsum("nck,ck->cn", w, bg)                                                          # fc1_background_cells
    sparse = torch.einsum("nck,rck->rn", w[:, mask_bits], maps[:, mask_bits]) + per_cell_bg[~mask_bits].sum(0)
    assert float((dense - sparse).abs().max()) < 1e-10
    masks = torch.tensor([0b1011, 0, -1, 0b1, 0b1110, -(1 << 63)], dtype=torch.int64)     # 3, 0, 64, 1, 3, 1 cells (bit 63 alone = int64 min)
    order = PackedHead.longest_first(masks).tolist()
    assert order == [2, 0, 4, 3, 5, 1]
