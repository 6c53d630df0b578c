"""GPU parity of the SGDET / SGCLS proposal front-end against goldens produced by executing the
reference's own evaluate.py:311-370 lines, utils.match_object_categories and utils.match_target_sgd
(oracle/make_golden_frontend.py), and against the oracle end to end.  Integer outputs (labels, NMS keep sets and
order, matched labels, targets, counters) are BIT-EXACT; fp32 box arithmetic and `conf * iou` are bit-exact; softmax
confidences are within 1e-6 absolute (exp implementation differs)."""
import numpy as np
import pytest
import torch

from oracle import frontend_oracle as FO
from oracle import hiercom_oracle as O
from scene_graph_commonsense_b200 import synthetic, tables
from tests import helpers
from tests.golden_cases import FRONTEND_CASES

pytestmark = pytest.mark.gpu
DEV = "cuda"
CONF_ATOL = 1e-6


def _split(flat, off):
    return [flat[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def _samples(case):
    return synthetic.make_batch(case["ids"], case["n_gt"], with_maps=False, p_rel=0.5)


@pytest.mark.parametrize("name", sorted(FRONTEND_CASES))
def test_detr_proposals_match_reference_golden(name):
    from scene_graph_commonsense_b200 import frontend
    g = helpers.golden(name)
    out_dict = {"pred_logits": torch.from_numpy(g["pred_logits"]).to(DEV), "pred_boxes": torch.from_numpy(g["pred_boxes"]).to(DEV)}
    p = frontend.detr_proposals(out_dict, synthetic.reference_args())
    np.testing.assert_array_equal(p.offsets_host, g["offsets"])
    np.testing.assert_array_equal(p.cats.cpu().numpy(), g["cats"])
    np.testing.assert_array_equal(p.box_f.cpu().numpy(), g["bbox"])
    np.testing.assert_allclose(p.conf.cpu().numpy(), g["conf"], rtol=0, atol=CONF_ATOL)
    np.testing.assert_array_equal(p.supers.cpu().numpy(), g["supers"])
    np.testing.assert_array_equal(p.box_i.cpu().numpy(), g["bbox"].astype(np.int32))
    np.testing.assert_array_equal(p.box_img.cpu().numpy(), np.repeat(np.arange(len(g["offsets"]) - 1), np.diff(g["offsets"])))
    # rasterised mask area of each kept proposal == reference masks_pred (evaluate.py:336-340)
    bi = p.box_i.cpu().numpy()
    area = np.clip(bi[:, 1] - bi[:, 0], 0, None) * np.clip(bi[:, 3] - bi[:, 2], 0, None)
    np.testing.assert_array_equal(area, g["masks_sum"])
    cats, conf, boxes, sup = p.to_lists()
    assert len(cats) == len(g["offsets"]) - 1 and all(len(c) == len(b) == len(s) for c, b, s in zip(cats, boxes, sup))


@pytest.mark.parametrize("name", sorted(FRONTEND_CASES))
def test_match_object_categories_matches_reference_golden(name):
    from scene_graph_commonsense_b200 import frontend
    g = helpers.golden(name)
    off = g["offsets"]
    samples = _samples(FRONTEND_CASES[name])
    t = lambda a: [torch.from_numpy(np.ascontiguousarray(x)).to(DEV) for x in a]
    cats, conf, box = frontend.match_object_categories(t(_split(g["cats"], off)), t(_split(g["conf"], off)), t(_split(g["bbox"], off)),
                                                       [s.bbox.to(DEV) for s in samples])
    if bool(g["moc_none"]):
        assert cats is None and conf is None and box is None
        return
    assert [len(c) for c in cats] == np.diff(g["moc_offsets"]).tolist()
    np.testing.assert_array_equal(np.array([int(c) for img in cats for c in img]), g["moc_cats"])
    np.testing.assert_array_equal(np.array([float(c) for img in conf for c in img], dtype=np.float32), g["moc_conf"])
    np.testing.assert_array_equal(torch.cat(box).cpu().numpy(), g["moc_box"])


def test_match_object_categories_batch_size_mismatch_returns_none():
    from scene_graph_commonsense_b200 import frontend
    assert frontend.match_object_categories([torch.zeros(3)], [torch.zeros(3)], [torch.zeros(3, 4)], []) == (None, None, None)


@pytest.mark.parametrize("name", sorted(FRONTEND_CASES))
def test_match_target_sgd_matches_reference_golden(name):
    from scene_graph_commonsense_b200 import frontend, ops, pipeline, targets
    g = helpers.golden(name)
    samples = _samples(FRONTEND_CASES[name])
    cs, co, bs_, bo_, rel = frontend.match_target_sgd(0, [s.relationships for s in samples], [s.subj_or_obj for s in samples],
                                                      [s.categories for s in samples], [s.bbox for s in samples])
    off = g["tgt_offsets"]
    for i in range(len(samples)):
        a, b = off[i], off[i + 1]
        if a == b:
            assert rel[i] is None and cs[i] is None and bs_[i] is None
            continue
        np.testing.assert_array_equal(rel[i].cpu().numpy(), g["tgt_rel"][a:b])
        np.testing.assert_array_equal(cs[i].cpu().numpy(), g["tgt_cat_sub"][a:b])
        np.testing.assert_array_equal(co[i].cpu().numpy(), g["tgt_cat_obj"][a:b])
        np.testing.assert_array_equal(bs_[i].cpu().numpy(), g["tgt_box_sub"][a:b])
        np.testing.assert_array_equal(bo_[i].cpu().numpy(), g["tgt_box_obj"][a:b])
    # device table == the host packer the SGDET pipeline used so far
    h = targets.flat_targets_sgd(samples)
    rel_tri, dir_tri, tri_off = frontend.pack_relationships([s.relationships for s in samples], [s.subj_or_obj for s in samples], DEV)
    box_off = torch.from_numpy(np.concatenate(([0], np.cumsum([len(s.categories) for s in samples]))).astype(np.int32)).to(DEV)
    gt_off, label, sub, obj = ops.targets_flat(dir_tri, rel_tri, tri_off, box_off)
    n = int(gt_off[-1])
    np.testing.assert_array_equal(gt_off.cpu().numpy(), h["offsets"])
    for got, want in ((label, h["label"]), (sub, h["sub"]), (obj, h["obj"])):
        np.testing.assert_array_equal(got[:n].cpu().numpy(), want)


def _gt_tables(samples):
    cats = torch.cat([s.categories for s in samples]).to(torch.int32).to(DEV)
    boxes = torch.cat([s.bbox for s in samples]).to(torch.int32).to(DEV)
    off = torch.from_numpy(np.concatenate(([0], np.cumsum([len(s.categories) for s in samples]))).astype(np.int32)).to(DEV)
    return cats, boxes, off


def _oracle_samples(samples, props):
    """GT samples + oracle proposals -> SGDET-shaped samples for hiercom_oracle.replay_sgdet."""
    s2s = tables.sub2super_table()
    out = []
    for s, p in zip(samples, props):
        t = synthetic.ImageSample(s.image_id, s.feat, s.depth, s.bbox, s.categories, s.super_categories, s.relationships, s.subj_or_obj)
        t.bbox_pred = torch.from_numpy(p["bbox"])
        t.categories_pred = torch.from_numpy(p["categories"])
        t.cat_conf_pred = torch.from_numpy(p["conf"])
        t.super_categories_pred = [torch.as_tensor([int(v) for v in s2s[int(c)] if v >= 0], dtype=torch.int64) for c in p["categories"]]
        out.append(t)
    return out


def _scores(samples, b, pairs, gain=3.0):
    off = b.box_offsets.cpu().numpy()
    img, sub, obj = (pairs[k].cpu().numpy() for k in ("img", "sub", "obj"))
    rel, sup, conn = [], [], []
    for i, s_, o_ in zip(img, sub, obj):
        r = synthetic.pair_scores(samples[i].image_id, int(s_ - off[i]), int(o_ - off[i]), helpers.SPLITS, gain=gain)
        rel.append(r[0]); sup.append(r[1]); conn.append(r[2])
    rel, sup, conn = torch.stack(rel), torch.stack(sup), torch.cat(conn)
    return rel.to(DEV), sup.to(DEV), conn.to(DEV), torch.log(torch.sigmoid(conn)).to(DEV)


@pytest.mark.parametrize("mode", ["sgdet", "sgcls"])
@pytest.mark.parametrize("name", ["fe_dense", "fe_ragged"])
def test_window_from_detr_outputs_counters_match_oracle(name, mode):
    """DETR outputs -> front-end kernels -> pair enumeration -> candidates/top-K/match on table-driven scores ==
    oracle front-end -> oracle SGDET replay (evaluate.py:382-446 / :614-694) -> oracle Evaluator.  Bit-exact counters.
    The proposal confidences enter the ranking (evaluator.py:247-249), so the golden's reference confidences are fed to
    the oracle and ours to the kernels: a rank flip from the <=1e-6 softmax difference would show up here."""
    from scene_graph_commonsense_b200 import frontend, pipeline
    case = FRONTEND_CASES[name]
    g = helpers.golden(name)
    samples = _samples(case)
    out_dict = {"pred_logits": torch.from_numpy(g["pred_logits"]).to(DEV), "pred_boxes": torch.from_numpy(g["pred_boxes"]).to(DEV)}
    gt_cats, gt_boxes, gt_off = _gt_tables(samples)
    rel_tri, dir_tri, tri_off = frontend.pack_relationships([s.relationships for s in samples], [s.subj_or_obj for s in samples], DEV)
    build = frontend.sgdet_batch if mode == "sgdet" else frontend.sgcls_batch
    b = build(out_dict, None, None, gt_cats, gt_boxes, gt_off, rel_tri, dir_tri, tri_off, synthetic.reference_args(), skip_mode="batch")
    assert b is not None
    al, vi = synthetic.synthetic_cs_keys(11, 0.5, 0.1)
    pipe = pipeline.RelationPipeline(None, DEV, commonsense=True, aligned_keys=al, violated_keys=vi, predcls=False)
    pairs = pipe.enumerate_pairs(b)
    assert pairs["n"] > 0
    # oracle side
    props = FO.detr_proposals(g["pred_logits"], g["pred_boxes"], tables.alp2fre())
    if mode == "sgcls":
        mc, mf, mb = FO.match_object_categories([p["categories"] for p in props], [p["conf"] for p in props], [p["bbox"] for p in props],
                                                [s.bbox.numpy() for s in samples])
        props = [dict(categories=c, conf=f, bbox=bx.astype(np.float32)) for c, f, bx in zip(mc, mf, mb)]
    osamples = _oracle_samples(samples, props)
    rel, sup, conn, logsig = _scores(osamples, b, pairs)
    pipe.evaluate(b, pairs, rel, sup, logsig)
    ev = O.OracleEvaluator(helpers.SPLITS, hierar=True, aligned=set(al.tolist()), violated=set(vi.tolist()),
                           zero_shot=set(tables.zero_shot_keys().tolist()))
    O.replay_sgdet(osamples, synthetic.batch_score_fn(osamples, helpers.SPLITS, gain=3.0), ev, features=False)
    ev.compute(per_class=True, predcls=False)
    got = pipe.counters.cpu().numpy()[:tables.EV_SIZE]
    np.testing.assert_array_equal(got, ev.counters())
    assert got[tables.EV_NGT] > 0


def test_sgcls_window_is_none_when_reference_skips_the_batch():
    from scene_graph_commonsense_b200 import frontend
    g = helpers.golden("fe_sparse")
    samples = _samples(FRONTEND_CASES["fe_sparse"])
    out_dict = {"pred_logits": torch.from_numpy(g["pred_logits"]).to(DEV), "pred_boxes": torch.from_numpy(g["pred_boxes"]).to(DEV)}
    gt_cats, gt_boxes, gt_off = _gt_tables(samples)
    rel_tri, dir_tri, tri_off = frontend.pack_relationships([s.relationships for s in samples], [s.subj_or_obj for s in samples], DEV)
    assert frontend.sgcls_batch(out_dict, None, None, gt_cats, gt_boxes, gt_off, rel_tri, dir_tri, tri_off) is None


def test_nms_kernel_against_oracle_on_random_clusters_and_ties():
    """Many images, heavy overlap, exact duplicate boxes: keep sets and order vs the fp32 NMS restatement."""
    from scene_graph_commonsense_b200 import frontend
    g = torch.Generator().manual_seed(5)
    b, q, c = 8, 100, 150
    logits = torch.randn(b, q, c + 1, generator=g)
    cls = torch.randint(0, 6, (b, q), generator=g)                      # few classes -> long NMS chains
    logits[torch.arange(b)[:, None], torch.arange(q)[None, :], cls] += 8.0
    centers = torch.rand(b, 5, 2, generator=g)
    which = torch.randint(0, 5, (b, q), generator=g)
    boxes = torch.cat((centers[torch.arange(b)[:, None], which] + 0.02 * torch.randn(b, q, 2, generator=g),
                       0.2 + 0.1 * torch.rand(b, q, 2, generator=g)), dim=2).clamp(0, 1)
    boxes[:, 50:] = boxes[:, :50]                                        # exact duplicates
    p = frontend.detr_proposals({"pred_logits": logits.to(DEV), "pred_boxes": boxes.to(DEV)})
    ref = FO.detr_proposals(logits.numpy(), boxes.numpy(), tables.alp2fre())
    np.testing.assert_array_equal(p.offsets_host, np.concatenate(([0], np.cumsum([len(r["categories"]) for r in ref]))))
    np.testing.assert_array_equal(p.cats.cpu().numpy(), np.concatenate([r["categories"] for r in ref]))
    np.testing.assert_array_equal(p.box_f.cpu().numpy(), np.concatenate([r["bbox"] for r in ref]))
    np.testing.assert_allclose(p.conf.cpu().numpy(), np.concatenate([r["conf"] for r in ref]), rtol=0, atol=CONF_ATOL)
