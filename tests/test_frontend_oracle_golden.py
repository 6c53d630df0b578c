"""CPU: the front-end oracle (oracle/frontend_oracle.py) against goldens produced by executing the reference's own
evaluate.py:311-370 source lines + utils.match_object_categories + utils.match_target_sgd (oracle/make_golden_frontend.py)."""
import numpy as np
import pytest

from oracle import frontend_oracle as FO
from oracle import hiercom_oracle as O
from scene_graph_commonsense_b200 import synthetic, tables
from tests.golden_cases import FRONTEND_CASES
from tests.helpers import golden


def split(flat, off):
    return [flat[off[i]:off[i + 1]] for i in range(len(off) - 1)]


@pytest.mark.parametrize("name", sorted(FRONTEND_CASES))
def test_detr_proposals_match_reference(name):
    g = golden(name)
    props = FO.detr_proposals(g["pred_logits"], g["pred_boxes"], tables.alp2fre())
    off = g["offsets"]
    assert [len(p["categories"]) for p in props] == np.diff(off).tolist()
    cats = np.concatenate([p["categories"] for p in props])
    assert np.array_equal(cats, g["cats"])                                      # labels + NMS keep set + order: bit-exact
    assert np.array_equal(np.concatenate([p["bbox"] for p in props]), g["bbox"])   # fp32 box arithmetic: bit-exact
    np.testing.assert_allclose(np.concatenate([p["conf"] for p in props]), g["conf"], rtol=0, atol=1e-7)
    assert np.array_equal(tables.sub2super_table()[cats], g["supers"])          # evaluate.py:368-370


@pytest.mark.parametrize("name", sorted(FRONTEND_CASES))
def test_match_object_categories_matches_reference(name):
    g = golden(name)
    case = FRONTEND_CASES[name]
    samples = synthetic.make_batch(case["ids"], case["n_gt"], with_maps=False, p_rel=0.5)
    off = g["offsets"]
    cats, conf, box = FO.match_object_categories(split(g["cats"], off), split(g["conf"], off), split(g["bbox"], off),
                                                 [s.bbox.numpy() for s in samples])
    if bool(g["moc_none"]):
        assert cats is None and conf is None and box is None
        return
    assert [len(c) for c in cats] == np.diff(g["moc_offsets"]).tolist()
    assert np.array_equal(np.concatenate(cats), g["moc_cats"])
    assert np.array_equal(np.concatenate(conf), g["moc_conf"])                  # conf * iou in fp32: bit-exact
    assert np.array_equal(np.concatenate(box), g["moc_box"])


@pytest.mark.parametrize("name", sorted(FRONTEND_CASES))
def test_match_target_sgd_matches_reference(name):
    g = golden(name)
    case = FRONTEND_CASES[name]
    samples = synthetic.make_batch(case["ids"], case["n_gt"], with_maps=False, p_rel=0.5)
    rel, cs, co, bs_, bo_ = O.match_target_sgd(samples)
    off = g["tgt_offsets"]
    for i in range(len(samples)):
        a, b = off[i], off[i + 1]
        if a == b:
            assert rel[i] is None
            continue
        assert np.array_equal(rel[i], g["tgt_rel"][a:b]) and np.array_equal(cs[i], g["tgt_cat_sub"][a:b])
        assert np.array_equal(co[i], g["tgt_cat_obj"][a:b])
        assert np.array_equal(bs_[i], g["tgt_box_sub"][a:b]) and np.array_equal(bo_[i], g["tgt_box_obj"][a:b])


def test_nms_restatement_against_torchvision_random():
    tv = pytest.importorskip("torchvision")
    import torch
    rng = np.random.default_rng(0)
    for trial in range(30):
        n = int(rng.integers(1, 40))
        xy = rng.random((n, 2)).astype(np.float32) * 24
        wh = rng.random((n, 2)).astype(np.float32) * 10
        b = np.concatenate((xy, xy + wh), axis=1).astype(np.float32)
        if trial % 3 == 0 and n > 3:
            b[n // 2:] = b[:n - n // 2]                                         # exact duplicates / ties
        s = rng.random(n).astype(np.float32)
        if trial % 5 == 0:
            s = np.round(s, 1)                                                  # score ties -> stable order
        ref = tv.ops.nms(torch.from_numpy(b), torch.from_numpy(s), 0.5).numpy()
        assert np.array_equal(FO.nms_xyxy(b, s, 0.5), ref)
