"""GPU parity of the integer stages (pair enumeration, candidates, commonsense filter, top-K, matching, counters)
against the oracle and the committed reference goldens.  Everything here is BIT-EXACT."""
import numpy as np
import pytest
import torch

from oracle import hiercom_oracle as O
from scene_graph_commonsense_b200 import synthetic, tables
from tests import helpers
from tests.golden_cases import PREDCLS_CASES, SGDET_CASES

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _mods():
    from scene_graph_commonsense_b200 import evaluator, ops, pipeline
    return evaluator, ops, pipeline


# ------------------------------------------------------------------------------------------- pair enumeration
def _ref_pairs(samples, groups):
    """evaluate.py:132-183 loop structure -> per image list of (img, sub, obj, ov, gt_directed, rel) in output order."""
    per_img = {i: [] for i in range(len(samples))}
    for grp in groups:
        n_obj = np.array([len(samples[i].bbox) for i in grp])
        for g in range(int(n_obj.max())):
            keep = [i for i, n in zip(grp, n_obj) if n > g]
            for e in range(g):
                ov = [O.masks_overlap(samples[i].bbox[g], samples[i].bbox[e]) for i in keep]
                if not any(ov):                               # evaluate.py:155-156 whole-batch skip
                    continue
                for i, o in zip(keep, ov):
                    rel = int(samples[i].relationships[g - 1][e])
                    d = float(samples[i].subj_or_obj[g - 1][e])
                    per_img[i].append((i, g, e, int(o), rel if d == 1 else -1, rel))
                    per_img[i].append((i, e, g, int(o), rel if d == 0 else -1, rel))
    return per_img


@pytest.mark.parametrize("mode", ["batch", "per_image"])
def test_pairs_enumerate_matches_reference_loops(mode):
    _, ops, pipeline = _mods()
    ns = [2, 16, 7, 3, 12, 9, 1, 20]
    samples = synthetic.make_batch(list(range(100, 100 + len(ns))), ns, with_maps=False, p_rel=0.5)
    gs = 3
    b = pipeline.batch_from_samples(samples, DEV, skip_mode=mode, group_size=gs, with_maps=False)
    pairs = ops.pairs_enumerate(b.boxes, b.box_offsets, b.tri_offsets, b.p_max, b.rel_tri, b.dir_tri, b.group_id, b.n_groups, b.max_tri)
    groups = [list(range(i, min(i + gs, len(ns)))) for i in range(0, len(ns), gs)] if mode == "batch" else [[i] for i in range(len(ns))]
    ref = _ref_pairs(samples, groups)
    off = b.box_offsets.cpu().numpy()
    got_off = pairs["offsets"].cpu().numpy()
    sub, obj, img, ov, gt, rel = (pairs[k].cpu().numpy() for k in ("sub", "obj", "img", "ov", "gt", "rel"))
    assert pairs["n"] == sum(len(v) for v in ref.values()) == got_off[-1]
    for i in range(len(ns)):
        seg = slice(got_off[i], got_off[i + 1])
        got = list(zip(img[seg].tolist(), (sub[seg] - off[i]).tolist(), (obj[seg] - off[i]).tolist(), ov[seg].tolist(), gt[seg].tolist(),
                       rel[seg].tolist()))
        assert got == ref[i], "image %d" % i


# ------------------------------------------------------------------------------------------- drop-in evaluators
def _dropin_evaluators(case):
    evaluator, _, _ = _mods()
    args = synthetic.reference_args(run_mode=case["run_mode"], hierar=case.get("hierar", True))
    ev = evaluator.Evaluator(args=args, num_classes=50, iou_thresh=0.5, top_k=[20, 50, 100])
    al, vi = helpers.cs_key_arrays(case["run_mode"], case.get("cs"))
    if al is not None:
        ev.set_commonsense(al, vi)
    t3 = evaluator.Evaluator_Top3(args=args, num_classes=50, iou_thresh=0.5, top_k=[20, 50, 100]) if case.get("hierar", True) else None
    return ev, t3


@pytest.mark.parametrize("name", sorted(PREDCLS_CASES))
def test_dropin_evaluator_predcls_matches_reference_golden(name):
    case = PREDCLS_CASES[name]
    g = helpers.golden(name)
    samples = synthetic.make_batch(case["ids"], case["n"], with_maps=False, p_rel=0.5)
    ev, t3 = _dropin_evaluators(case)
    m = m3 = None
    for w in case.get("windows", [list(range(len(samples)))]):
        batch = [samples[i] for i in w]
        fn = synthetic.batch_score_fn(batch, helpers.SPLITS, **case["kw"])
        O.replay_predcls(batch, fn, ev, t3, features=False)
        m = ev.compute(per_class=True)
        ev.clear_data()
        if t3 is not None:
            m3 = t3.compute(per_class=True)
            t3.clear_data()
    np.testing.assert_array_equal(ev.counters().cpu().numpy(), g["ev"])
    np.testing.assert_allclose(helpers.flat_metrics(m), g["metrics"], rtol=0, atol=0, equal_nan=True)
    if t3 is not None:
        np.testing.assert_array_equal(t3.counters().cpu().numpy(), g["t3"])
        np.testing.assert_allclose(helpers.flat_metrics(m3), g["metrics3"], rtol=0, atol=0, equal_nan=True)


@pytest.mark.parametrize("name", sorted(SGDET_CASES))
def test_dropin_evaluator_sgdet_matches_reference_golden(name):
    case = SGDET_CASES[name]
    g = helpers.golden(name)
    prel = case.get("p_rel", [0.5] * len(case["ids"]))
    batch = [synthetic.make_sgdet_image(i, a, b, p_rel=pr, with_maps=False)
             for i, a, b, pr in zip(case["ids"], case["n_gt"], case["n_prop"], prel)]
    ev, _ = _dropin_evaluators(dict(case, hierar=True))
    O.replay_sgdet(batch, synthetic.batch_score_fn(batch, helpers.SPLITS, gain=3.0), ev, features=False)
    m = ev.compute(per_class=True, predcls=False)
    np.testing.assert_array_equal(ev.counters().cpu().numpy(), g["ev"])
    np.testing.assert_allclose(helpers.flat_metrics(m), g["metrics"], rtol=0, atol=0, equal_nan=True)


# ------------------------------------------------------------------------------------------- batched pipeline
def _scores_for_pairs(samples, b, pairs, kw, sgdet=False):
    off = b.box_offsets.cpu().numpy()
    img, sub, obj = (pairs[k].cpu().numpy() for k in ("img", "sub", "obj"))
    rel, sup, conn = [], [], []
    for i, s_, o_ in zip(img, sub, obj):
        r = synthetic.pair_scores(samples[i].image_id, int(s_ - off[i]), int(o_ - off[i]), helpers.SPLITS, **kw)
        rel.append(r[0]); sup.append(r[1]); conn.append(r[2])
    rel, sup, conn = torch.stack(rel), torch.stack(sup), torch.cat(conn)
    logsig = torch.log(torch.sigmoid(conn))                       # train_utils.py:190, computed exactly as the reference does
    return rel.to(DEV), sup.to(DEV), conn.to(DEV), logsig.to(DEV)


def _stable_topk(conf, offsets, k_per_row, top=100):
    out = []
    for i in range(len(offsets) - 1):
        seg = conf[offsets[i] * k_per_row:offsets[i + 1] * k_per_row]
        order = np.argsort(-seg.astype(np.float64), kind="stable")[:top]
        out.append(order)
    return out


@pytest.mark.parametrize("name", sorted(PREDCLS_CASES))
def test_pipeline_integer_stages_match_reference_golden(name):
    _, ops, pipeline = _mods()
    case = PREDCLS_CASES[name]
    g = helpers.golden(name)
    samples = synthetic.make_batch(case["ids"], case["n"], with_maps=False, p_rel=0.5)
    windows = case.get("windows", [list(range(len(samples)))])
    al, vi = helpers.cs_key_arrays(case["run_mode"], case.get("cs"))
    pipe = pipeline.RelationPipeline(None, DEV, commonsense=al is not None, aligned_keys=al, violated_keys=vi, hier=case["hierar"])
    for w in windows:
        batch = [samples[i] for i in w]
        b = pipeline.batch_from_samples(batch, DEV, skip_mode="batch", with_maps=False)
        pairs = pipe.enumerate_pairs(b)
        rel, sup, conn, logsig = _scores_for_pairs(batch, b, pairs, case["kw"])
        res = pipe.evaluate(b, pairs, rel, sup if case["hierar"] else None, logsig, connectivity=conn, want_topk=True)
        # top-K sets: (confidence desc, candidate index asc)
        k = 3 if case["hierar"] else 1
        ref_top = _stable_topk(res["cand_conf"].cpu().numpy(), pairs["offsets"].cpu().numpy(), k)
        got_top = res["topk"].cpu().numpy()
        for i, r in enumerate(ref_top):
            np.testing.assert_array_equal(got_top[i][:len(r)], r)
            assert (got_top[i][len(r):] == -1).all()
    c = pipe.counters.cpu().numpy()
    np.testing.assert_array_equal(c[:tables.EV_SIZE], g["ev"])
    if case["hierar"]:
        np.testing.assert_array_equal(c[tables.EV_SIZE:], g["t3"])
    np.testing.assert_array_equal(pipe.stats.cpu().numpy().astype(np.float64), g["stats"])
    m = pipe.metrics()
    flat = helpers.flat_metrics(m["evaluator"])
    np.testing.assert_allclose(flat, g["metrics"], rtol=0, atol=0, equal_nan=True)


@pytest.mark.parametrize("name", sorted(SGDET_CASES))
def test_pipeline_sgdet_matches_reference_golden(name):
    _, ops, pipeline = _mods()
    case = SGDET_CASES[name]
    g = helpers.golden(name)
    prel = case.get("p_rel", [0.5] * len(case["ids"]))
    batch = [synthetic.make_sgdet_image(i, a, b, p_rel=pr, with_maps=False)
             for i, a, b, pr in zip(case["ids"], case["n_gt"], case["n_prop"], prel)]
    al, vi = helpers.cs_key_arrays(case["run_mode"], case.get("cs"))
    pipe = pipeline.RelationPipeline(None, DEV, commonsense=al is not None, aligned_keys=al, violated_keys=vi, predcls=False)
    b = pipeline.batch_from_samples(batch, DEV, skip_mode="batch", sgdet=True, with_maps=False)
    pairs = pipe.enumerate_pairs(b)
    rel, sup, conn, logsig = _scores_for_pairs(batch, b, pairs, dict(gain=3.0))
    pipe.evaluate(b, pairs, rel, sup, logsig)
    np.testing.assert_array_equal(pipe.counters.cpu().numpy()[:tables.EV_SIZE], g["ev"])


def test_per_image_mode_equals_reference_at_batch_size_one():
    """skip_mode='per_image' == the reference loop run with one image per batch."""
    _, ops, pipeline = _mods()
    case = dict(ids=[200, 201, 202, 203, 204], n=[9, 14, 4, 11, 17], run_mode="eval_cs", hierar=True, kw=dict(gain=3.0), cs=(9, 0.5, 0.1),
                windows=[[0], [1], [2], [3], [4]])
    ev, t3, _, _, stats, samples = helpers.replay_predcls_case(case)
    al, vi = helpers.cs_key_arrays(case["run_mode"], case["cs"])
    pipe = pipeline.RelationPipeline(None, DEV, commonsense=True, aligned_keys=al, violated_keys=vi)
    b = pipeline.batch_from_samples(samples, DEV, skip_mode="per_image", with_maps=False)
    pairs = pipe.enumerate_pairs(b)
    rel, sup, conn, logsig = _scores_for_pairs(samples, b, pairs, case["kw"])
    pipe.evaluate(b, pairs, rel, sup, logsig, connectivity=conn)
    c = pipe.counters.cpu().numpy()
    np.testing.assert_array_equal(c[:tables.EV_SIZE], ev.counters())
    np.testing.assert_array_equal(c[tables.EV_SIZE:], t3.counters())
    assert c[tables.EV_NGT] > 0 and c[tables.EV_HITS + 2] > 0


def test_topk_edge_cases_ties_all_minus_inf_and_short_lists():
    """All candidates -inf (every key ties), fewer than 100 candidates, duplicate boxes."""
    _, ops, pipeline = _mods()
    samples = synthetic.make_batch([300, 301], [3, 25], with_maps=False, p_rel=0.8)
    samples[1].bbox[:] = samples[1].bbox[0]                        # identical boxes: every pair overlaps, IoU == 1
    pipe = pipeline.RelationPipeline(None, DEV, commonsense=True, aligned_keys=np.zeros(0, np.int64), violated_keys=np.zeros(0, np.int64))
    b = pipeline.batch_from_samples(samples, DEV, skip_mode="per_image", with_maps=False)
    pairs = pipe.enumerate_pairs(b)
    rel, sup, conn, logsig = _scores_for_pairs(samples, b, pairs, dict(gain=3.0))
    res = pipe.evaluate(b, pairs, rel, sup, logsig, want_topk=True)
    assert torch.isinf(res["cand_conf"]).all()                     # empty aligned set: filter kills everything
    got = res["topk"].cpu().numpy()
    off = pairs["offsets"].cpu().numpy() * 3
    for i in range(2):
        n = min(100, off[i + 1] - off[i])
        np.testing.assert_array_equal(got[i][:n], np.arange(n))     # ties resolve in candidate-index order
    # oracle agrees on the counters (-inf candidates still match)
    ev = O.OracleEvaluator(helpers.SPLITS, True, aligned=set(), violated=set(), zero_shot=set(tables.zero_shot_keys().tolist()))
    t3 = O.OracleEvaluatorTop3(helpers.SPLITS)
    for s in samples:
        O.replay_predcls([s], synthetic.batch_score_fn([s], helpers.SPLITS, gain=3.0), ev, t3, features=False)
        ev.compute(per_class=True); ev.clear_data(); t3.compute(per_class=True); t3.clear_data()
    c = pipe.counters.cpu().numpy()
    np.testing.assert_array_equal(c[:tables.EV_SIZE], ev.counters())
    np.testing.assert_array_equal(c[tables.EV_SIZE:], t3.counters())


def test_hier_head_kernel_matches_torch_fp32():
    _, ops, _ = _mods()
    g = torch.Generator().manual_seed(3)
    n = 517
    raw = torch.randn(n, 512, generator=g)
    sd = synthetic.head_state_dict(seed=4, logit_gain=30.0)
    emb = sd["fc2.weight"][:, 4096:].t().contiguous()
    cats = torch.randint(0, 150, (40,), generator=g)
    s2s = tables.sub2super_table()
    supers = torch.as_tensor(s2s[cats.numpy()])
    row_sub = torch.randint(0, 40, (n,), generator=g)
    row_obj = torch.randint(0, 40, (n,), generator=g)
    w = torch.cat([sd["fc3_1.weight"], sd["fc3_2.weight"], sd["fc3_3.weight"], sd["fc4.weight"], sd["fc5.weight"]])
    bh = torch.cat([sd["fc3_1.bias"], sd["fc3_2.bias"], sd["fc3_3.bias"], sd["fc4.bias"], sd["fc5.bias"]])
    c = lambda t, dt=None: (t.to(dt) if dt else t).to(DEV).contiguous()
    rel, sup, conn, logsig, pred = ops.hier_head(c(raw), c(sd["fc2.bias"]), c(emb), c(row_sub, torch.int32), c(row_obj, torch.int32),
                                                 c(cats, torch.int32), c(supers, torch.int8), c(w), c(bh), (15, 11, 24), want_pred=True)
    x = raw + sd["fc2.bias"] + emb[cats[row_sub]] + emb[150 + cats[row_obj]]
    # utils.py:136-149: the first entry of a box's super-class list plus the LAST entry of a 2..4-entry list (never the middle ones)
    n_sup = (supers >= 0).sum(1)
    assert int((n_sup == 3).sum()) > 0
    for k in range(4):
        for role, rows, base in ((0, row_sub, 300), (1, row_obj, 317)):
            sv = supers[rows][:, k].long()
            used = (sv >= 0) & ((k == 0) | (n_sup[rows] - 1 == k))
            x = x + torch.where(used[:, None], emb[(base + sv.clamp(min=0))], torch.zeros(1, 512))
    p = torch.relu(x)
    r1, r2, r3, s_ref, c_ref = O.hier_head(sd, p)
    np.testing.assert_allclose(pred.cpu().numpy(), p.numpy(), atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(rel.cpu().numpy(), torch.cat((r1, r2, r3), 1).numpy(), atol=2e-4, rtol=0)
    np.testing.assert_allclose(sup.cpu().numpy(), s_ref.numpy(), atol=2e-4, rtol=0)
    np.testing.assert_allclose(conn.cpu().numpy(), c_ref[:, 0].numpy(), atol=2e-4, rtol=0)
    np.testing.assert_allclose(logsig.cpu().numpy(), torch.log(torch.sigmoid(c_ref[:, 0])).numpy(), atol=2e-4, rtol=0)


def test_bayesian_head_module_matches_reference_golden():
    from scene_graph_commonsense_b200 import model
    g = helpers.golden("head")
    bh = model.BayesianHead(input_dim=512).to(DEV)
    sd = {k: v for k, v in synthetic.head_state_dict(seed=2, logit_gain=20.0).items() if k.startswith(("fc3_", "fc5"))}
    bh.load_state_dict(sd)
    h = torch.randn(16, 512, generator=torch.Generator().manual_seed(5)).to(DEV)
    b1, b2, b3, bs = bh(h)
    np.testing.assert_allclose(torch.cat((b1, b2, b3), 1).cpu().numpy(), g["bhead_relation"], atol=2e-4, rtol=0)
    np.testing.assert_allclose(bs.cpu().numpy(), g["bhead_super"], atol=2e-4, rtol=0)


@pytest.mark.parametrize("mode,gs", [("per_image", None), ("batch", None), ("batch", 3)])
def test_host_counted_pair_offsets_equal_the_device_enumeration(mode, gs):
    """The offsets `host_batch_from_samples` counts on the host (so that a step needs no device -> host read) are exactly what
    hc_pairs_enumerate produces, and the pair lists do not depend on which of the two sized them."""
    _, ops, pipeline = _mods()
    samples = synthetic.make_batch([31, 32, 33, 34, 35, 36, 37], [9, 1, 14, 2, 7, 40, 11], with_maps=False)
    samples[0].bbox[:5] = torch.tensor([(0, 32, 0, 32), (5, 5, 3, 9), (9, 3, 2, 7), (-4, 32, 3, 12), (30, 40, -2, 2)], dtype=samples[0].bbox.dtype)
    b = pipeline.batch_from_samples(samples, DEV, skip_mode=mode, group_size=gs, with_maps=False)
    pipe = pipeline.RelationPipeline(None, DEV, commonsense=False)
    assert pipe.host_offsets and b.pair_offsets_host is not None
    fast = pipe.enumerate_pairs(b)
    pipe.host_offsets = False
    slow = pipe.enumerate_pairs(b)
    np.testing.assert_array_equal(slow["offsets_host"], b.pair_offsets_host)
    np.testing.assert_array_equal(fast["offsets"].cpu().numpy(), b.pair_offsets_host)
    assert fast["n"] == slow["n"] > 0
    for k in ("sub", "obj", "img", "ov", "gt", "rel"):
        assert torch.equal(fast[k], slow[k]), k
