"""GPU parity of the full relation head (bf16 tcgen05 path) against the fp32 reference goldens and the oracle.
Tolerance (BASELINE.json north_star): joint relation probabilities within 2e-3 absolute of the fp32 reference."""
import numpy as np
import pytest
import torch

from oracle import hiercom_oracle as O
from scene_graph_commonsense_b200 import synthetic, tables
from tests import helpers

pytestmark = pytest.mark.gpu
DEV = "cuda"
PROB_TOL = 2e-3
PAIRS = [(1, 0), (0, 1), (3, 2), (2, 3), (4, 1), (1, 4)]


def _inputs():
    s = synthetic.make_image(900, 5)
    hs = torch.stack([O._masked_input(s, s.bbox[a]) for a, b in PAIRS])
    ho = torch.stack([O._masked_input(s, s.bbox[b]) for a, b in PAIRS])
    c1 = torch.stack([s.categories[a] for a, b in PAIRS])
    c2 = torch.stack([s.categories[b] for a, b in PAIRS])
    s1 = [s.super_categories[a] for a, b in PAIRS]
    s2 = [s.super_categories[b] for a, b in PAIRS]
    return s, hs, ho, c1, c2, s1, s2


def _prob_err(log_got, log_ref):
    return float(np.abs(np.exp(log_got.astype(np.float64)) - np.exp(log_ref.astype(np.float64))).max())


@pytest.mark.parametrize("tag,gain", [("init", 1.0), ("trained", 40.0)])
def test_legacy_forward_matches_reference_golden(tag, gain):
    from scene_graph_commonsense_b200 import model
    g = helpers.golden("head")
    s, hs, ho, c1, c2, s1, s2 = _inputs()
    net = model.BayesianRelationClassifier(synthetic.reference_args()).to(DEV)
    net.load_state_dict({"module." + k: v for k, v in synthetic.head_state_dict(seed=0, logit_gain=gain).items()})   # DDP-prefixed keys
    r1, r2, r3, sup, conn, pred, pred_aug = net(hs.to(DEV), ho.to(DEV), c1.to(DEV), c2.to(DEV), s1, s2, 0)
    assert r1.shape == (6, 15) and r2.shape == (6, 11) and r3.shape == (6, 24) and sup.shape == (6, 3) and conn.shape == (6, 1)
    assert pred.shape == (6, 512) and pred_aug is None
    rel = torch.cat((r1, r2, r3), 1).cpu().numpy()
    assert _prob_err(rel, g["hier_%s_relation" % tag]) <= PROB_TOL
    assert _prob_err(sup.cpu().numpy(), g["hier_%s_super" % tag]) <= PROB_TOL
    sig = lambda x: 1.0 / (1.0 + np.exp(-x.astype(np.float64)))
    assert np.abs(sig(conn.cpu().numpy()) - sig(g["hier_%s_conn" % tag])).max() <= PROB_TOL
    scale = np.abs(g["hier_%s_pred" % tag]).max()
    assert np.abs(pred.cpu().numpy() - g["hier_%s_pred" % tag]).max() <= 0.02 * scale


def test_flat_classifier_matches_reference_golden():
    from scene_graph_commonsense_b200 import model
    g = helpers.golden("head")
    s, hs, ho, c1, c2, s1, s2 = _inputs()
    net = model.FlatRelationClassifier(synthetic.reference_args(hierar=False)).to(DEV)
    net.load_state_dict(synthetic.head_state_dict(seed=1, flat=True))
    rel, conn, pred, _ = net(hs.to(DEV), ho.to(DEV), c1.to(DEV), c2.to(DEV), s1, s2, 0)
    assert np.abs(rel.cpu().numpy() - g["flat_relation"]).max() <= 2e-3
    assert np.abs(conn.cpu().numpy() - g["flat_conn"]).max() <= 2e-3


def test_batched_pipeline_matches_legacy_and_oracle():
    """Factored path (conv1 per image, conv2 halves per box) == unfactored reference formulation."""
    from scene_graph_commonsense_b200 import model, pipeline
    g = helpers.golden("head")
    s = synthetic.make_image(900, 5)
    sd = synthetic.head_state_dict(seed=0, logit_gain=40.0)
    pk = model.PackedHead(sd, DEV)
    pipe = pipeline.RelationPipeline(pk, DEV, commonsense=False)
    b = pipeline.batch_from_samples([s], DEV, skip_mode="batch")
    dev_pairs = dict(n=len(PAIRS), sub=torch.tensor([a for a, _ in PAIRS], dtype=torch.int32, device=DEV),
                     obj=torch.tensor([o for _, o in PAIRS], dtype=torch.int32, device=DEV))
    rel, sup, conn, logsig = pipe.forward_pairs(b, dev_pairs)
    assert _prob_err(rel.cpu().numpy(), g["hier_trained_relation"]) <= PROB_TOL
    assert _prob_err(sup.cpu().numpy(), g["hier_trained_super"]) <= PROB_TOL
    ref_ls = np.log(1.0 / (1.0 + np.exp(-g["hier_trained_conn"][:, 0].astype(np.float64))))
    assert np.abs(np.exp(logsig.cpu().numpy()) - np.exp(ref_ls)).max() <= PROB_TOL


def test_forward_pairs_with_the_survey_argument_list_equals_the_batch_form():
    """`forward_pairs(feat, depth, boxes, cats, supercats, pair_index, box_img)` == the DeviceBatch form, two images."""
    from scene_graph_commonsense_b200 import model, pipeline
    samples = [synthetic.make_image(903, 5), synthetic.make_image(904, 4)]
    pk = model.PackedHead(synthetic.preset_state_dict("trained"), DEV)
    pipe = pipeline.RelationPipeline(pk, DEV, commonsense=False)
    b = pipeline.batch_from_samples(samples, DEV, skip_mode="per_image")
    pairs = pipe.enumerate_pairs(b)
    want = pipe.forward_pairs(b, pairs)
    feat = torch.stack([s.feat for s in samples]).to(DEV)
    depth = torch.stack([s.depth for s in samples]).to(DEV)
    boxes = torch.cat([s.bbox for s in samples]).float()                   # floats are truncated like the reference's int()
    cats = torch.cat([s.categories for s in samples])
    supercats = [sc for s in samples for sc in s.super_categories]
    pair_index = torch.stack((pairs["sub"], pairs["obj"]), dim=1).long()
    box_img = torch.tensor([0] * 5 + [1] * 4)
    got = pipe.forward_pairs(feat, depth, boxes, cats, supercats, pair_index, box_img)
    for g, w in zip(got, want):
        assert float((g - w).abs().max()) <= 1e-5         # same kernels; the generic pair-gather pooling vs the tiled one is bit-identical
    empty = pipe.forward_pairs(feat, depth, boxes, cats, supercats, pair_index[:0], box_img)
    assert empty[0].shape == (0, 50) and empty[1].shape == (0, 3)


def test_end_to_end_step_counters_match_oracle_small():
    """cfg1-shaped: one image, 8 boxes, full model: counters of the CUDA path == oracle replay with the CUDA scores'
    own candidates is covered elsewhere; here the whole step runs and its scores stay within tolerance of the oracle
    for every directed pair (the oracle finishes in a few seconds at this size)."""
    from scene_graph_commonsense_b200 import model, pipeline
    s = synthetic.make_image(901, 8, p_rel=0.5)
    sd = synthetic.head_state_dict(seed=0, logit_gain=40.0)
    pk = model.PackedHead(sd, DEV)
    pipe = pipeline.RelationPipeline(pk, DEV, commonsense=True)
    b = pipeline.batch_from_samples([s], DEV, skip_mode="batch")
    pairs = pipe.enumerate_pairs(b)
    rel, sup, conn, logsig = pipe.forward_pairs(b, pairs)
    captured = {}

    def head_fn(h_sub, h_obj, c1, c2, s1, s2, ctx):
        out = O.make_head_fn(sd)(h_sub, h_obj, c1, c2, s1, s2)
        captured[(ctx[1], ctx[2])] = out
        return out
    ev = O.OracleEvaluator(helpers.SPLITS, True, aligned=set(tables.commonsense_aligned_keys().tolist()),
                           violated=set(tables.commonsense_violated_keys().tolist()), zero_shot=set(tables.zero_shot_keys().tolist()))
    n = O.replay_predcls([s], head_fn, ev)
    assert n == pairs["n"]
    sub, obj = pairs["sub"].cpu().numpy(), pairs["obj"].cpu().numpy()
    rel_c = rel.cpu().numpy()
    worst = 0.0
    for p in range(pairs["n"]):
        r_ref = captured[(int(sub[p]), int(obj[p]))][0][0].numpy()
        worst = max(worst, _prob_err(rel_c[p], r_ref))
    assert worst <= PROB_TOL, worst
    pipe.evaluate(b, pairs, rel, sup, logsig, connectivity=conn)
    ev.compute(per_class=True)
    assert int(pipe.counters[tables.EV_NGT]) == int(ev.num_connected_target) > 0
