"""BASELINE-size checks through size-independent properties (the oracle's Python loops would take hours here):
cfg2 (64 x 40 boxes, 99 840 pairs) and cfg3 (100 proposals / image, 29 700 candidates / image)."""
import numpy as np
import pytest
import torch

from scene_graph_commonsense_b200 import synthetic, tables

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _random_scores(n, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    logits = torch.randn(n, 54, generator=g, device=DEV) * 3
    sup = torch.log_softmax(logits[:, 50:53], 1)
    rel = torch.cat((torch.log_softmax(logits[:, :15], 1) + sup[:, 0:1], torch.log_softmax(logits[:, 15:26], 1) + sup[:, 1:2],
                     torch.log_softmax(logits[:, 26:50], 1) + sup[:, 2:3]), 1).contiguous()
    conn = logits[:, 53].contiguous()
    return rel, sup.contiguous(), conn, torch.log(torch.sigmoid(conn))


@pytest.mark.parametrize("n_img,n_box,mode", [(64, 40, "batch"), (6, 100, "per_image")])
def test_integer_stages_properties_at_baseline_sizes(n_img, n_box, mode):
    from scene_graph_commonsense_b200 import pipeline
    samples = synthetic.make_batch(list(range(700, 700 + n_img)), n_box, with_maps=False, p_rel=0.3)
    al, vi = synthetic.synthetic_cs_keys(1, 0.5, 0.1)
    mk = lambda: pipeline.RelationPipeline(None, DEV, commonsense=True, aligned_keys=al, violated_keys=vi)
    pipe = mk()
    b = pipeline.batch_from_samples(samples, DEV, skip_mode=mode, with_maps=False)
    pairs = pipe.enumerate_pairs(b)
    off = pairs["offsets"].cpu().numpy()
    if mode == "batch":
        assert pairs["n"] == n_img * n_box * (n_box - 1)              # 64 images in lock-step: every (g,e) survives
    # pair indexing: directed pairs come in mirrored couples, subjects/objects inside their image, GT only on one direction
    sub, obj, img, gt, ov = (pairs[k].cpu().numpy() for k in ("sub", "obj", "img", "gt", "ov"))
    assert (sub[0::2] == obj[1::2]).all() and (obj[0::2] == sub[1::2]).all() and (ov[0::2] == ov[1::2]).all()
    assert ((gt[0::2] == -1) | (gt[1::2] == -1)).all()
    box_off = b.box_offsets.cpu().numpy()
    assert (sub >= box_off[img]).all() and (sub < box_off[img + 1]).all() and (np.diff(img) >= 0).all()
    if mode == "per_image":
        assert ov.all()                                               # per-image rule keeps exactly the overlapping pairs
    rel, sup, conn, logsig = _random_scores(pairs["n"], 5)
    res = pipe.evaluate(b, pairs, rel, sup, logsig, connectivity=conn, want_topk=True)
    conf = res["cand_conf"].cpu().numpy()
    top = res["topk"].cpu().numpy()
    for i in range(n_img):
        seg = conf[3 * off[i]:3 * off[i + 1]]
        order = np.argsort(-seg.astype(np.float64), kind="stable")[:100]
        np.testing.assert_array_equal(top[i][:len(order)], order)     # sortedness + tie order on up to 29 700 candidates
    c1 = pipe.counters.cpu().numpy().copy()
    n_gt = int((gt != -1).sum())
    assert c1[tables.EV_NGT] == n_gt == c1[tables.EV_SIZE + tables.T3_NGT]
    assert c1[tables.EV_HITS] <= c1[tables.EV_HITS + 1] <= c1[tables.EV_HITS + 2] <= n_gt
    assert c1[tables.EV_HITS_PC:tables.EV_HITS_PC + 50].sum() == c1[tables.EV_HITS] and c1[tables.EV_NGT_PC:tables.EV_NGT_PC + 50].sum() == n_gt
    assert (c1[tables.EV_BLOCK:tables.EV_SIZE] <= c1[:tables.EV_BLOCK]).all()       # zero-shot twin is a subset
    # idempotence / linearity: a second pass adds exactly the same integers
    pipe.evaluate(b, pairs, rel, sup, logsig, connectivity=conn)
    np.testing.assert_array_equal(pipe.counters.cpu().numpy(), 2 * c1)
    if mode == "per_image":
        # shard invariance: images dealt round-robin to 4 "ranks", counters summed == single pass
        total = np.zeros_like(c1)
        for r in range(4):
            ids = list(range(r, n_img, 4))
            ps = mk()
            bs = pipeline.batch_from_samples([samples[i] for i in ids], DEV, skip_mode="per_image", with_maps=False)
            prs = ps.enumerate_pairs(bs)
            rows = torch.cat([torch.arange(off[i], off[i + 1], device=DEV) for i in ids])
            ps.evaluate(bs, prs, rel[rows].contiguous(), sup[rows].contiguous(), logsig[rows].contiguous(), connectivity=conn[rows].contiguous())
            total += ps.counters.cpu().numpy()
        np.testing.assert_array_equal(total, c1)


def test_cfg1_full_image_scores_within_tolerance_of_oracle():
    """BASELINE config 1 in full: one image, 20 GT boxes, every processed directed pair through the bf16 tcgen05 head vs the
    fp32 oracle (about 10 s of CPU); joint probabilities within 2e-3, and the evaluator counters agree when both sides
    consume the CUDA scores."""
    from oracle import hiercom_oracle as O
    from scene_graph_commonsense_b200 import model, pipeline
    from tests import helpers
    s = synthetic.make_image(990, 20, p_rel=0.3)
    sd = synthetic.head_state_dict(seed=0, logit_gain=40.0)
    pipe = pipeline.RelationPipeline(model.PackedHead(sd, DEV), DEV, commonsense=True)
    b = pipeline.batch_from_samples([s], DEV, skip_mode="batch")
    pairs = pipe.enumerate_pairs(b)
    rel, sup, conn, logsig = pipe.forward_pairs(b, pairs)
    pipe.evaluate(b, pairs, rel, sup, logsig, connectivity=conn)
    sub, obj = pairs["sub"].cpu().numpy(), pairs["obj"].cpu().numpy()
    lut = {(int(a), int(c)): i for i, (a, c) in enumerate(zip(sub, obj))}
    relc, supc, connc = rel.cpu(), sup.cpu(), conn.cpu()
    worst = [0.0]
    head = O.make_head_fn(sd)

    def head_fn(h_sub, h_obj, c1, c2, s1, s2, ctx):
        r, sp, cn = head(h_sub, h_obj, c1, c2, s1, s2)
        i = lut[(ctx[1], ctx[2])]
        worst[0] = max(worst[0], float((torch.exp(r[0].double()) - torch.exp(relc[i].double())).abs().max()))
        return relc[i:i + 1], supc[i:i + 1], connc[i:i + 1].view(1, 1)          # evaluator consumes the CUDA scores
    ev, t3 = helpers.oracle_evaluators(dict(run_mode="eval_cs"), True)
    n = O.replay_predcls([s], head_fn, ev, t3)
    ev.compute(per_class=True); t3.compute(per_class=True)
    assert n == pairs["n"] and worst[0] <= 2e-3, worst[0]
    c = pipe.counters.cpu().numpy()
    # logsig is recomputed by the oracle replay with torch.log(torch.sigmoid(.)) on the same fp32 connectivity
    np.testing.assert_array_equal(c[:tables.EV_SIZE], ev.counters())
    np.testing.assert_array_equal(c[tables.EV_SIZE:], t3.counters())


def test_streamed_windows_equal_sequential_steps():
    """RelationPipeline.run (H2D of window k+1 prefetched on a copy stream under window k's kernels) returns, window by
    window, exactly the counters of stepping the same windows one after the other."""
    from scene_graph_commonsense_b200 import model, pipeline
    sd = synthetic.head_state_dict(seed=0, logit_gain=40.0)
    pk = model.PackedHead(sd, DEV)
    hosts = [pipeline.host_batch_from_samples(synthetic.make_batch([880 + 3 * w, 881 + 3 * w, 882 + 3 * w][:2 + w % 2], 4 + w, p_rel=0.5))
             for w in range(4)]
    seq = pipeline.RelationPipeline(pk, DEV, commonsense=True)
    want = []
    for hb in hosts:
        n = seq.step(hb.to_device(DEV))
        want.append((n, seq.counters.cpu().numpy().copy()))
    stream = pipeline.RelationPipeline(pk, DEV, commonsense=True)
    got = [(n, c.numpy().copy()) for n, c in stream.run(iter(hosts))]
    assert len(got) == len(want)
    for (n0, c0), (n1, c1) in zip(want, got):
        assert n0 == n1
        np.testing.assert_array_equal(c0, c1)
    assert want[-1][1][tables.EV_NGT] > 0
    # with a reset hook every window stands alone
    solo = pipeline.RelationPipeline(pk, DEV, commonsense=True)
    per_window = [c.numpy().copy() for _, c in solo.run(iter(hosts), before_step=lambda p: p.reset())]
    np.testing.assert_array_equal(np.sum(per_window, axis=0), want[-1][1])


def test_cfg2_full_size_counters_bit_exact_vs_oracle():
    """BASELINE config 2 in full (64 images x 40 boxes, 99 840 directed pairs, 299 520 candidates, shipped commonsense sets,
    reference batch skip rule): every Evaluator / Evaluator_Top3 counter and the connectivity statistics equal the loop
    oracle's on identical scores (about 15 s of CPU for the oracle)."""
    from oracle import hiercom_oracle as O
    from scene_graph_commonsense_b200 import pipeline
    from tests import helpers
    from tests.test_gpu_eval import _scores_for_pairs
    samples = synthetic.make_batch(list(range(64)), 40, base_seed=0, with_maps=False, p_rel=0.3)
    pipe = pipeline.RelationPipeline(None, DEV, commonsense=True)
    b = pipeline.batch_from_samples(samples, DEV, skip_mode="batch", with_maps=False)
    pairs = pipe.enumerate_pairs(b)
    assert pairs["n"] == 99840
    rel, sup, conn, logsig = _scores_for_pairs(samples, b, pairs, dict(gain=3.0))
    pipe.evaluate(b, pairs, rel, sup, logsig, connectivity=conn)
    ev, t3 = helpers.oracle_evaluators(dict(run_mode="eval_cs"), True)
    stats = dict(num_not_connected=0, num_connected=0, num_connected_pred=0, connectivity_precision=0, connectivity_recall=0)
    n = O.replay_predcls(samples, synthetic.batch_score_fn(samples, helpers.SPLITS, gain=3.0), ev, t3, stats=stats, features=False)
    m = ev.compute(per_class=True)
    m3 = t3.compute(per_class=True)
    assert n == pairs["n"]
    c = pipe.counters.cpu().numpy()
    np.testing.assert_array_equal(c[:tables.EV_SIZE], ev.counters())
    np.testing.assert_array_equal(c[tables.EV_SIZE:], t3.counters())
    assert c[tables.EV_NGT] > 10000 and c[tables.EV_HITS + 2] > 0
    got = pipe.metrics()
    np.testing.assert_allclose(helpers.flat_metrics(got["evaluator"]), helpers.flat_metrics(m), rtol=0, atol=0, equal_nan=True)
    np.testing.assert_allclose(helpers.flat_metrics(got["top3"]), helpers.flat_metrics(m3), rtol=0, atol=0, equal_nan=True)
    s = pipe.stats.cpu().numpy()
    want = [stats[k] for k in ("num_not_connected", "num_connected", "num_connected_pred", "connectivity_precision", "connectivity_recall")]
    np.testing.assert_array_equal(s.astype(np.float64), np.asarray([float(x) for x in want]))


def test_cfg3_full_size_sgdet_counters_bit_exact_vs_oracle():
    """BASELINE config 3 shape (100 proposals / image = 9 900 directed pairs and 29 700 candidates per image, 20 GT boxes,
    object-confidence add, synonym matching, top-100): counters equal the loop oracle's on identical scores."""
    from oracle import hiercom_oracle as O
    from scene_graph_commonsense_b200 import pipeline
    from tests import helpers
    from tests.test_gpu_eval import _scores_for_pairs
    batch = [synthetic.make_sgdet_image(i, 20, 100, p_rel=0.3, with_maps=False) for i in range(400, 404)]
    pipe = pipeline.RelationPipeline(None, DEV, commonsense=True, predcls=False)
    b = pipeline.batch_from_samples(batch, DEV, skip_mode="batch", sgdet=True, with_maps=False)
    pairs = pipe.enumerate_pairs(b)
    rel, sup, conn, logsig = _scores_for_pairs(batch, b, pairs, dict(gain=3.0))
    pipe.evaluate(b, pairs, rel, sup, logsig)
    ev, _ = helpers.oracle_evaluators(dict(run_mode="eval_cs"), True)
    n = O.replay_sgdet(batch, synthetic.batch_score_fn(batch, helpers.SPLITS, gain=3.0), ev, features=False)
    ev.compute(per_class=True, predcls=False)
    assert n == pairs["n"] and n > 30000
    c = pipe.counters.cpu().numpy()
    np.testing.assert_array_equal(c[:tables.EV_SIZE], ev.counters())
    assert c[tables.EV_NGT] > 0
