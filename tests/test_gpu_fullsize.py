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
