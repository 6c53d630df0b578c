"""GPU parity for the configuration variants the reference supports on the path (SURVEY §2a "Label/cluster maps",
main.py:56-84): other (num_geo, num_pos, num_sem) splits (gpt2 9/32/9, bert 12/25/13, clip 27/15/8, OIv6 4/2/24), the
fc2 layout without super-class columns (model.py:125-128), temperatures != 1, and degenerate windows (images with zero or
one box, a window with no surviving pair)."""
import numpy as np
import pytest
import torch

from oracle import hiercom_oracle as O
from scene_graph_commonsense_b200 import synthetic, tables
from tests import helpers

pytestmark = pytest.mark.gpu
DEV = "cuda"

SPLIT_CASES = [(9, 32, 9), (12, 25, 13), (27, 15, 8), (4, 2, 24), (15, 11, 24)]


def _torch_head(pred, w, b, splits, temps):
    """model.py:176-184 in torch fp64->fp32: super = log_softmax(fc5), rel_k = log_softmax(fc3_k / T_k) + super[:, k]."""
    g, p, s = splits
    R = g + p + s
    z = (pred.double() @ w.double().t() + b.double())
    sup = torch.log_softmax(z[:, R + 1:R + 4], dim=1)
    outs, a = [], 0
    for k, n in enumerate(splits):
        outs.append(torch.log_softmax(z[:, a:a + n] / temps[k], dim=1) + sup[:, k:k + 1])
        a += n
    return torch.cat(outs, 1).float(), sup.float(), z[:, R].float()


@pytest.mark.parametrize("splits", SPLIT_CASES)
@pytest.mark.parametrize("n_rows", [1, 3, 64, 517])
def test_head_and_candidates_with_other_splits(splits, n_rows):
    from scene_graph_commonsense_b200 import ops
    g = torch.Generator().manual_seed(sum(splits) * 7 + n_rows)
    R = sum(splits)
    temps = (1.0, 0.7, 1.6)
    raw = torch.randn(n_rows, 512, generator=g)
    w = torch.randn(R + 4, 512, generator=g) * 0.2
    b = torch.randn(R + 4, generator=g) * 0.1
    emb = torch.randn(300, 512, generator=g) * 0.3                   # non-VG layout: one-hot columns only (model.py:125-128)
    fc2_b = torch.randn(512, generator=g) * 0.1
    cats = torch.randint(0, 150, (23,), generator=g)
    row_sub = torch.randint(0, 23, (n_rows,), generator=g)
    row_obj = torch.randint(0, 23, (n_rows,), generator=g)
    c = lambda t, dt=None: (t.to(dt) if dt else t).to(DEV).contiguous()
    rel, sup, conn, logsig, pred = ops.hier_head(c(raw), c(fc2_b), c(emb), c(row_sub, torch.int32), c(row_obj, torch.int32),
                                                 c(cats, torch.int32), None, c(w), c(b), splits, temps=temps, want_pred=True)
    p_ref = torch.relu(raw + fc2_b + emb[cats[row_sub]] + emb[150 + cats[row_obj]])
    np.testing.assert_allclose(pred.cpu().numpy(), p_ref.numpy(), atol=1e-5, rtol=1e-5)
    rel_ref, sup_ref, conn_ref = _torch_head(p_ref, w, b, splits, temps)
    np.testing.assert_allclose(rel.cpu().numpy(), rel_ref.numpy(), atol=2e-4, rtol=0)
    np.testing.assert_allclose(sup.cpu().numpy(), sup_ref.numpy(), atol=2e-4, rtol=0)
    np.testing.assert_allclose(conn.cpu().numpy(), conn_ref.numpy(), atol=2e-4, rtol=0)
    # candidates on the kernel's own scores: per-super max / first argmax, bit-exact against numpy on the same floats
    ov = (torch.rand(n_rows, generator=g) < 0.8).to(torch.uint8)
    conf, label, t3c, t3s = ops.candidates(rel, splits, True, c(ov), logsig, c(row_sub, torch.int32), c(row_obj, torch.int32),
                                           c(cats, torch.int32), None, sup, want_top3=True)
    relh, lsh, suph = rel.cpu().numpy(), logsig.cpu().numpy(), sup.cpu().numpy()
    a = 0
    for k, n in enumerate(splits):
        seg = relh[:, a:a + n]
        want_label = seg.argmax(1) + a
        want_conf = np.where(ov.numpy() != 0, seg.max(1), -np.inf).astype(np.float32) + lsh
        np.testing.assert_array_equal(label.cpu().numpy()[k::3], want_label)
        np.testing.assert_array_equal(conf.cpu().numpy()[k::3], want_conf)
        a += n
    np.testing.assert_array_equal(t3s.cpu().numpy(), suph.argmax(1))


@pytest.mark.parametrize("splits", [(9, 32, 9), (4, 2, 24)])
def test_counters_with_other_splits_match_oracle(splits):
    """Full integer stage (candidates -> filter -> top-K -> match -> counters) under a non-default clustering."""
    from scene_graph_commonsense_b200 import pipeline
    ids, ns = [400, 401, 402], [11, 6, 14]
    samples = synthetic.make_batch(ids, ns, with_maps=False, p_rel=0.5)
    num_pred = sum(splits)
    for s in samples:                                              # keep GT labels inside the variant's predicate range
        s.relationships = [torch.where(r >= 0, r % num_pred, r) for r in s.relationships]
    al, vi = synthetic.synthetic_cs_keys(13, 0.5, 0.1)
    pipe = pipeline.RelationPipeline(None, DEV, commonsense=True, aligned_keys=al, violated_keys=vi, splits=splits, hier=True)
    b = pipeline.batch_from_samples(samples, DEV, skip_mode="batch", with_maps=False)
    pairs = pipe.enumerate_pairs(b)
    off = b.box_offsets.cpu().numpy()
    rel, sup, conn = [], [], []
    for i, s_, o_ in zip(pairs["img"].cpu().numpy(), pairs["sub"].cpu().numpy(), pairs["obj"].cpu().numpy()):
        r = synthetic.pair_scores(samples[i].image_id, int(s_ - off[i]), int(o_ - off[i]), splits, gain=3.0)
        rel.append(r[0]); sup.append(r[1]); conn.append(r[2])
    rel, sup, conn = torch.stack(rel), torch.stack(sup), torch.cat(conn)
    pipe.evaluate(b, pairs, rel.to(DEV), sup.to(DEV), torch.log(torch.sigmoid(conn)).to(DEV), connectivity=conn.to(DEV))
    ev = O.OracleEvaluator(splits, hierar=True, aligned=set(al.tolist()), violated=set(vi.tolist()),
                           zero_shot=set(tables.zero_shot_keys().tolist()))
    t3 = O.OracleEvaluatorTop3(splits)
    O.replay_predcls(samples, synthetic.batch_score_fn(samples, splits, gain=3.0), ev, t3, features=False)
    ev.compute(per_class=True); t3.compute(per_class=True)
    cnt = pipe.counters.cpu().numpy()
    np.testing.assert_array_equal(cnt[:tables.EV_SIZE], ev.counters())
    np.testing.assert_array_equal(cnt[tables.EV_SIZE:], t3.counters())
    assert cnt[tables.EV_NGT] > 0


def test_degenerate_windows_zero_and_one_box_images():
    """Images with 0 or 1 boxes contribute no pairs and no GT; a window where nothing survives returns 0 pairs."""
    from scene_graph_commonsense_b200 import pipeline
    samples = synthetic.make_batch([410, 411, 412, 413], [1, 7, 2, 5], with_maps=False, p_rel=0.6)
    empty = synthetic.make_image(414, 2, with_maps=False)
    empty.bbox, empty.categories, empty.super_categories = empty.bbox[:0], empty.categories[:0], []
    empty.relationships, empty.subj_or_obj = [], []
    samples.insert(2, empty)
    al, vi = synthetic.synthetic_cs_keys(14, 0.5, 0.1)
    for mode in ("batch", "per_image"):
        pipe = pipeline.RelationPipeline(None, DEV, commonsense=True, aligned_keys=al, violated_keys=vi)
        b = pipeline.batch_from_samples(samples, DEV, skip_mode=mode, with_maps=False)
        pairs = pipe.enumerate_pairs(b)
        po = pairs["offsets_host"]
        assert po[1] - po[0] == 0 and po[3] - po[2] == 0           # the 1-box and the 0-box image
        off = b.box_offsets.cpu().numpy()
        rel, sup, conn = [], [], []
        for i, s_, o_ in zip(pairs["img"].cpu().numpy(), pairs["sub"].cpu().numpy(), pairs["obj"].cpu().numpy()):
            r = synthetic.pair_scores(samples[i].image_id, int(s_ - off[i]), int(o_ - off[i]), helpers.SPLITS, gain=3.0)
            rel.append(r[0]); sup.append(r[1]); conn.append(r[2])
        rel, sup, conn = torch.stack(rel), torch.stack(sup), torch.cat(conn)
        pipe.evaluate(b, pairs, rel.to(DEV), sup.to(DEV), torch.log(torch.sigmoid(conn)).to(DEV), connectivity=conn.to(DEV))
        ev = O.OracleEvaluator(helpers.SPLITS, hierar=True, aligned=set(al.tolist()), violated=set(vi.tolist()),
                               zero_shot=set(tables.zero_shot_keys().tolist()))
        t3 = O.OracleEvaluatorTop3(helpers.SPLITS)
        real = [s for s in samples if len(s.categories) >= 2]
        if mode == "batch":
            O.replay_predcls(real, synthetic.batch_score_fn(real, helpers.SPLITS, gain=3.0), ev, t3, features=False)
            ev.compute(per_class=True); t3.compute(per_class=True)
        else:
            for s in real:
                O.replay_predcls([s], synthetic.batch_score_fn([s], helpers.SPLITS, gain=3.0), ev, t3, features=False)
                ev.compute(per_class=True); ev.clear_data(); t3.compute(per_class=True); t3.clear_data()
        cnt = pipe.counters.cpu().numpy()
        np.testing.assert_array_equal(cnt[:tables.EV_SIZE], ev.counters())
        np.testing.assert_array_equal(cnt[tables.EV_SIZE:], t3.counters())
    # nothing survives: two far-apart boxes per image, per_image mode
    far = synthetic.make_batch([420, 421], [2, 2], with_maps=False, p_rel=1.0)
    for s in far:
        s.bbox = torch.tensor([[0, 4, 0, 4], [20, 28, 20, 28]], dtype=torch.int32)
    pipe = pipeline.RelationPipeline(None, DEV, commonsense=False)
    b = pipeline.batch_from_samples(far, DEV, skip_mode="per_image", with_maps=False)
    assert pipe.enumerate_pairs(b)["n"] == 0
    assert pipe.step(b) == 0 and int(pipe.counters.sum()) == 0
