"""Pins the oracle to goldens produced END TO END by the unmodified reference (oracle/make_golden.py `real`):
BayesianRelationClassifier -> train_utils.evaluate_one_direction -> Evaluator / Evaluator_Top3 on images whose object
classes have 1, 2 and 3 super-classes, with an empty box, ragged sizes and the whole-batch skip rule.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import hiercom_oracle as O
from scene_graph_commonsense_b200 import synthetic, tables
from tests import helpers
from tests.golden_cases import HEAD3_CASE, REAL_PIPELINE_CASE, install_gt, real_pipeline_samples


def test_process_super_class_is_first_plus_last_entry():
    """utils.py:136-149: a 3-entry list [4,5,6] sets {4,6}, not {4,5,6}; duplicates add up; 5 entries keep only s[0]."""
    got = O.process_super_class([[4], [4, 5], [4, 5, 6], [4, 5, 6, 7], [3, 3], [2, 9, 2], [1, 2, 3, 4, 5]]).numpy()
    want = np.zeros((7, 17), dtype=np.int64)
    for r, idx in enumerate(([4], [4, 5], [4, 6], [4, 7], [3, 3], [2, 2], [1])):
        for i in idx:
            want[r, i] += 1
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("preset", HEAD3_CASE["presets"])
def test_head_with_three_super_classes_matches_reference(preset):
    c = HEAD3_CASE
    g = helpers.golden("head3")
    s = synthetic.with_categories(synthetic.make_image(c["id"], c["n"]), c["cats"])
    assert sorted({len(x) for x in s.super_categories}) == [1, 2, 3]
    hs = torch.stack([O._masked_input(s, s.bbox[a]) for a, b in c["pairs"]])
    ho = torch.stack([O._masked_input(s, s.bbox[b]) for a, b in c["pairs"]])
    c1 = torch.stack([s.categories[a] for a, b in c["pairs"]])
    c2 = torch.stack([s.categories[b] for a, b in c["pairs"]])
    s1 = [s.super_categories[a] for a, b in c["pairs"]]
    s2 = [s.super_categories[b] for a, b in c["pairs"]]
    sd = synthetic.preset_state_dict(preset)
    r1, r2, r3, sup, conn, pred = O.bayesian_relation_classifier(sd, hs, ho, c1, c2, s1, s2)
    np.testing.assert_allclose(torch.cat((r1, r2, r3), 1).numpy(), g[preset + "_relation"], atol=5e-5, rtol=0)
    np.testing.assert_allclose(sup.numpy(), g[preset + "_super"], atol=5e-5, rtol=0)
    np.testing.assert_allclose(pred.numpy(), g[preset + "_pred"], atol=2e-5, rtol=1e-5)
    # and the sum over ALL entries (the round-1 misreading) is measurably different on these inputs
    hc_wrong = torch.cat((torch.zeros(len(s1), 4096 + 300), torch.stack([torch.bincount(x, minlength=17) for x in s1]).float(),
                          torch.stack([torch.bincount(x, minlength=17) for x in s2]).float()), 1)
    hc_right = torch.cat((torch.zeros(len(s1), 4096 + 300), O.process_super_class(s1).float(), O.process_super_class(s2).float()), 1)
    assert (hc_wrong != hc_right).any()


@pytest.mark.parametrize("preset", REAL_PIPELINE_CASE["presets"])
def test_pipeline_through_real_reference(preset):
    g = helpers.golden("pl_real")
    samples = install_gt(real_pipeline_samples(), g[preset + "_gt"])
    sd = synthetic.preset_state_dict(preset)
    rows = {}
    base = O.make_head_fn(sd)

    def head_fn(h_sub, h_obj, c1, c2, s1, s2, ctx):
        out = base(h_sub, h_obj, c1, c2, s1, s2)
        keep, sub, obj = ctx
        for r, i in enumerate(keep):
            rows[(int(i), int(sub), int(obj))] = (out[0][r].numpy(), out[1][r].numpy(), out[2][r].numpy())
        return out
    ev = O.OracleEvaluator(helpers.SPLITS, True, aligned=set(tables.commonsense_aligned_keys().tolist()),
                           violated=set(tables.commonsense_violated_keys().tolist()), zero_shot=set(tables.zero_shot_keys().tolist()))
    t3 = O.OracleEvaluatorTop3(helpers.SPLITS)
    stats = dict(num_not_connected=0, num_connected=0, num_connected_pred=0, connectivity_precision=0, connectivity_recall=0)
    O.replay_predcls(samples, head_fn, ev, t3, stats=stats)
    m, m3 = ev.compute(per_class=True), t3.compute(per_class=True)
    keys = [tuple(int(v) for v in k) for k in g[preset + "_keys"]]
    assert sorted(rows) == keys                      # same directed pairs survive the skip rule
    np.testing.assert_allclose(np.stack([rows[k][0] for k in keys]), g[preset + "_relation"], atol=5e-5, rtol=0)
    np.testing.assert_allclose(np.stack([rows[k][1] for k in keys]), g[preset + "_super"], atol=5e-5, rtol=0)
    np.testing.assert_array_equal(ev.counters(), g[preset + "_ev"])
    np.testing.assert_array_equal(t3.counters(), g[preset + "_t3"])
    assert g[preset + "_ev"][0] > 0 and g[preset + "_ev"][153] > g[preset + "_ev"][2]      # real hits, and real misses
    np.testing.assert_allclose(helpers.flat_metrics(m), g[preset + "_metrics"], rtol=0, atol=0, equal_nan=True)
    np.testing.assert_allclose(helpers.flat_metrics(m3), g[preset + "_metrics3"], rtol=0, atol=0, equal_nan=True)
    got = [stats["num_not_connected"], stats["num_connected"], stats["num_connected_pred"], stats["connectivity_precision"],
           stats["connectivity_recall"]]
    np.testing.assert_array_equal(np.asarray(got, dtype=np.float64), g[preset + "_stats"])
