"""GPU parity of the Scene-Graph-Benchmark twin (R14 / N1) against goldens produced by the real SGB code."""
import numpy as np
import pytest
import torch

from scene_graph_commonsense_b200 import synthetic
from tests import helpers
from tests.golden_cases import SGB_CASES

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _load(name):
    c = SGB_CASES[name]
    return c, helpers.golden(name), synthetic.make_sgb_batch(c["num_objs"], seed=c["seed"]), synthetic.sgb_state_dict(seed=c["seed"])


def _pairs(num_objs):
    out = []
    for n in num_objs:
        idx = torch.nonzero(torch.ones(n, n) - torch.eye(n)).view(-1, 2)
        out.append(idx)
    return out


@pytest.mark.parametrize("name", sorted(SGB_CASES))
def test_relation_tail_matches_real_sgb_within_2e3(name):
    from scene_graph_commonsense_b200 import sgb
    c, g, batch, sd = _load(name)
    post_cat = torch.nn.Linear(1024, 4096).to(DEV)
    head = sgb.BayesHead(input_dim=4096).to(DEV)
    with torch.no_grad():
        post_cat.weight.copy_(sd["post_cat.weight"]); post_cat.bias.copy_(sd["post_cat.bias"])
        for n in ("fc3_1", "fc3_2", "fc3_3", "fc5"):
            getattr(head, n).weight.copy_(sd[n + ".weight"]); getattr(head, n).bias.copy_(sd[n + ".bias"])
    edge_rep = torch.nn.functional.linear(batch["edge_ctx"], sd["post_emb.weight"], sd["post_emb.bias"]).to(DEV)   # upstream of the path
    r1, r2, r3, sup = sgb.hierarchical_relation_tail(edge_rep, _pairs(batch["num_objs"]), batch["num_objs"], batch["obj_labels"].to(DEV),
                                                     batch["union_features"].to(DEV), post_cat, head, sd["freq_bias"].to(DEV))
    worst = 0.0
    for i in range(len(batch["num_objs"])):
        for got, key in ((r1, "rel1"), (r2, "rel2"), (r3, "rel3"), (sup, "sup")):
            worst = max(worst, float(np.abs(np.exp(got[i].cpu().numpy().astype(np.float64)) - np.exp(g["%s_%d" % (key, i)].astype(np.float64))).max()))
    assert worst <= 2e-3, worst


def test_bayes_head_modules_match_torch():
    from scene_graph_commonsense_b200 import sgb
    h = torch.randn(300, 4096, generator=torch.Generator().manual_seed(1)).to(DEV)
    for cls in (sgb.BayesHead, sgb.BayesHeadProb):
        m = cls(input_dim=4096).to(DEV)
        m.layer_init()
        with torch.no_grad():
            for p in m.parameters():
                if p.dim() == 1:
                    p.normal_(0, 0.1)
        a = m(h)
        w = lambda l: torch.nn.functional.linear(h.double(), l.weight.double(), l.bias.double()).float()
        z1, z2, z3, z5 = w(m.fc3_1), w(m.fc3_2), w(m.fc3_3), w(m.fc5)
        if cls is sgb.BayesHead:
            ref = (z1, z2, z3, z5)
        else:
            s = torch.log_softmax(z5, 1)
            ref = (torch.log_softmax(z1, 1) + s[:, 1:2], torch.log_softmax(z2, 1) + s[:, 2:3], torch.log_softmax(z3, 1) + s[:, 3:4], s)
        for x, y in zip(a, ref):
            assert float((x - y).detach().abs().max()) <= 5e-4           # bf16x3 split operands: ~2^-16 relative on logits of magnitude ~4


@pytest.mark.parametrize("name", sorted(SGB_CASES))
def test_postprocessor_matches_real_sgb(name):
    from scene_graph_commonsense_b200 import sgb
    c, g, batch, sd = _load(name)
    nimg = len(batch["num_objs"])
    t = lambda k: [torch.from_numpy(g["%s_%d" % (k, i)]).to(DEV) for i in range(nimg)]
    pp = sgb.HierarchPostProcessor(False, use_gt_box=True, validator=synthetic.sgb_validator)
    boxes = [sgb.SimpleBoxList(b.to(DEV), (800, 600)) for b in batch["boxes"]]
    refine = [l.to(DEV) for l in batch["obj_logits"].split(batch["num_objs"], 0)]
    res = pp((t("rel1"), t("rel2"), t("rel3"), t("sup"), refine), [p.to(DEV) for p in _pairs(batch["num_objs"])], boxes, [None] * nimg)
    for i, r in enumerate(res):
        np.testing.assert_array_equal(r.get_field("pred_labels").cpu().numpy(), g["pred_labels_%d" % i])
        np.testing.assert_allclose(r.get_field("pred_scores").cpu().numpy(), g["pred_scores_%d" % i], atol=1e-6)
        np.testing.assert_array_equal(r.get_field("rel_pair_idxs").cpu().numpy(), g["rel_pair_idxs_%d" % i])
        np.testing.assert_array_equal(r.get_field("pred_rel_labels").cpu().numpy(), g["pred_rel_labels_%d" % i])
        np.testing.assert_allclose(r.get_field("pred_rel_scores").cpu().numpy(), g["pred_rel_scores_%d" % i], atol=1e-5)


@pytest.mark.parametrize("name", sorted(SGB_CASES))
def test_recall_kernels_match_real_sgrecall_and_sgmeanrecall(name):
    from scene_graph_commonsense_b200 import sgb
    c, g, batch, sd = _load(name)
    nimg = len(batch["num_objs"])
    t = lambda k: [torch.from_numpy(g["%s_%d" % (k, i)]).to(DEV) for i in range(nimg)]
    pp = sgb.HierarchPostProcessor(False, use_gt_box=True)
    refine = [l.to(DEV) for l in batch["obj_logits"].split(batch["num_objs"], 0)]
    cand = pp.candidates(t("rel1"), t("rel2"), t("rel3"), refine, [p.to(DEV) for p in _pairs(batch["num_objs"])])
    ev = sgb.SGBRecall()
    from scene_graph_commonsense_b200 import ops
    ranked = ops.topk_select((cand["pair_off"] * 3).contiguous(), cand["score"], 128)
    rej = ev.reject_mask(cand, ranked, synthetic.sgb_validator)
    gt_rels = [g["gt_rels_%d" % i] for i in range(nimg)]
    gt_cls = [g["gt_classes_%d" % i] for i in range(nimg)]
    final_rank, _ = ev.evaluate_batch(cand, gt_rels, gt_cls, batch["boxes"], reject=rej)
    # the ranked window after the validator equals the head of the reference's fully sorted list
    pair_off = cand["pair_off"].cpu().numpy()
    obj_off = np.concatenate(([0], np.cumsum(batch["num_objs"])))
    for i in range(nimg):
        ids = final_rank[i][final_rank[i] >= 0]
        gc = torch.from_numpy(ids).long().to(DEV) + 3 * int(pair_off[i])
        rows = cand["row"][gc].long()
        pairs_local = (cand["pair_idx"][rows] - int(obj_off[i])).cpu().numpy()
        np.testing.assert_array_equal(pairs_local, g["rel_pair_idxs_%d" % i][:len(ids)])
        np.testing.assert_array_equal(cand["label"][gc].cpu().numpy(), g["pred_rel_labels_%d" % i][:len(ids)])
    res = ev.result()
    np.testing.assert_array_equal(np.array([res["recall"][k] for k in (20, 50, 100)]), g["recall"])
    np.testing.assert_array_equal(np.array([res["mean_recall"][k] for k in (20, 50, 100)]), g["mean_recall"])
    np.testing.assert_array_equal(np.array([res["mean_recall_list"][k] for k in (20, 50, 100)]), g["mean_recall_list"])
