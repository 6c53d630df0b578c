"""Golden vectors for the SGB twin (R14 / N1), produced by the REAL Scene-Graph-Benchmark code of the reference.

Run in the build container only:  python oracle/make_golden_sgb.py      -> tests/golden/sgb_case*.npz
The SGB package imports a native extension (`maskrcnn_benchmark._C`, detector ops) and optional packages (yacs, pycocotools,
h5py, replicate, apex ...) that are neither built nor installed here and are NOT on the relation path; they are replaced by
mocks before import.  The relation-path classes are then exercised unmodified:
  MotifHierarchicalPredictor.forward (tail after the context layer; the LSTM context encoder is replaced by a stub that
      returns the seeded edge_ctx, it is outside the path), with the real BayesHead and FrequencyBias
  RelationSampling.prepare_test_pairs, HierarchPostProcessor.forward (real code; CommonsenseValidator replaced by the
      deterministic synthetic.sgb_validator - the LLM is outside the path)
  SGRecall.calculate_recall, SGMeanRecall.collect_mean_recall_items / calculate_mean_recall
torch.sort is patched to stable=True while the reference sorts (H1).
"""
import os
import sys
from unittest.mock import MagicMock

import numpy as np
import torch
import torch.nn as nn

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SGB = os.path.join(os.environ.get("HIERCOM_REFERENCE", "/root/reference"), "scenegraph_benchmark", "Scene-Graph-Benchmark.pytorch")
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)
sys.path.insert(0, SGB)
for name in ['maskrcnn_benchmark._C', 'pycocotools', 'pycocotools.mask', 'pycocotools.coco', 'pycocotools.cocoeval', 'h5py', 'replicate',
             'apex', 'apex.amp', 'yacs', 'yacs.config', 'matplotlib', 'matplotlib.pyplot']:
    sys.modules[name] = MagicMock()


class _CfgNode(dict):
    def __init__(self, *a, **k):
        super().__init__()

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return self


sys.modules['yacs.config'].CfgNode = _CfgNode

from maskrcnn_benchmark.data.datasets.evaluation.vg import sgg_eval as ref_eval          # noqa: E402
from maskrcnn_benchmark.modeling.roi_heads.relation_head import inference as ref_inf      # noqa: E402
from maskrcnn_benchmark.modeling.roi_heads.relation_head import model_motifs as ref_motifs  # noqa: E402
from maskrcnn_benchmark.modeling.roi_heads.relation_head import model_motifs_hierarchical as ref_hier  # noqa: E402
from maskrcnn_benchmark.modeling.roi_heads.relation_head import roi_relation_predictors as ref_pred    # noqa: E402
from maskrcnn_benchmark.modeling.roi_heads.relation_head import sampling as ref_sampling  # noqa: E402
from maskrcnn_benchmark.structures.bounding_box import BoxList                            # noqa: E402

from scene_graph_commonsense_b200 import synthetic                                        # noqa: E402
from tests.golden_cases import SGB_CASES                                                  # noqa: E402

_orig_sort = torch.sort


def _stable_sort(x, dim=-1, descending=False, stable=False):
    return _orig_sort(x, dim=dim, descending=descending, stable=True)


class _Ctx(nn.Module):
    def __init__(self, edge_ctx, obj_preds):
        super().__init__()
        self.edge_ctx, self.obj_preds = edge_ctx, obj_preds

    def forward(self, roi_features, proposals, logger=None):
        return torch.zeros(self.edge_ctx.shape[0], 151), self.obj_preds, self.edge_ctx, None


def build_predictor(sd, batch):
    p = object.__new__(ref_pred.MotifHierarchicalPredictor)
    nn.Module.__init__(p)
    p.attribute_on, p.use_vision, p.use_bias, p.union_single_not_match = False, True, True, False
    p.hidden_dim, p.pooling_dim = 512, 4096
    p.post_emb, p.post_cat = nn.Linear(512, 1024), nn.Linear(1024, 4096)
    p.rel_compress = ref_hier.BayesHead(input_dim=4096)
    fb = object.__new__(ref_motifs.FrequencyBias)
    nn.Module.__init__(fb)
    fb.num_objs, fb.num_rels = 151, 51
    fb.obj_baseline = nn.Embedding(151 * 151, 51)
    p.freq_bias = fb
    p.geo_label_tensor = torch.tensor(ref_pred_labels('geo'))
    p.pos_label_tensor = torch.tensor(ref_pred_labels('pos'))
    p.sem_label_tensor = torch.tensor(ref_pred_labels('sem'))
    p.context_layer = _Ctx(batch["edge_ctx"], batch["obj_labels"])
    with torch.no_grad():
        p.post_emb.weight.copy_(sd["post_emb.weight"]); p.post_emb.bias.copy_(sd["post_emb.bias"])
        p.post_cat.weight.copy_(sd["post_cat.weight"]); p.post_cat.bias.copy_(sd["post_cat.bias"])
        for n in ("fc3_1", "fc3_2", "fc3_3", "fc5"):
            getattr(p.rel_compress, n).weight.copy_(sd[n + ".weight"]); getattr(p.rel_compress, n).bias.copy_(sd[n + ".bias"])
        fb.obj_baseline.weight.copy_(sd["freq_bias"])
    return p.eval()


class _CtxT(_Ctx):
    """TransformerContext stand-in: 3-tuple (roi_relation_predictors.py:200)."""

    def forward(self, roi_features, proposals, logger=None):
        return torch.zeros(self.edge_ctx.shape[0], 151), self.obj_preds, self.edge_ctx


class _CtxV(_Ctx):
    """VCTreeLSTMContext stand-in: takes rel_pair_idxs, returns binary_preds last (roi_relation_predictors.py:649)."""

    def forward(self, roi_features, proposals, rel_pair_idxs, logger=None):
        return torch.zeros(self.edge_ctx.shape[0], 151), self.obj_preds, self.edge_ctx, None


def _copy_head(head, sd, prefix=""):
    with torch.no_grad():
        for n in ("fc3_1", "fc3_2", "fc3_3", "fc5"):
            getattr(head, n).weight.copy_(sd[prefix + n + ".weight"]); getattr(head, n).bias.copy_(sd[prefix + n + ".bias"])


def build_variant(kind, sd, sd_ctx, batch):
    """The REAL TransformerHierPredictor / VCTreeHierPredictor forward (tail after a stub context layer)."""
    cls = {"transformer": ref_pred.TransformerHierPredictor, "vctree": ref_pred.VCTreeHierPredictor}[kind]
    p = object.__new__(cls)
    nn.Module.__init__(p)
    p.attribute_on, p.use_vision, p.use_bias, p.union_single_not_match = False, True, True, False
    p.hidden_dim, p.pooling_dim = 512, 4096
    p.post_emb, p.post_cat = nn.Linear(512, 1024), nn.Linear(1024, 4096)
    with torch.no_grad():
        p.post_emb.weight.copy_(sd["post_emb.weight"]); p.post_emb.bias.copy_(sd["post_emb.bias"])
        p.post_cat.weight.copy_(sd["post_cat.weight"]); p.post_cat.bias.copy_(sd["post_cat.bias"])
    if kind == "transformer":
        p.rel_compress, p.ctx_compress = ref_hier.BayesHead(4096), ref_hier.BayesHead(1024)
        _copy_head(p.rel_compress, sd)
        _copy_head(p.ctx_compress, sd_ctx)
        p.context_layer = _CtxT(batch["edge_ctx"], batch["obj_labels"])
    else:
        p.ctx_compress = ref_hier.BayesHeadProb(4096)
        _copy_head(p.ctx_compress, sd)
        p.context_layer = _CtxV(batch["edge_ctx"], batch["obj_labels"])
    return p.eval()


def gen_variants():
    """sgb_variants.npz: Transformer / VCTree hierarchical predictors through the real SGB forward (SGB_VARIANT_CASE)."""
    from tests.golden_cases import SGB_VARIANT_CASE as c
    batch = synthetic.make_sgb_batch(c["num_objs"], seed=c["seed"])
    sd = synthetic.sgb_state_dict(seed=c["seed"])
    sd_ctx = synthetic.sgb_state_dict(seed=c["seed"] + 100, pooling=1024)          # a BayesHead over the 1024-d pair representation
    sampler = object.__new__(ref_sampling.RelationSampling)
    sampler.use_gt_box, sampler.test_overlap = True, False
    proposals = [BoxList(b, (800, 600), 'xyxy') for b in batch["boxes"]]
    rel_pair_idxs = sampler.prepare_test_pairs('cpu', proposals)
    out = {}
    for kind in ("transformer", "vctree"):
        pred = build_variant(kind, sd, sd_ctx, batch)
        with torch.no_grad():
            _, r1, r2, r3, sup, _ = pred(proposals, rel_pair_idxs, None, None, None, batch["union_features"], None)
        for i in range(len(c["num_objs"])):
            out["%s_rel_%d" % (kind, i)] = torch.cat((r1[i], r2[i], r3[i]), dim=1).numpy()
            out["%s_sup_%d" % (kind, i)] = sup[i].numpy()
        print("variant", kind, "max joint prob", float(torch.cat([torch.cat((a, b, c_), 1) for a, b, c_ in zip(r1, r2, r3)]).exp().max()))
    np.savez_compressed(os.path.join(OUT, "sgb_variants.npz"), **out)


def ref_pred_labels(which):
    from scene_graph_commonsense_b200 import sgb
    return {'geo': sgb.GEO_LABEL, 'pos': sgb.POS_LABEL, 'sem': sgb.SEM_LABEL}[which]


class _FakeLLM:
    top_k = 10

    def query(self, combined_obj_label, rel_labels, image, boxlist):
        return synthetic.sgb_validator(combined_obj_label, rel_labels)


def build_postprocessor():
    pp = object.__new__(ref_inf.HierarchPostProcessor)
    nn.Module.__init__(pp)
    pp.attribute_on, pp.use_gt_box, pp.later_nms_pred_thres = False, True, 0.3
    pp.geo_label_tensor = torch.tensor(ref_pred_labels('geo'))
    pp.pos_label_tensor = torch.tensor(ref_pred_labels('pos'))
    pp.sem_label_tensor = torch.tensor(ref_pred_labels('sem'))
    pp.llm, pp.skip_top = _FakeLLM(), 10
    return pp


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, c in SGB_CASES.items():
        batch = synthetic.make_sgb_batch(c["num_objs"], seed=c["seed"])
        sd = synthetic.sgb_state_dict(seed=c["seed"])
        num_objs = batch["num_objs"]
        sampler = object.__new__(ref_sampling.RelationSampling)
        sampler.use_gt_box, sampler.test_overlap = True, False
        proposals = [BoxList(b, (800, 600), 'xyxy') for b in batch["boxes"]]
        rel_pair_idxs = sampler.prepare_test_pairs('cpu', proposals)
        pred = build_predictor(sd, batch)
        with torch.no_grad():
            _, r1, r2, r3, sup, _ = pred(proposals, rel_pair_idxs, None, None, None, batch["union_features"], None)
        pp = build_postprocessor()
        refine = batch["obj_logits"].split(num_objs, 0)
        torch.sort = _stable_sort
        try:
            with torch.no_grad():
                results = pp((r1, r2, r3, sup, refine), rel_pair_idxs, proposals, [None] * len(num_objs))
        finally:
            torch.sort = _orig_sort
        # GT relations: half of them copy one of the image's ranked predictions (so matches exist), half are random
        g = torch.Generator().manual_seed(77 + c["seed"])
        result_dict = {}
        ev_r = ref_eval.SGRecall(result_dict); ev_r.register_container('predcls')
        ev_m = ref_eval.SGMeanRecall(result_dict, 51, ['__background__'] + ['p%d' % i for i in range(1, 51)]); ev_m.register_container('predcls')
        out = {}
        for i, res in enumerate(results):
            n = num_objs[i]
            pri = res.get_field('rel_pair_idxs'); prl = res.get_field('pred_rel_labels')
            n_gt = 0 if (c.get("empty_gt") == i) else max(2, n)
            gt = []
            for j in range(n_gt):
                if j % 2 == 0:
                    r = int(torch.randint(0, min(len(pri), 150), (1,), generator=g))
                    gt.append([int(pri[r, 0]), int(pri[r, 1]), int(prl[r])])
                else:
                    a, b = torch.randperm(n, generator=g)[:2].tolist()
                    gt.append([a, b, int(torch.randint(1, 51, (1,), generator=g))])
            gt_rels = np.asarray(gt, dtype=np.int64).reshape(-1, 3)
            gt_classes = res.get_field('pred_labels').numpy()      # PredCLS: labels are given; any consistent table works
            gt_boxes = batch["boxes"][i].numpy()
            out["gt_rels_%d" % i] = gt_rels
            out["gt_classes_%d" % i] = gt_classes
            out["rel_pair_idxs_%d" % i] = pri.numpy(); out["pred_rel_labels_%d" % i] = prl.numpy()
            out["pred_rel_scores_%d" % i] = res.get_field('pred_rel_scores').numpy().astype(np.float32)
            out["pred_labels_%d" % i] = res.get_field('pred_labels').numpy(); out["pred_scores_%d" % i] = res.get_field('pred_scores').numpy()
            out["rel1_%d" % i], out["rel2_%d" % i], out["rel3_%d" % i], out["sup_%d" % i] = (t[i].numpy() for t in (r1, r2, r3, sup))
            if len(gt_rels) == 0:
                continue
            lc = dict(gt_rels=gt_rels, gt_classes=gt_classes, gt_boxes=gt_boxes, pred_rel_inds=pri.numpy(),
                      rel_scores=res.get_field('pred_rel_scores').numpy(), pred_rel_labels=prl.numpy(), pred_boxes=gt_boxes,
                      pred_classes=gt_classes, obj_scores=np.ones(n))
            lc = ev_r.calculate_recall({'iou_thres': 0.5}, lc, 'predcls')
            ev_m.collect_mean_recall_items({'iou_thres': 0.5}, lc, 'predcls')
        ev_m.calculate_mean_recall('predcls')
        out["recall"] = np.array([np.mean(result_dict['predcls_recall'][k]) for k in (20, 50, 100)])
        out["recall_per_image"] = np.array([result_dict['predcls_recall'][k] for k in (20, 50, 100)])
        out["mean_recall"] = np.array([result_dict['predcls_mean_recall'][k] for k in (20, 50, 100)])
        out["mean_recall_list"] = np.array([result_dict['predcls_mean_recall_list'][k] for k in (20, 50, 100)])
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, "R@K", out["recall"], "mR@K", out["mean_recall"])


if __name__ == "__main__":
    if len(sys.argv) < 2 or "cases" in sys.argv:
        main()
    if len(sys.argv) < 2 or "variants" in sys.argv:
        gen_variants()
