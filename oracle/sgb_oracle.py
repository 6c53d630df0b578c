"""CPU ORACLE for the Scene-Graph-Benchmark plug-and-play twin of the path (R14 / N1).  TEST INFRASTRUCTURE ONLY.

Restates, in torch-CPU / NumPy, the SGB code the HIERCOM fork adds or routes through
(SGB = scenegraph_benchmark/Scene-Graph-Benchmark.pytorch/maskrcnn_benchmark):
  predictor tail        SGB/modeling/roi_heads/relation_head/roi_relation_predictors.py:400-469
  BayesHead(Prob)       SGB/modeling/roi_heads/relation_head/model_motifs_hierarchical.py:22-33,56-66
  post-processor        SGB/modeling/roi_heads/relation_head/inference.py:187-312
  pair enumeration      SGB/modeling/roi_heads/relation_head/sampling.py:30-45
  recall                SGB/data/datasets/evaluation/vg/sgg_eval.py:56-99,347-385,494-565; structures/boxlist_ops.py:54-90

Pinned by `oracle/make_golden_sgb.py`, which imports the REAL SGB modules (native extension and optional packages
mocked, they are not on this path) and runs the real MotifHierarchicalPredictor.forward tail, HierarchPostProcessor.forward,
RelationSampling.prepare_test_pairs, SGRecall and SGMeanRecall on seeded inputs -> tests/golden/sgb_*.npz.
torch.sort is patched to stable=True while the reference sorts (H1).
"""
from functools import reduce

import numpy as np
import torch
import torch.nn.functional as F

GEO_LABEL = [1, 2, 3, 4, 5, 6, 8, 10, 22, 23, 29, 31, 32, 33, 43]
POS_LABEL = [9, 16, 17, 20, 27, 30, 36, 42, 48, 49, 50]
SEM_LABEL = [7, 11, 12, 13, 14, 15, 18, 19, 21, 24, 25, 26, 28, 34, 35, 37, 38, 39, 40, 41, 44, 45, 46, 47]


def prepare_test_pairs(num_objs):
    """sampling.py:30-45, use_gt_box branch: all ordered pairs in torch.nonzero(ones - eye) row-major order."""
    out = []
    for n in num_objs:
        m = torch.ones((n, n)) - torch.eye(n)
        idx = torch.nonzero(m).view(-1, 2)
        out.append(idx if len(idx) > 0 else torch.zeros((1, 2), dtype=torch.int64))
    return out


def predictor_tail(w, edge_ctx, rel_pair_idxs, num_objs, obj_preds, union_features, use_bias=True):
    """roi_relation_predictors.py:399-459.  `w`: dict of fp32 tensors post_emb.*, post_cat.*, fc3_1.*, fc3_2.*, fc3_3.*, fc5.*,
    freq_bias [151*151, 51]."""
    hidden = w["post_emb.weight"].shape[1]
    edge_rep = F.linear(edge_ctx, w["post_emb.weight"], w["post_emb.bias"]).view(-1, 2, hidden)
    head_reps = edge_rep[:, 0].contiguous().split(num_objs, 0)
    tail_reps = edge_rep[:, 1].contiguous().split(num_objs, 0)
    preds = obj_preds.split(num_objs, 0)
    prod, pair_pred = [], []
    for idx, h, t, op in zip(rel_pair_idxs, head_reps, tail_reps, preds):
        prod.append(torch.cat((h[idx[:, 0]], t[idx[:, 1]]), dim=-1))
        pair_pred.append(torch.stack((op[idx[:, 0]], op[idx[:, 1]]), dim=1))
    prod, pair_pred = torch.cat(prod), torch.cat(pair_pred)
    prod = F.linear(prod, w["post_cat.weight"], w["post_cat.bias"]) * union_features
    r1 = F.linear(prod, w["fc3_1.weight"], w["fc3_1.bias"])
    r2 = F.linear(prod, w["fc3_2.weight"], w["fc3_2.bias"])
    r3 = F.linear(prod, w["fc3_3.weight"], w["fc3_3.bias"])
    sup = F.linear(prod, w["fc5.weight"], w["fc5.bias"])
    if use_bias:
        bias = w["freq_bias"][pair_pred[:, 0].long() * 151 + pair_pred[:, 1].long()]
        b1, b2, b3 = bias[:, GEO_LABEL], bias[:, POS_LABEL], bias[:, SEM_LABEL]
        sb = torch.log(torch.stack((torch.exp(b1).sum(1), torch.exp(b2).sum(1), torch.exp(b3).sum(1)), dim=1))
        r1, r2, r3 = r1 + b1, r2 + b2, r3 + b3
        sup = sup.clone()
        sup[:, 1:] = sup[:, 1:] + sb
    sup = F.log_softmax(sup, dim=1)
    r1 = F.log_softmax(r1, dim=1) + sup[:, 1].view(-1, 1)
    r2 = F.log_softmax(r2, dim=1) + sup[:, 2].view(-1, 1)
    r3 = F.log_softmax(r3, dim=1) + sup[:, 3].view(-1, 1)
    num_rels = [r.shape[0] for r in rel_pair_idxs]
    return r1.split(num_rels, 0), r2.split(num_rels, 0), r3.split(num_rels, 0), sup.split(num_rels, 0)


def post_process_image(rel1, rel2, rel3, obj_logit, rel_pair_idx, validator=None, skip_top=10, llm_top_k=10):
    """inference.py:210-305 for use_gt_box=True.  Returns dict(obj_class, obj_scores, triple_scores (first sort),
    rel_pair_idx, rel_labels (after the second sort), rel_class_prob (first-sort order), first_order, second_order)."""
    obj_class_prob = F.softmax(obj_logit, -1).clone()
    obj_class_prob[:, 0] = 0
    obj_scores, obj_pred = obj_class_prob[:, 1:].max(dim=1)
    obj_class = obj_pred + 1
    s0, s1 = obj_scores[rel_pair_idx[:, 0]], obj_scores[rel_pair_idx[:, 1]]
    p1, p2, p3 = torch.exp(rel1), torch.exp(rel2), torch.exp(rel3)
    sc1, c1 = p1.max(dim=1)
    sc2, c2 = p2.max(dim=1)
    sc3, c3 = p3.max(dim=1)
    c1, c2, c3 = torch.tensor(GEO_LABEL)[c1], torch.tensor(POS_LABEL)[c2], torch.tensor(SEM_LABEL)[c3]
    cat_prob = torch.cat((p1, p2, p3), dim=1)
    cat_prob = torch.cat((cat_prob, cat_prob, cat_prob), dim=0)
    cat_pair = torch.cat((rel_pair_idx, rel_pair_idx, rel_pair_idx), dim=0)
    cat_s0, cat_s1 = torch.cat((s0, s0, s0)), torch.cat((s1, s1, s1))
    cat_labels, cat_scores = torch.cat((c1, c2, c3)), torch.cat((sc1, sc2, sc3))
    triple = cat_scores * cat_s0 * cat_s1
    triple_sorted, order = torch.sort(triple.view(-1), dim=0, descending=True, stable=True)
    pair_sorted, prob_sorted, labels_sorted = cat_pair[order], cat_prob[order], cat_labels[order]
    order2 = torch.arange(len(order))
    if validator is not None:
        a, b = skip_top, skip_top + llm_top_k
        combined = torch.stack((obj_class[pair_sorted[a:b, 0]], obj_class[pair_sorted[a:b, 1]]), dim=1)
        resp = torch.as_tensor(validator(combined, labels_sorted[a:b], None, None))
        triple_sorted = triple_sorted.clone()
        win = triple_sorted[a:b]
        win[resp == -1] = float("-inf")
        _, order2 = torch.sort(triple_sorted.view(-1), dim=0, descending=True, stable=True)
        pair_sorted, labels_sorted = pair_sorted[order2], labels_sorted[order2]
    return dict(obj_class=obj_class, obj_scores=obj_scores, triple_first=triple, first_order=order, second_order=order2,
                rel_pair_idx=pair_sorted, rel_labels=labels_sorted, rel_class_prob=prob_sorted, triple_sorted=triple_sorted)


def boxlist_iou(box1, box2):
    """boxlist_ops.py:54-90 (fp32 torch, TO_REMOVE = 1)."""
    box1, box2 = torch.as_tensor(box1, dtype=torch.float32), torch.as_tensor(box2, dtype=torch.float32)
    area1 = (box1[:, 2] - box1[:, 0] + 1) * (box1[:, 3] - box1[:, 1] + 1)
    area2 = (box2[:, 2] - box2[:, 0] + 1) * (box2[:, 3] - box2[:, 1] + 1)
    lt = torch.max(box1[:, None, :2], box2[:, :2])
    rb = torch.min(box1[:, None, 2:], box2[:, 2:])
    wh = (rb - lt + 1).clamp(min=0)
    inter = wh[:, :, 0] * wh[:, :, 1]
    return (inter / (area1[:, None] + area2 - inter)).numpy()


def compute_pred_matches(gt_triplets, pred_triplets, gt_boxes, pred_boxes, iou_thres):
    """sgg_eval.py:528-565 (phrdet=False)."""
    keeps = (gt_triplets[..., None] == pred_triplets.T[None, ...]).all(1)
    pred_to_gt = [[] for _ in range(pred_boxes.shape[0])]
    for gt_ind in np.where(keeps.any(1))[0]:
        keep_inds = keeps[gt_ind]
        boxes = pred_boxes[keep_inds]
        gt_box = gt_boxes[gt_ind]
        sub_iou = boxlist_iou(gt_box[None, :4], boxes[:, :4])[0]
        obj_iou = boxlist_iou(gt_box[None, 4:], boxes[:, 4:])[0]
        inds = (sub_iou >= iou_thres) & (obj_iou >= iou_thres)
        for i in np.where(keep_inds)[0][inds]:
            pred_to_gt[i].append(int(gt_ind))
    return pred_to_gt


class SGBRecallOracle:
    """SGRecall.calculate_recall (:56-99) + SGMeanRecall.collect_mean_recall_items / calculate_mean_recall (:347-385),
    PredCLS (pred boxes/classes = GT, vg_eval.py:267-270)."""

    def __init__(self, top_k=(20, 50, 100), iou_thres=0.5, num_rel=51):
        self.top_k, self.iou_thres, self.num_rel = top_k, iou_thres, num_rel
        self.recall = {k: [] for k in top_k}
        self.collect = {k: [[] for _ in range(num_rel)] for k in top_k}
        self.per_image = []

    def add_image(self, gt_rels, gt_classes, gt_boxes, pred_rel_inds, pred_rel_labels):
        gt_rels, gt_classes, gt_boxes = np.asarray(gt_rels), np.asarray(gt_classes), np.asarray(gt_boxes, dtype=np.float32)
        if len(gt_rels) == 0:
            return
        pred_rel_inds, pred_rel_labels = np.asarray(pred_rel_inds), np.asarray(pred_rel_labels)
        gt_trip = np.column_stack((gt_classes[gt_rels[:, 0]], gt_rels[:, 2], gt_classes[gt_rels[:, 1]]))
        gt_tb = np.column_stack((gt_boxes[gt_rels[:, 0]], gt_boxes[gt_rels[:, 1]]))
        pr_trip = np.column_stack((gt_classes[pred_rel_inds[:, 0]], pred_rel_labels, gt_classes[pred_rel_inds[:, 1]]))
        pr_tb = np.column_stack((gt_boxes[pred_rel_inds[:, 0]], gt_boxes[pred_rel_inds[:, 1]]))
        pred_to_gt = compute_pred_matches(gt_trip, pr_trip, gt_tb, pr_tb, self.iou_thres)
        hits = []
        for k in self.top_k:
            match = reduce(np.union1d, pred_to_gt[:k])
            self.recall[k].append(float(len(match)) / float(gt_rels.shape[0]))
            hits.append(len(match))
            recall_hit, recall_count = [0] * self.num_rel, [0] * self.num_rel
            for idx in range(gt_rels.shape[0]):
                recall_count[int(gt_rels[idx, 2])] += 1
                recall_count[0] += 1
            for idx in range(len(match)):
                recall_hit[int(gt_rels[int(match[idx]), 2])] += 1
                recall_hit[0] += 1
            for n in range(self.num_rel):
                if recall_count[n] > 0:
                    self.collect[k][n].append(float(recall_hit[n] / recall_count[n]))
        self.per_image.append((hits, int(gt_rels.shape[0])))

    def result(self):
        recall = {k: np.mean(v) for k, v in self.recall.items()}
        mean_recall, lists = {}, {}
        for k in self.top_k:
            s, lst = 0, []
            for idx in range(self.num_rel - 1):
                tmp = 0.0 if len(self.collect[k][idx + 1]) == 0 else np.mean(self.collect[k][idx + 1])
                lst.append(tmp)
                s += tmp
            mean_recall[k] = s / float(self.num_rel - 1)
            lists[k] = lst
        return dict(recall=recall, mean_recall=mean_recall, mean_recall_list=lists)
