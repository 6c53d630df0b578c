"""CPU ORACLE for the HIERCOM relation-prediction hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain torch-CPU / NumPy / Python loops, the algorithm of the reference
(bowen-upenn/scene_graph_commonsense @ 3388036f) for the path named in BASELINE.json `north_star`.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
it; the product package `scene_graph_commonsense_b200` never does (tests/test_no_oracle_in_product.py).

Parity pinning: the reference ships NO tests or golden vectors for this path, so the oracle
is pinned against the *imported, unmodified reference itself*: `oracle/make_golden.py` runs the reference
classes on seeded inputs in the build container and commits inputs+outputs under `tests/golden/`;
`tests/test_oracle_golden.py` replays them through this file.  The one deliberate normalisation is H1
(SURVEY §7): the reference's `torch.argsort(descending=True)` has an implementation-defined tie order; both the
golden generator (by patching `torch.argsort` to `stable=True`, reference source untouched) and this oracle
use (confidence desc, flat candidate index asc).

Each function cites the reference file:line it follows (paths relative to the reference root).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

TOP_K = (20, 50, 100)


# =====================================================================================================
# R5-R7: relation head (model.py)


def process_super_class(s_list, num_super_classes=17):
    """utils.py:136-149 - list of super-class ids per object -> int64 [B,17]: one_hot(s[0]) plus, for i in 1..3, one_hot(s[i])
    added to the rows with len(s) == i + 1 ONLY (`idx = nonzero([len(s) == i + 1 ...])`, :139-141): a list of 2..4 entries
    contributes its first and its LAST entry, the middle ones are never added; longer lists keep s[0] alone."""
    out = torch.zeros(len(s_list), num_super_classes, dtype=torch.int64)
    for r, s in enumerate(s_list):
        s = [int(v) for v in s]
        out[r, s[0]] += 1
        if 2 <= len(s) <= 4:
            out[r, s[-1]] += 1
    return out


def conv_layers(sd, h_sub, h_obj):
    """model.py:138-150 (eval mode: dropout is identity)."""
    a = torch.tanh(F.conv2d(h_sub, sd["conv1_1.weight"], sd["conv1_1.bias"]))
    b = torch.tanh(F.conv2d(h_obj, sd["conv1_2.weight"], sd["conv1_2.bias"]))
    h = torch.cat((a, b), dim=1)
    h = F.max_pool2d(F.relu(F.conv2d(h, sd["conv2_1.weight"], sd["conv2_1.bias"], padding=1)), 2, 2)
    h = F.max_pool2d(F.relu(F.conv2d(h, sd["conv3_1.weight"], sd["conv3_1.bias"], padding=1)), 2, 2)
    h = torch.reshape(h, (h.shape[0], -1))           # NCHW flatten: index = c*64 + y*8 + x
    return F.relu(F.linear(h, sd["fc1.weight"], sd["fc1.bias"]))


def concat_labels(h, c1, c2, s1, s2, num_classes=150, num_super_classes=17):
    """model.py:152-168."""
    parts = [h, F.one_hot(c1, num_classes).to(h.dtype), F.one_hot(c2, num_classes).to(h.dtype)]
    if s1 is not None:
        parts += [process_super_class(s1, num_super_classes).to(h.dtype),
                  process_super_class(s2, num_super_classes).to(h.dtype)]
    return torch.cat(parts, dim=1)


def hier_head(sd, pred, T=(1.0, 1.0, 1.0)):
    """model.py:176-184 - Bayes rule in log space: log p(rel, super) = log p(rel | super) + log p(super)."""
    connectivity = F.linear(pred, sd["fc4.weight"], sd["fc4.bias"])
    super_relation = F.log_softmax(F.linear(pred, sd["fc5.weight"], sd["fc5.bias"]), dim=1)
    rels = []
    for k in range(3):
        r = F.linear(pred, sd["fc3_%d.weight" % (k + 1)], sd["fc3_%d.bias" % (k + 1)])
        rels.append(F.log_softmax(r / T[k], dim=1) + super_relation[:, k].view(-1, 1))
    return rels[0], rels[1], rels[2], super_relation, connectivity


def bayesian_relation_classifier(sd, h_sub, h_obj, c1, c2, s1, s2, T=(1.0, 1.0, 1.0)):
    """model.py:170-186 BayesianRelationClassifier.forward (no augmentation branch)."""
    with torch.no_grad():
        h = conv_layers(sd, h_sub, h_obj)
        hc = concat_labels(h, c1, c2, s1, s2)
        pred = F.relu(F.linear(hc, sd["fc2.weight"], sd["fc2.bias"]))
        r1, r2, r3, sup, conn = hier_head(sd, pred, T)
    return r1, r2, r3, sup, conn, pred


def flat_relation_classifier(sd, h_sub, h_obj, c1, c2, s1, s2):
    """model.py:94-102 FlatRelationClassifier.forward."""
    with torch.no_grad():
        h = conv_layers(sd, h_sub, h_obj)
        hc = concat_labels(h, c1, c2, s1, s2)
        pred = F.relu(F.linear(hc, sd["fc2.weight"], sd["fc2.bias"]))
        relation = F.linear(pred, sd["fc3.weight"], sd["fc3.bias"])
        conn = F.linear(pred, sd["fc4.weight"], sd["fc4.bias"])
    return relation, conn, pred


def bayesian_head(sd, h, T=(1.0, 1.0, 1.0)):
    """model.py:24-34 BayesianHead.forward (head alone, 512 -> 15/11/24/3)."""
    with torch.no_grad():
        sup = F.log_softmax(F.linear(h, sd["fc5.weight"], sd["fc5.bias"]), dim=1)
        out = []
        for k in range(3):
            r = F.linear(h, sd["fc3_%d.weight" % (k + 1)], sd["fc3_%d.bias" % (k + 1)])
            out.append(F.log_softmax(r / T[k], dim=1) + sup[:, k].view(-1, 1))
    return out[0], out[1], out[2], sup


# =====================================================================================================
# R1-R4: masks, pair enumeration, overlap pre-filter (evaluate.py:111-183)


def _slice_bounds(lo, hi, size):
    """Python `a[int(lo):int(hi)]` on an axis of length `size`: negatives wrap once, then clamp."""
    lo, hi = int(lo), int(hi)
    if lo < 0:
        lo = max(lo + size, 0)
    if hi < 0:
        hi = max(hi + size, 0)
    lo, hi = min(lo, size), min(hi, size)
    return lo, max(hi, lo)


def box_mask(box, size=32):
    """evaluate.py:113-115 - mask[int(y0):int(y1), int(x0):int(x1)] = 1 with box = (xmin,xmax,ymin,ymax)."""
    m = np.zeros((size, size), dtype=np.uint8)
    m[int(box[2]):int(box[3]), int(box[0]):int(box[1])] = 1
    return m


def masks_overlap(box_a, box_b, size=32):
    """evaluate.py:150-154 - the `or/and` ratio with inf->0 and nan->False reduces to any(mask_a & mask_b)."""
    return bool(np.any(box_mask(box_a, size) & box_mask(box_b, size)))


def grid_iou_ge_half(box_t, box_p, size=32, thresh=0.5):
    """evaluator.py:84-94 - rasterised-mask IoU, returns iou >= thresh (0 when union == 0)."""
    mt, mp = box_mask(box_t, size), box_mask(box_p, size)
    inter = int(np.sum(mt & mp))
    union = int(np.sum(mt | mp))
    if union == 0:
        return 0 >= thresh
    return float(inter) / float(union) >= thresh


def compare_object_cat(pred_cat, target_cat):
    """utils.py:355-373."""
    equiv = [[1, 5, 11, 23, 38, 44, 121, 124, 148, 149], [0, 50], [92, 137]]
    unsymm = {123: [14, 63, 95, 87, 123], 108: [89, 102, 67, 72, 71, 81, 96, 105, 90, 111, 108],
              60: [145, 106, 142, 144, 77, 60]}
    pred_cat, target_cat = int(pred_cat), int(target_cat)
    if pred_cat == target_cat:
        return True
    for grp in equiv:
        if pred_cat in grp and target_cat in grp:
            return True
    for key, lst in unsymm.items():
        if pred_cat == key and target_cat in lst:
            return True
        if target_cat == key and pred_cat in lst:
            return True
    return False


# =====================================================================================================
# R8-R13: evaluators (evaluator.py)


def _pack(s, p, o):
    return (int(s) * 50 + int(p)) * 150 + int(o)


class OracleEvaluator:
    """evaluator.py:15-367 (hierarchical and flat) restated with per-call row storage.

    `aligned`/`violated` are python sets of packed keys (or None when run_mode is not *_cs, evaluator.py:76-81);
    `zero_shot` is a set of packed keys (evaluator.py:39)."""

    def __init__(self, splits=(15, 11, 24), hierar=True, aligned=None, violated=None, zero_shot=None,
                 top_k=TOP_K, num_classes=50, iou_thresh=0.5, feature_size=32):
        self.G, self.Pn, self.S = splits
        self.hierar = hierar
        self.aligned, self.violated = aligned, violated
        self.zero_shot = zero_shot if zero_shot is not None else set()
        self.top_k = tuple(top_k)
        self.num_classes = num_classes
        self.iou_thresh = iou_thresh
        self.fs = feature_size
        z = lambda: np.zeros(num_classes, dtype=np.float64)
        self.result_dict = {k: 0.0 for k in self.top_k}
        self.result_per_class = {k: z() for k in self.top_k}
        self.num_connected_target = 0.0
        self.num_conn_target_per_class = z()
        self.result_dict_zs = {k: 0.0 for k in self.top_k}
        self.result_per_class_zs = {k: z() for k in self.top_k}
        self.num_connected_target_zs = 0.0
        self.num_conn_target_per_class_zs = z()
        self.clear_data()

    def clear_data(self):                                   # evaluator.py:568-583
        self.which, self.conf, self.conn, self.rel = [], [], [], []
        self.cs, self.co, self.bs, self.bo = [], [], [], []
        self.t_which, self.t_rel, self.t_cs, self.t_co, self.t_bs, self.t_bo = [], [], [], [], [], []
        self.sgd_targets = None

    def accumulate(self, which_in_batch, relation_pred, relation_target, super_relation_pred, connectivity,
                   subject_cat_pred, object_cat_pred, subject_cat_target, object_cat_target,
                   subject_bbox_pred, object_bbox_pred, subject_bbox_target, object_bbox_target, iou_mask,
                   predcls=True, cat_subject_confidence=None, cat_object_confidence=None):
        """evaluator.py:118-269.  All arguments array-likes with leading dim bs."""
        rp = np.asarray(relation_pred, dtype=np.float32)
        which = np.asarray(which_in_batch, dtype=np.int64)
        conn = np.asarray(connectivity, dtype=np.float32)
        cs = np.asarray(subject_cat_pred, dtype=np.int64)
        co = np.asarray(object_cat_pred, dtype=np.int64)
        bs_ = np.asarray(subject_bbox_pred)
        bo_ = np.asarray(object_bbox_pred)
        iou_mask = np.asarray(iou_mask, dtype=bool)
        G, Pn = self.G, self.Pn
        if self.hierar:                                      # :157-179 / :231-251
            segs = [(0, G), (G, G + Pn), (G + Pn, rp.shape[1])]
            conf = np.concatenate([rp[:, a:b].max(axis=1) for a, b in segs]).astype(np.float32)
            rel = np.concatenate([rp[:, a:b].argmax(axis=1) + a for a, b in segs]).astype(np.int64)
            rep = 3
        else:                                                # :128-134 / :199-205
            conf = rp.max(axis=1).astype(np.float32)
            rel = rp.argmax(axis=1).astype(np.int64)
            rep = 1
        if not predcls:                                      # :164-166
            ins = (np.asarray(cat_subject_confidence, dtype=np.float32) +
                   np.asarray(cat_object_confidence, dtype=np.float32)).astype(np.float32)
            conf = (conf + np.tile(ins, rep)).astype(np.float32)
        conf[~np.tile(iou_mask, rep)] = -math.inf            # :167-168
        cs_r, co_r = np.tile(cs, rep), np.tile(co, rep)
        if self.aligned is not None:                         # :189-194 / :261-266
            for i in range(len(conf)):
                key = _pack(cs_r[i], rel[i], co_r[i])
                if key in self.violated or key not in self.aligned:
                    conf[i] = -math.inf
        self.which.append(np.tile(which, rep))
        self.conf.append(conf)
        self.conn.append(np.tile(conn, rep))
        self.rel.append(rel)
        self.cs.append(cs_r)
        self.co.append(co_r)
        self.bs.append(np.tile(bs_, (rep, 1)))
        self.bo.append(np.tile(bo_, (rep, 1)))
        if predcls:                                          # :181-187 / :253-259
            self.t_which.append(which)
            self.t_rel.append(np.asarray(relation_target, dtype=np.int64))
            self.t_cs.append(np.asarray(subject_cat_target, dtype=np.int64))
            self.t_co.append(np.asarray(object_cat_target, dtype=np.int64))
            self.t_bs.append(np.asarray(subject_bbox_target))
            self.t_bo.append(np.asarray(object_bbox_target))

    def accumulate_target(self, relation_target, subject_cat_target, object_cat_target,
                          subject_bbox_target, object_bbox_target):
        """evaluator.py:272-277 - SGDET/SGCLS: per-image lists (None for images without GT)."""
        self.sgd_targets = (relation_target, subject_cat_target, object_cat_target,
                            subject_bbox_target, object_bbox_target)

    def compute(self, per_class=True, predcls=True):
        """evaluator.py:280-367."""
        if not self.conf:
            return self.metrics()
        which = np.concatenate(self.which)
        conf = (np.concatenate(self.conf) + np.concatenate(self.conn)).astype(np.float32)   # :292
        rel = np.concatenate(self.rel)
        cs, co = np.concatenate(self.cs), np.concatenate(self.co)
        bs_, bo_ = np.concatenate(self.bs), np.concatenate(self.bo)
        if self.sgd_targets is None:
            t_which = np.concatenate(self.t_which)
            t_rel = np.concatenate(self.t_rel)
            t_cs, t_co = np.concatenate(self.t_cs), np.concatenate(self.t_co)
            t_bs, t_bo = np.concatenate(self.t_bs), np.concatenate(self.t_bo)
        for image in np.unique(which):                       # :294
            cur = np.nonzero(which == image)[0]
            if self.sgd_targets is None:
                tcur = np.nonzero(t_which == image)[0]
                g_rel, g_cs, g_co, g_bs, g_bo = t_rel[tcur], t_cs[tcur], t_co[tcur], t_bs[tcur], t_bo[tcur]
            else:
                if self.sgd_targets[0][int(image)] is None:  # :298-299
                    continue
                g_rel, g_cs, g_co, g_bs, g_bo = (np.asarray(x[int(image)]) for x in self.sgd_targets)
                g_bs, g_bo = g_bs.reshape(-1, 4), g_bo.reshape(-1, 4)
            order = np.argsort(-conf[cur].astype(np.float64), kind="stable")   # :304 with H1 tie order
            # (negating keeps -inf last and ties in index order; NaN never occurs on this path)
            keep = cur[order[:min(self.top_k[-1], len(cur))]]                  # :315-316
            for i in range(len(g_rel)):                      # :306
                t = int(g_rel[i])
                if t == -1:
                    continue
                zs = _pack(g_cs[i], t, g_co[i]) in self.zero_shot              # :310-311,341
                for j, c in enumerate(keep):                 # :319
                    if predcls:                              # :320-325
                        lab = (g_cs[i] == cs[c]) and (g_co[i] == co[c])
                    else:
                        lab = compare_object_cat(g_cs[i], cs[c]) and compare_object_cat(g_co[i], co[c])
                    if not lab:
                        continue
                    if not (grid_iou_ge_half(g_bs[i], bs_[c], self.fs, self.iou_thresh) and
                            grid_iou_ge_half(g_bo[i], bo_[c], self.fs, self.iou_thresh)):
                        continue
                    if t == rel[c]:                          # :331-348
                        for k in self.top_k:
                            if j >= k:
                                continue
                            self.result_dict[k] += 1.0
                            if per_class:
                                self.result_per_class[k][t] += 1.0
                            if zs:
                                self.result_dict_zs[k] += 1.0
                                if per_class:
                                    self.result_per_class_zs[k][t] += 1.0
                        break
                self.num_connected_target += 1.0             # :350-356
                self.num_conn_target_per_class[t] += 1.0
                if zs:
                    self.num_connected_target_zs += 1.0
                    self.num_conn_target_per_class_zs[t] += 1.0
        return self.metrics()

    def metrics(self):
        """evaluator.py:358-365 (float64 ratio for R@k, float32 nanmean for mR@k)."""
        return metrics_from_counts(self.result_dict, self.result_per_class, self.num_connected_target,
                                   self.num_conn_target_per_class, self.top_k) + \
            metrics_from_counts(self.result_dict_zs, self.result_per_class_zs, self.num_connected_target_zs,
                                self.num_conn_target_per_class_zs, self.top_k)

    def counters(self):
        """Flat int64 view in the product's counter layout (tables.EV_*), for bit-exact comparison."""
        def block(hits, hits_pc, n, n_pc):
            return np.concatenate([[hits[k] for k in self.top_k],
                                   np.concatenate([hits_pc[k] for k in self.top_k]), [n], n_pc])
        return np.concatenate([
            block(self.result_dict, self.result_per_class, self.num_connected_target, self.num_conn_target_per_class),
            block(self.result_dict_zs, self.result_per_class_zs, self.num_connected_target_zs,
                  self.num_conn_target_per_class_zs)]).astype(np.int64)


def metrics_from_counts(hits, hits_pc, n, n_pc, top_k=TOP_K):
    recall = [hits[k] / max(n, 1e-3) for k in top_k]
    per_class = [torch.as_tensor(hits_pc[k], dtype=torch.float32) / torch.as_tensor(n_pc, dtype=torch.float32)
                 for k in top_k]
    mean_recall = [torch.nanmean(r) for r in per_class]
    return recall, per_class, mean_recall


class OracleEvaluatorTop3:
    """evaluator.py:589-773 - R@k* / mR@k*: one candidate per directed pair (max over the three heads), a GT is
    found if ANY of the three per-head argmaxes equals it; plus the top-1-super variant (:746-760)."""

    def __init__(self, splits=(15, 11, 24), top_k=TOP_K, num_classes=50, iou_thresh=0.5, feature_size=32):
        self.G, self.Pn, self.S = splits
        self.top_k = tuple(top_k)
        self.iou_thresh = iou_thresh
        self.fs = feature_size
        z = lambda: np.zeros(num_classes, dtype=np.float64)
        self.result_dict = {k: 0.0 for k in self.top_k}
        self.result_dict_top1 = {k: 0.0 for k in self.top_k}
        self.result_per_class = {k: z() for k in self.top_k}
        self.result_per_class_top1 = {k: z() for k in self.top_k}
        self.num_connected_target = 0.0
        self.num_conn_target_per_class = z()
        self.clear_data()

    def clear_data(self):
        self.rows = []

    def accumulate(self, which_in_batch, relation_pred, relation_target, super_relation_pred, connectivity,
                   subject_cat_pred, object_cat_pred, subject_cat_target, object_cat_target,
                   subject_bbox_pred, object_bbox_pred, subject_bbox_target, object_bbox_target, iou_mask):
        """evaluator.py:639-685."""
        rp = np.asarray(relation_pred, dtype=np.float32)
        G, Pn = self.G, self.Pn
        conf = np.maximum(np.maximum(rp[:, :G].max(axis=1), rp[:, G:G + Pn].max(axis=1)),
                          rp[:, G + Pn:].max(axis=1)).astype(np.float32)       # :646-648
        conf[~np.asarray(iou_mask, dtype=bool)] = -math.inf                    # :649
        self.rows.append(dict(which=np.asarray(which_in_batch, dtype=np.int64), conf=conf,
                              conn=np.asarray(connectivity, dtype=np.float32), rp=rp,
                              sup=np.asarray(super_relation_pred, dtype=np.float32),
                              t=np.asarray(relation_target, dtype=np.int64),
                              cs=np.asarray(subject_cat_pred, dtype=np.int64), co=np.asarray(object_cat_pred, dtype=np.int64),
                              tcs=np.asarray(subject_cat_target, dtype=np.int64), tco=np.asarray(object_cat_target, dtype=np.int64),
                              bs=np.asarray(subject_bbox_pred), bo=np.asarray(object_bbox_pred),
                              tbs=np.asarray(subject_bbox_target), tbo=np.asarray(object_bbox_target)))

    def compute(self, per_class=True):
        """evaluator.py:697-773."""
        if not self.rows:
            return self.metrics()
        cat = lambda k: np.concatenate([r[k] for r in self.rows])
        which, rp, sup, t_all = cat("which"), cat("rp"), cat("sup"), cat("t")
        conf = (cat("conf") + cat("conn")).astype(np.float32)                  # :702
        cs, co, tcs, tco = cat("cs"), cat("co"), cat("tcs"), cat("tco")
        bs_, bo_, tbs, tbo = cat("bs"), cat("bo"), cat("tbs"), cat("tbo")
        G, Pn = self.G, self.Pn
        a1 = rp[:, :G].argmax(axis=1)
        a2 = rp[:, G:G + Pn].argmax(axis=1) + G
        a3 = rp[:, G + Pn:].argmax(axis=1) + G + Pn
        labs = np.stack((a1, a2, a3), axis=1)
        top_super = sup.argmax(axis=1)
        for image in np.unique(which):
            cur = np.nonzero(which == image)[0]
            order = np.argsort(-conf[cur].astype(np.float64), kind="stable")
            keep = cur[order[:min(self.top_k[-1], len(cur))]]
            num_target = int(np.sum(t_all[cur] != -1))                         # :716
            for gi in cur:                                                     # :711
                t = int(t_all[gi])
                if t == -1:
                    continue
                found = found_top1 = False
                for j, c in enumerate(keep):                                   # :722
                    if not (tcs[gi] == cs[c] and tco[gi] == co[c]):
                        continue
                    if not (grid_iou_ge_half(tbs[gi], bs_[c], self.fs, self.iou_thresh) and
                            grid_iou_ge_half(tbo[gi], bo_[c], self.fs, self.iou_thresh)):
                        continue
                    if not found and t in labs[c]:                             # :730-744
                        for k in self.top_k:
                            if j >= max(k, num_target):
                                continue
                            self.result_dict[k] += 1.0
                            if per_class:
                                self.result_per_class[k][t] += 1.0
                        found = True
                    if not found_top1 and t == labs[c][top_super[c]]:          # :746-760
                        for k in self.top_k:
                            if j >= max(k, num_target):
                                continue
                            self.result_dict_top1[k] += 1.0
                            if per_class:
                                self.result_per_class_top1[k][t] += 1.0
                        found_top1 = True
                    if found and found_top1:
                        break
                self.num_connected_target += 1.0                               # :765-766
                self.num_conn_target_per_class[t] += 1.0
        return self.metrics()

    def metrics(self):
        return metrics_from_counts(self.result_dict, self.result_per_class, self.num_connected_target,
                                   self.num_conn_target_per_class, self.top_k)

    def counters(self):
        """Flat int64 view in the product's layout (tables.T3_*)."""
        return np.concatenate([
            [self.result_dict[k] for k in self.top_k], np.concatenate([self.result_per_class[k] for k in self.top_k]),
            [self.result_dict_top1[k] for k in self.top_k], np.concatenate([self.result_per_class_top1[k] for k in self.top_k]),
            [self.num_connected_target], self.num_conn_target_per_class]).astype(np.int64)


# =====================================================================================================
# L1 driver loops restated (evaluate.py:111-217 PredCLS, :382-446 SGDET/SGCLS)


def _masked_input(sample, box):
    """evaluate.py:136-137 - h = cat(image_feature * mask, image_depth * mask) -> [257,32,32]."""
    m = torch.from_numpy(box_mask(box, sample.feat.shape[-1]).astype(np.float32))
    return torch.cat((sample.feat * m, sample.depth * m), dim=0)


def replay_predcls(batch, head_fn, evaluator, evaluator_top3=None, stats=None, features=True):
    """evaluate.py:111-183 for ONE batch (a list of ImageSample): lock-step (graph_iter, edge_iter) loops over the
    batch with the whole-batch skip rule (:155-156), two directed passes per surviving pair (:160-183 ->
    train_utils.py:160-196).  `head_fn(h_sub, h_obj, cat_sub, cat_obj, spcat_sub, spcat_obj)` returns
    `(relation [bs,50], super_relation [bs,3] or None, connectivity [bs,1])`.  Returns #directed pairs run."""
    n_obj = np.array([len(s.categories) for s in batch])
    ran = 0
    for g in range(int(n_obj.max())):
        keep = np.nonzero(n_obj > g)[0]
        for e in range(g):
            iou_mask = np.array([masks_overlap(batch[i].bbox[g], batch[i].bbox[e]) for i in keep])
            if iou_mask.sum() == 0:                          # :155-156
                continue
            hg = torch.stack([_masked_input(batch[i], batch[i].bbox[g]) for i in keep]) if features else None
            he = torch.stack([_masked_input(batch[i], batch[i].bbox[e]) for i in keep]) if features else None
            cg = torch.stack([batch[i].categories[g] for i in keep])
            ce = torch.stack([batch[i].categories[e] for i in keep])
            sg = [batch[i].super_categories[g] for i in keep]
            se = [batch[i].super_categories[e] for i in keep]
            bg = torch.stack([batch[i].bbox[g] for i in keep])
            be = torch.stack([batch[i].bbox[e] for i in keep])
            rel_t = torch.stack([batch[i].relationships[g - 1][e] for i in keep])
            dir_t = torch.stack([batch[i].subj_or_obj[g - 1][e] for i in keep])
            for first in (True, False):
                if first:
                    args = (hg, he, cg, ce, sg, se, bg, be)
                else:
                    args = (he, hg, ce, cg, se, sg, be, bg)
                relation, sup, conn = head_fn(*args[:6], (keep, g, e) if first else (keep, e, g))
                not_conn = dir_t != (1 if first else 0)      # train_utils.py:169-174
                t = rel_t.clone()
                t[not_conn] = -1                             # train_utils.py:186-187
                logsig = torch.log(torch.sigmoid(conn[:, 0]))    # train_utils.py:190
                evaluator.accumulate(keep, relation.numpy(), t.numpy(), None if sup is None else sup.numpy(),
                                     logsig.numpy(), args[2].numpy(), args[3].numpy(), args[2].numpy(), args[3].numpy(),
                                     args[6].numpy(), args[7].numpy(), args[6].numpy(), args[7].numpy(), iou_mask)
                if evaluator_top3 is not None:
                    evaluator_top3.accumulate(keep, relation.numpy(), t.numpy(), sup.numpy(), logsig.numpy(),
                                              args[2].numpy(), args[3].numpy(), args[2].numpy(), args[3].numpy(),
                                              args[6].numpy(), args[7].numpy(), args[6].numpy(), args[7].numpy(), iou_mask)
                if stats is not None:                        # train_utils.py:175-183
                    connected = ~not_conn
                    pred_conn = torch.sigmoid(conn[:, 0]) >= 0.5
                    stats["num_not_connected"] += int(not_conn.sum())
                    stats["num_connected"] += int(connected.sum())
                    stats["num_connected_pred"] += int(pred_conn.sum())
                    stats["connectivity_precision"] += int((rel_t[pred_conn] != -1).sum())
                    stats["connectivity_recall"] += int(torch.round(torch.sigmoid(conn[connected, 0])).sum())
                ran += len(keep)
    return ran


def match_target_sgd(batch):
    """utils.py:294-352 - flat per-image GT triplet lists in (g,e) loop order; None for images without GT.
    Reference quirk kept on purpose: the outer loop is `range(len(relationships[image]))` (utils.py:312) and
    `relationships` has N-1 rows, so g stops at N-2 and relations of the LAST box (g = N-1) never become targets."""
    out = ([], [], [], [], [])
    for s in batch:
        rel, cs, co, bs_, bo_ = [], [], [], [], []
        for g in range(1, len(s.relationships)):
            for e in range(g):
                d = float(s.subj_or_obj[g - 1][e])
                if d == 1:
                    a, b = g, e
                elif d == 0:
                    a, b = e, g
                else:
                    continue
                rel.append(int(s.relationships[g - 1][e]))
                cs.append(int(s.categories[a]))
                co.append(int(s.categories[b]))
                bs_.append(s.bbox[a].numpy())
                bo_.append(s.bbox[b].numpy())
        if rel:
            for lst, v in zip(out, (np.array(rel), np.array(cs), np.array(co), np.stack(bs_), np.stack(bo_))):
                lst.append(v)
        else:
            for lst in out:
                lst.append(None)
    return out


def replay_sgdet(batch, head_fn, evaluator, features=True):
    """evaluate.py:382-446 for one batch whose proposals are already prepared (bbox_pred, categories_pred,
    cat_conf_pred, super_categories_pred on each sample): pair loop over *predicted* boxes, object-confidence add,
    flat GT list via match_target_sgd, `compute(predcls=False)` by the caller."""
    n_obj = np.array([len(s.categories_pred) for s in batch])
    ran = 0
    for g in range(int(n_obj.max())):
        keep = np.nonzero(n_obj > g)[0]
        for e in range(g):
            iou_mask = np.array([masks_overlap(batch[i].bbox_pred[g], batch[i].bbox_pred[e]) for i in keep])
            if iou_mask.sum() == 0:                          # :407-408
                continue
            hg = torch.stack([_masked_input(batch[i], batch[i].bbox_pred[g]) for i in keep]) if features else None
            he = torch.stack([_masked_input(batch[i], batch[i].bbox_pred[e]) for i in keep]) if features else None
            cg = torch.stack([batch[i].categories_pred[g] for i in keep])
            ce = torch.stack([batch[i].categories_pred[e] for i in keep])
            sg = [batch[i].super_categories_pred[g] for i in keep]
            se = [batch[i].super_categories_pred[e] for i in keep]
            bg = torch.stack([batch[i].bbox_pred[g] for i in keep])
            be = torch.stack([batch[i].bbox_pred[e] for i in keep])
            fg = torch.stack([batch[i].cat_conf_pred[g] for i in keep])
            fe = torch.stack([batch[i].cat_conf_pred[e] for i in keep])
            for first in (True, False):
                a = (hg, he, cg, ce, sg, se, bg, be, fg, fe) if first else (he, hg, ce, cg, se, sg, be, bg, fe, fg)
                relation, sup, conn = head_fn(*a[:6], (keep, g, e) if first else (keep, e, g))
                logsig = torch.log(torch.sigmoid(conn[:, 0]))
                evaluator.accumulate(keep, relation.numpy(), None, None if sup is None else sup.numpy(), logsig.numpy(),
                                     a[2].numpy(), a[3].numpy(), None, None, a[6].numpy(), a[7].numpy(), None, None,
                                     iou_mask, False, a[8].numpy(), a[9].numpy())
                ran += len(keep)
    evaluator.accumulate_target(*match_target_sgd(batch))    # :446
    return ran


def make_head_fn(sd, hierar=True):
    def fn(h_sub, h_obj, c1, c2, s1, s2, ctx=None):
        if hierar:
            r1, r2, r3, sup, conn, _ = bayesian_relation_classifier(sd, h_sub, h_obj, c1, c2, s1, s2)
            return torch.cat((r1, r2, r3), dim=1), sup, conn
        rel, conn, _ = flat_relation_classifier(sd, h_sub, h_obj, c1, c2, s1, s2)
        return rel, None, conn
    return fn
