"""Plug-and-play twins for the Scene-Graph-Benchmark host (BASELINE config 5: PredCLS Motifs,
4096-d union features, 51 classes).  Citations: SGB = scenegraph_benchmark/Scene-Graph-Benchmark.pytorch/maskrcnn_benchmark.

  BayesHead / BayesHeadProb      SGB/modeling/roi_heads/relation_head/model_motifs_hierarchical.py:8-72
  hierarchical_relation_tail     SGB/.../roi_relation_predictors.py:400-469 (everything after the context encoder)
  HierarchPostProcessor          SGB/.../inference.py:147-312 (validator = the LLM call, injected; default: accept all)
  SGBRecall                      SGB/data/datasets/evaluation/vg/sgg_eval.py:41-99 (SGRecall), :316-385 (SGMeanRecall),
                                 :494-565 (_triplet, _compute_pred_matches), structures/boxlist_ops.py:54-90 (IoU, +1)

Dense work (post_cat 1024->4096 with the `* union_features` epilogue, BayesHead 4096->54) runs on the tcgen05 kernel;
gather / frequency bias / hierarchical log-softmax / candidates / ranking window / matching are the kernels of csrc/sgb.cu.
The LSTM/Tree/Transformer context encoders, ROI feature extractors and the detector stay SGB code.
"""
import os

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._lib import EPI_BF16, EPI_F32, EPI_SPLIT3_BF16

GEO_LABEL = [1, 2, 3, 4, 5, 6, 8, 10, 22, 23, 29, 31, 32, 33, 43]                  # roi_relation_predictors.py:376
POS_LABEL = [9, 16, 17, 20, 27, 30, 36, 42, 48, 49, 50]                             # :377
SEM_LABEL = [7, 11, 12, 13, 14, 15, 18, 19, 21, 24, 25, 26, 28, 34, 35, 37, 38, 39, 40, 41, 44, 45, 46, 47]   # :378-379
LABEL_IDS = GEO_LABEL + POS_LABEL + SEM_LABEL
SPLITS = (15, 11, 24)
NUM_OBJ_SGB = 151
NUM_REL_SGB = 51


_LABEL_IDS_DEV = {}


def _label_ids(device):
    key = str(device)
    if key not in _LABEL_IDS_DEV:
        _LABEL_IDS_DEV[key] = torch.tensor(LABEL_IDS, dtype=torch.int32, device=device)
    return _LABEL_IDS_DEV[key]


class SimpleBoxList:
    """The slice of maskrcnn_benchmark.structures.bounding_box.BoxList the relation path touches: `.bbox` (xyxy f32),
    `.size`, `get_field / add_field`, `len`.  Real BoxLists work unchanged (duck typing)."""

    def __init__(self, bbox, size=(0, 0), mode="xyxy"):
        self.bbox = torch.as_tensor(bbox, dtype=torch.float32)
        self.size, self.mode, self.extra_fields = size, mode, {}

    def add_field(self, k, v):
        self.extra_fields[k] = v

    def get_field(self, k):
        return self.extra_fields[k]

    def has_field(self, k):
        return k in self.extra_fields

    def __len__(self):
        return self.bbox.shape[0]


class BayesHead(nn.Module):
    """model_motifs_hierarchical.py:8-39 - four Linear layers returning RAW logits (super has a background slot)."""

    def __init__(self, input_dim=512, num_geometric=15, num_possessive=11, num_semantic=24, T1=1, T2=1, T3=1):
        super().__init__()
        self.fc3_1 = nn.Linear(input_dim, num_geometric)
        self.fc3_2 = nn.Linear(input_dim, num_possessive)
        self.fc3_3 = nn.Linear(input_dim, num_semantic)
        self.fc5 = nn.Linear(input_dim, 4)
        self.T1, self.T2, self.T3 = T1, T2, T3
        self._packed, self._versions = None, None

    def layer_init(self):
        for m in (self.fc3_1, self.fc3_2, self.fc3_3, self.fc5):
            nn.init.xavier_normal_(m.weight)                               # utils_relation.layer_init(xavier=True)
            nn.init.zeros_(m.bias)

    def splits(self):
        return (self.fc3_1.out_features, self.fc3_2.out_features, self.fc3_3.out_features)

    def packed(self):
        versions = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._packed is None or versions != self._versions:
            w = torch.cat((self.fc3_1.weight, self.fc3_2.weight, self.fc3_3.weight, self.fc5.weight)).detach().float()
            b = torch.cat((self.fc3_1.bias, self.fc3_2.bias, self.fc3_3.bias, self.fc5.bias)).detach().float()
            wp = torch.zeros(128, w.shape[1], device=w.device)
            wp[:w.shape[0]] = w
            bp = torch.zeros(128, device=w.device)
            bp[:b.shape[0]] = b
            self._packed = (ops.pack_weight_bf16x3(wp), bp.contiguous())
            self._versions = versions
        return self._packed

    def packed_bf16(self, dtype=torch.bfloat16):
        """Plain 16-bit [128, input_dim] packing (precision "bf16" / "fp16": one MMA per product instead of three)."""
        versions = tuple((p.data_ptr(), p._version) for p in self.parameters()) + (dtype,)
        if getattr(self, "_packed1", None) is None or versions != self._versions1:
            w = torch.cat((self.fc3_1.weight, self.fc3_2.weight, self.fc3_3.weight, self.fc5.weight)).detach().float()
            b = torch.cat((self.fc3_1.bias, self.fc3_2.bias, self.fc3_3.bias, self.fc5.bias)).detach().float()
            wp = torch.zeros(128, w.shape[1], device=w.device)
            wp[:w.shape[0]] = w
            bp = torch.zeros(128, device=w.device)
            bp[:b.shape[0]] = b
            self._packed1 = (wp.to(dtype).contiguous(), bp.contiguous())
            self._versions1 = versions
        return self._packed1

    def logits_from_bf16(self, a):
        """bf16 or fp16 [n, input_dim] -> f32 [n,128] logits with plain 16-bit operands of that format (fp32 accumulate)."""
        w, b = self.packed_bf16(a.dtype)
        n, k = a.shape
        out = torch.empty(n, 128, dtype=torch.float32, device=a.device)
        ops.tc_gemm(a, w, out, n, 128, k, bias=b, lda=k, ldc=128, epilogue=EPI_F32, group_m=8, tag="bayes_head")
        return out

    def logits_from_split(self, a):
        """bf16 [n, 3*input_dim] in the bf16x3 layout (hc_split_bf16x3 / HC_EPI_SPLIT3_BF16) -> f32 [n,128] logits."""
        w, b = self.packed()
        n, k3 = a.shape
        out = torch.empty(n, 128, dtype=torch.float32, device=a.device)
        ops.tc_gemm(a, w, out, n, 128, k3, bias=b, lda=k3, ldc=128, epilogue=EPI_F32, group_m=8, tag="bayes_head")
        return out

    def logits(self, h):
        """[n, input_dim] f32 -> f32 [n,128] (columns: heads then the 4 super logits, rest zero padding).
        bf16x3 split operands on the bf16 tensor cores: ~fp32 accuracy for the 4096-long dot products."""
        w, b = self.packed()
        a = ops.split_bf16x3(h)
        n, k3 = a.shape
        out = torch.empty(n, 128, dtype=torch.float32, device=h.device)
        ops.tc_gemm(a, w, out, n, 128, k3, bias=b, lda=k3, ldc=128, epilogue=EPI_F32, group_m=8, tag="bayes_head")
        return out

    @torch.no_grad()
    def forward(self, h):
        z = self.logits(h.float().contiguous())
        g, p, s = self.splits()
        return z[:, :g], z[:, g:g + p], z[:, g + p:g + p + s], z[:, g + p + s:g + p + s + 4]


class BayesHeadProb(BayesHead):
    """model_motifs_hierarchical.py:42-72 - same layers, log-softmax + Bayes add with super indices 1..3."""

    @torch.no_grad()
    def forward(self, h):
        z = self.logits(h.float().contiguous())
        rel, sup = ops.sgb_hier_softmax(z, self.splits())
        g, p, _ = self.splits()
        return rel[:, :g], rel[:, g:g + p], rel[:, g + p:], sup


_PACKED_LINEAR = {}
# Operand precision of the two SGB GEMMs (post_cat 1024 -> 4096, BayesHead 4096 -> 54): "bf16x3" = split operands
# (A_hi*W_hi + A_lo*W_hi + A_hi*W_lo, ~fp32 accuracy, 3x the MMAs), "bf16" = plain bf16 in / fp32 accumulate (north_star's nominal
# format; misses the 2e-3 probability bar on this tail: 8.3e-3 at 64 x 40), or "fp16" = plain IEEE half operands (same tensor-core
# rate as bf16, 8x smaller operand rounding: holds the bar with one MMA per product; stores saturate at +-65504).
PRECISIONS = ("bf16x3", "bf16", "fp16")
DEFAULT_PRECISION = "bf16x3"


def _packed_linear(lin, precision="bf16x3"):
    """B-side packing of an nn.Linear (bf16x3 split or plain bf16), cached until its parameters change (data_ptr / version)."""
    key = (id(lin), precision)
    ver = (lin.weight.data_ptr(), lin.weight._version, lin.bias.data_ptr(), lin.bias._version)
    hit = _PACKED_LINEAR.get(key)
    if hit is None or hit[0] != ver:
        if precision == "bf16x3":
            w = ops.pack_weight_bf16x3(lin.weight)
        else:
            if precision == "fp16" and float(lin.weight.detach().abs().max()) > 6.0e4:
                raise RuntimeError("hiercom_b200: a weight exceeds the fp16 range - use precision='bf16x3'")
            w = lin.weight.detach().to(torch.float16 if precision == "fp16" else torch.bfloat16).contiguous()
        hit = (ver, w, lin.bias.detach().float().contiguous())
        _PACKED_LINEAR[key] = hit
    return hit[1], hit[2]


class PerImage(tuple):
    """Per-image views (what the SGB interfaces exchange) that remember the whole-batch tensor they are slices of, so the
    next stage of this package takes `.whole` instead of re-concatenating 64 views."""
    whole = None


def _per_image(whole, counts):
    out = PerImage(whole.split(counts, 0))
    out.whole = whole
    return out


def _whole(parts):
    w = getattr(parts, "whole", None)
    return w if w is not None else torch.cat(list(parts))


_PAIR_INDEX_CACHE = {}


def global_pair_index(rel_pair_idxs, num_objs, device):
    """list of [P_i,2] per-image index tensors -> (int32 [P,2] global ids, int32 [B+1] pair offsets, int32 [P] image id,
    per-image pair counts).  A handful of whole-batch device ops (no per-image launches); the last result is cached on the
    identity + version of the index tensors because the relation tail and the post-processor receive the same list."""
    key = (tuple((p.data_ptr(), p._version, p.shape[0]) for p in rel_pair_idxs), tuple(int(n) for n in num_objs), str(device))
    hit = _PAIR_INDEX_CACHE.get("last")
    if hit is not None and hit[0] == key:
        return hit[1]
    num_rels = [int(p.shape[0]) for p in rel_pair_idxs]
    n = int(sum(num_rels))
    obj_base = np.concatenate(([0], np.cumsum(num_objs)))[:-1]
    host = torch.from_numpy(np.concatenate((np.concatenate(([0], np.cumsum(num_rels))), obj_base, num_rels)).astype(np.int32))
    dev_meta = host.to(device)
    b = len(num_rels)
    pair_off, base, counts = dev_meta[:b + 1], dev_meta[b + 1:2 * b + 1], dev_meta[2 * b + 1:]
    pair_img = torch.repeat_interleave(torch.arange(b, dtype=torch.int32, device=device), counts.long(), output_size=n)
    idx = torch.cat([p for p in rel_pair_idxs]).to(device=device, dtype=torch.int32)
    idx = (idx + base[pair_img.long()].unsqueeze(1)).contiguous()
    out = (idx, pair_off.contiguous(), pair_img, num_rels)
    _PAIR_INDEX_CACHE["last"] = (key, out, rel_pair_idxs)          # holds the list so the data_ptr identity stays valid
    return out


@torch.no_grad()
def hierarchical_relation_tail(edge_rep, rel_pair_idxs, num_objs, obj_preds, union_features, post_cat, rel_compress,
                               freq_bias_weight=None, use_vision=True, precision=None, ctx_compress=None):
    """roi_relation_predictors.py:400-469 after `edge_rep = self.post_emb(edge_ctx)`:
    pair gather -> post_cat -> * union_features -> BayesHead -> frequency bias -> hierarchical log-softmax.
    edge_rep f32 [sum N, 2*hidden]; obj_preds int [sum N]; union_features f32 [P, pooling_dim] (pooling_dim == MLP_HEAD_DIM,
    i.e. no `up_dim`, as in config 5); freq_bias_weight = FrequencyBias.obj_baseline.weight [151*151, 51] or None.
    ctx_compress (TransformerHierPredictor, roi_relation_predictors.py:233-238): a second BayesHead over the un-gated pair
    representation whose logits are ADDED to rel_compress's before the softmax.
    Returns (relation1_dist, relation2_dist, relation3_dist, superrelation_dist) split per image."""
    dev = edge_rep.device
    hidden = edge_rep.shape[1] // 2
    pair_idx, pair_off, pair_img, num_rels = global_pair_index(rel_pair_idxs, num_objs, dev)
    n = pair_idx.shape[0]
    precision = precision or DEFAULT_PRECISION
    if precision not in PRECISIONS:
        raise ValueError("precision must be one of %s" % (PRECISIONS,))
    pooling = post_cat.out_features
    if use_vision and union_features.shape[1] != pooling:
        raise NotImplementedError("union_single_not_match (up_dim) is not on the config-5 path")
    w, bias = _packed_linear(post_cat, precision)
    mul = union_features.float().contiguous() if use_vision else None
    if precision == "bf16x3":
        prod = ops.sgb_pair_gather(edge_rep.float().contiguous(), pair_idx, hidden, split=True)   # [P, 3*2*hidden] bf16x3 layout
        # post_cat GEMM whose epilogue applies `* union_features` and writes the bf16x3 A operand of the BayesHead GEMM directly:
        # the f32 [P, 4096] product never goes to HBM (was: write 16 KB/pair, read it back, split, write 24 KB/pair)
        prod3 = torch.empty(n, 3 * pooling, dtype=torch.bfloat16, device=dev)
        ops.tc_gemm(prod, w, prod3, n, pooling, 6 * hidden, bias=bias, lda=6 * hidden, ldc=3 * pooling, epilogue=EPI_SPLIT3_BF16,
                    mul=mul, group_m=16, m_sub=1, tag="post_cat")
        logits = rel_compress.logits_from_split(prod3)
    else:
        f16 = precision == "fp16"
        prod = ops.sgb_pair_gather(edge_rep.float().contiguous(), pair_idx, hidden, split=False, f16=f16)  # [P, 2*hidden] bf16 / fp16
        prod1 = torch.empty(n, pooling, dtype=torch.float16 if f16 else torch.bfloat16, device=dev)
        # CTA pairs (pooling % 256 == 0): eight epilogue warps per CTA for the `* union_features` epilogue, which bounds this GEMM
        pairs = int(os.environ.get("HC_SGB_PAIRS", "1")) if pooling % 256 == 0 else 0
        ops.tc_gemm(prod, w, prod1, n, pooling, 2 * hidden, bias=bias, lda=2 * hidden, ldc=pooling, epilogue=EPI_BF16, mul=mul,
                    group_m=16, m_sub=1, tag="post_cat", cta_pairs=pairs)
        logits = rel_compress.logits_from_bf16(prod1)
    if ctx_compress is not None:
        er = edge_rep.float()
        li = pair_idx.long()
        prod_rep = torch.cat((er[:, :hidden][li[:, 0]], er[:, hidden:][li[:, 1]]), dim=1).contiguous()     # :216-221, f32 [P, 2*hidden]
        logits = logits + ctx_compress.logits(prod_rep)
    pair_pred = obj_preds.to(dev, torch.int32)[pair_idx.long()].contiguous() if freq_bias_weight is not None else None
    rel, sup = ops.sgb_hier_softmax(logits, rel_compress.splits(), None if freq_bias_weight is None else freq_bias_weight.float().contiguous(),
                                    NUM_OBJ_SGB, pair_pred, _label_ids(dev))
    g, p, _ = rel_compress.splits()
    out = (_per_image(rel[:, :g], num_rels), _per_image(rel[:, g:g + p], num_rels), _per_image(rel[:, g + p:], num_rels),
           _per_image(sup, num_rels))
    out[0].joint = rel                                              # [P, G+P+S] as one tensor for HierarchPostProcessor.candidates
    return out


class HierarchPostProcessor(nn.Module):
    """inference.py:147-312 for `use_gt_box=True` (PredCLS / SGCLS).  `validator(combined_obj_label [k,2], rel_labels [k],
    image, boxlist) -> tensor of +1/-1` stands in for `CommonsenseValidator.query` (the LLM call stays reference code);
    `None` accepts everything.  Sorting is stable (score desc, index asc) - the reference's torch.sort leaves tie order
    unspecified.  The full-length ordering uses torch.sort on the device (library call); the evaluation path
    (`SGBRecall`) only needs the ranked window and uses the selection kernel instead."""

    def __init__(self, attribute_on=False, use_gt_box=False, later_nms_pred_thres=0.3, validator=None, llm_top_k=10, skip_top=10):
        super().__init__()
        self.attribute_on, self.use_gt_box, self.later_nms_pred_thres = attribute_on, use_gt_box, later_nms_pred_thres
        self.validator, self.llm_top_k, self.skip_top = validator, llm_top_k, skip_top

    @torch.no_grad()
    def candidates(self, rel1, rel2, rel3, refine_logits, rel_pair_idxs):
        """Object scores/labels (inference.py:214-222) and the 3P triple-score candidates (:246-281) for a whole batch."""
        dev = rel1[0].device
        num_objs = [int(l.shape[0]) for l in refine_logits]
        logit = _whole(refine_logits).float()
        prob = torch.softmax(logit, -1)
        prob[:, 0] = 0
        obj_scores, obj_pred = prob[:, 1:].max(dim=1)
        obj_pred = obj_pred + 1
        pair_idx, pair_off, pair_img, num_rels = global_pair_index(rel_pair_idxs, num_objs, dev)
        rel = getattr(rel1, "joint", None)                         # the tail's own [P,50] tensor when rel1..3 come from it
        if rel is None or rel.shape[1] != sum(SPLITS):
            rel = torch.cat((_whole(rel1), _whole(rel2), _whole(rel3)), dim=1)
        rel = rel.float().contiguous()
        score, label, row = ops.sgb_candidates(rel, SPLITS, pair_off, pair_img, pair_idx, obj_scores.contiguous(), _label_ids(dev))
        return dict(score=score, label=label, row=row, rel=rel, pair_idx=pair_idx, pair_off=pair_off, pair_img=pair_img, num_rels=num_rels,
                    num_objs=num_objs, obj_scores=obj_scores, obj_pred=obj_pred)

    @torch.no_grad()
    def forward(self, x, rel_pair_idxs, boxes, images=None):
        if not self.use_gt_box:
            raise NotImplementedError("sgdet post-processing (late NMS, box regression) stays SGB code (inference.py:219-242)")
        rel1, rel2, rel3, _super, refine_logits = x
        c = self.candidates(rel1, rel2, rel3, refine_logits, rel_pair_idxs)
        prob = torch.exp(c["rel"])
        obj_off = np.concatenate(([0], np.cumsum(c["num_objs"])))
        pair_off = c["pair_off"].cpu().numpy()
        results = []
        for i, box in enumerate(boxes):
            p0, p1 = int(pair_off[i]), int(pair_off[i + 1])
            o0, o1 = int(obj_off[i]), int(obj_off[i + 1])
            obj_class, obj_scores = c["obj_pred"][o0:o1], c["obj_scores"][o0:o1]
            box.add_field('pred_labels', obj_class)
            box.add_field('pred_scores', obj_scores)
            scores = c["score"][3 * p0:3 * p1].clone()
            labels = c["label"][3 * p0:3 * p1].long()
            rows = (c["row"][3 * p0:3 * p1] - p0).long()
            pair_local = (c["pair_idx"][p0:p1] - o0).long()
            triple_scores, sorting_idx = torch.sort(scores, dim=0, descending=True, stable=True)          # :282
            rel_pair_idx = pair_local[rows][sorting_idx]
            rel_class_prob = prob[p0:p1][rows][sorting_idx]
            rel_labels = labels[sorting_idx]
            if self.validator is not None:                                                                # :292-302
                a, b = self.skip_top, self.skip_top + self.llm_top_k
                combined = torch.stack((obj_class[rel_pair_idx[a:b, 0]], obj_class[rel_pair_idx[a:b, 1]]), dim=1)
                resp = torch.as_tensor(self.validator(combined, rel_labels[a:b], None if images is None else images[i], box),
                                       device=triple_scores.device)
                window = triple_scores[a:b]
                window[resp == -1] = float("-inf")
                _, sorting_idx2 = torch.sort(triple_scores, dim=0, descending=True, stable=True)
                rel_pair_idx = rel_pair_idx[sorting_idx2]
                rel_labels = rel_labels[sorting_idx2]
            box.add_field('rel_pair_idxs', rel_pair_idx)
            box.add_field('pred_rel_scores', rel_class_prob)       # NOT re-sorted by the second sort (reference behaviour)
            box.add_field('pred_rel_labels', rel_labels)
            results.append(box)
        return results


class SGBRecall:
    """SGRecall + SGMeanRecall (sgg_eval.py:41-99, 316-385) for PredCLS: R@K is the MEAN over images of per-image recall;
    mR@K the mean over the 50 predicates of the mean-over-images per-predicate recall.  The kernels return per-image
    integer hit / GT counts; the float reductions below follow the reference's NumPy operations, so results are identical
    for identical integers (and integers can be all-gathered across ranks)."""

    def __init__(self, top_k=(20, 50, 100), iou_thresh=0.5, num_rel=NUM_REL_SGB):
        self.top_k, self.iou_thresh, self.num_rel = tuple(top_k), iou_thresh, num_rel
        self.img_hits, self.img_ngt, self.img_hits_pc, self.img_cnt_pc = [], [], [], []

    @torch.no_grad()
    def evaluate_batch(self, cand, gt_rels, gt_classes, gt_boxes, reject=None):
        """cand: dict from HierarchPostProcessor.candidates; gt_rels: list of int [G_i,3] (sub id, obj id, label) with ids
        local to the image; gt_classes: list of int [N_i]; gt_boxes: list of f32 [N_i,4] xyxy.  PredCLS: predictions use the
        GT boxes and labels (vg_eval.py:267-270).  reject: optional uint8 [B,128] validator verdict per first-sort rank."""
        dev = cand["score"].device
        obj_base = np.concatenate(([0], np.cumsum(cand["num_objs"])))[:-1]
        g_cnt = np.asarray([int(g.shape[0]) for g in gt_rels], dtype=np.int64)
        g_off = np.concatenate(([0], np.cumsum(g_cnt))).astype(np.int32)
        n_gt = int(g_off[-1])
        # whole-batch packing: one cat per field, the per-image object-id base added by a gather (no per-image launches)
        if n_gt:
            rel_all = torch.cat([torch.as_tensor(g).view(-1, 3) for g in gt_rels])
            add = torch.from_numpy(np.repeat(obj_base, g_cnt).astype(np.int64)).to(rel_all.device, rel_all.dtype)
            rel_all = rel_all.clone()
            rel_all[:, :2] += add.unsqueeze(1)
            gt_rel = rel_all.to(dev, torch.int32).contiguous()
        else:
            gt_rel = torch.zeros(1, 3, dtype=torch.int32, device=dev)
        gt_cls = torch.cat([torch.as_tensor(c) for c in gt_classes]).to(dev, torch.int32).contiguous()
        gt_box = torch.cat([torch.as_tensor(b).view(-1, 4) for b in gt_boxes]).to(dev, torch.float32).contiguous()
        cand_off = (cand["pair_off"] * 3).contiguous()
        ranked = ops.topk_select(cand_off, cand["score"], 128)
        out = ops.sgb_rank_match(ranked, reject, cand["pair_off"], cand["score"], cand["label"], cand["row"], cand["pair_idx"], gt_cls, gt_box,
                                 torch.from_numpy(g_off).to(dev), gt_rel, gt_cls, gt_box, self.iou_thresh, self.top_k)
        # one D2H read for the five int32 result tables
        sizes = [t.numel() for t in out]
        flat = torch.cat([t.reshape(-1) for t in out]).cpu().numpy()
        cuts = np.cumsum(sizes)[:-1]
        final_rank, hits, ngt, hits_pc, cnt_pc = (a.reshape(t.shape) for a, t in zip(np.split(flat, cuts), out))
        self.img_hits.append(hits); self.img_ngt.append(ngt); self.img_hits_pc.append(hits_pc); self.img_cnt_pc.append(cnt_pc)
        return final_rank, ranked

    @staticmethod
    def reject_mask(cand, ranked, validator, skip_top=10, llm_top_k=10):
        """Validator verdicts for first-sort ranks [skip_top, skip_top+llm_top_k) of every image (inference.py:292-297) as the
        uint8 [B,128] mask `hc_sgb_rank_match` consumes.  The validator (LLM) runs on the host by nature."""
        dev = cand["score"].device
        n_img = ranked.shape[0]
        rej = torch.zeros(n_img, ranked.shape[1], dtype=torch.uint8, device=dev)
        pair_off = cand["pair_off"].cpu().numpy()
        for i in range(n_img):
            ids = ranked[i, skip_top:skip_top + llm_top_k]
            ids = ids[ids >= 0].long()
            if ids.numel() == 0:
                continue
            gc = ids + 3 * int(pair_off[i])
            rows = cand["row"][gc].long()
            so = cand["pair_idx"][rows].long()
            combined = torch.stack((cand["obj_pred"][so[:, 0]], cand["obj_pred"][so[:, 1]]), dim=1)
            resp = torch.as_tensor(validator(combined.cpu(), cand["label"][gc].cpu().long(), None, None))
            rej[i, skip_top:skip_top + ids.numel()] = (resp == -1).to(torch.uint8).to(dev)
        return rej

    def per_image_arrays(self):
        return (np.concatenate(self.img_hits), np.concatenate(self.img_ngt), np.concatenate(self.img_hits_pc), np.concatenate(self.img_cnt_pc))

    def result(self):
        return recall_from_image_counts(*self.per_image_arrays(), top_k=self.top_k, num_rel=self.num_rel)


def recall_from_image_counts(hits, ngt, hits_pc, cnt_pc, top_k=(20, 50, 100), num_rel=NUM_REL_SGB):
    """sgg_eval.py:95-97,51 (R@K = np.mean of per-image float ratios; images without GT are skipped, vg_eval.py:241-242)
    and :368-384 (mR@K)."""
    keep = ngt > 0
    recall, mean_recall, mean_recall_list = {}, {}, {}
    for qi, k in enumerate(top_k):
        vals = [float(h) / float(n) for h, n in zip(hits[keep, qi], ngt[keep])]
        recall[k] = np.mean(vals) if vals else float("nan")
        lst, s = [], 0
        for n in range(1, num_rel):
            col = [float(h / c) for h, c in zip(hits_pc[keep, qi, n], cnt_pc[keep, n]) if c > 0]
            tmp = 0.0 if len(col) == 0 else np.mean(col)
            lst.append(tmp)
            s += tmp
        mean_recall[k] = s / float(num_rel - 1)
        mean_recall_list[k] = lst
    return dict(recall=recall, mean_recall=mean_recall, mean_recall_list=mean_recall_list)
