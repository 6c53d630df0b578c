"""Drop-in twins of the reference's `Evaluator` and `Evaluator_Top3` (reference evaluator.py:15-367, 589-790).

Same constructor, `accumulate(...)`, `accumulate_target(...)`, `compute(...)`, `clear_data()` signatures and return
tuples, same public counters (`result_dict`, `result_per_class`, `num_connected_target`,
`num_conn_target_per_class` and the `_zs` / `_top1` twins, accumulated across `compute()` calls and never reset by
`clear_data`, evaluator.py:568-583).  `accumulate` only records device tensors; `compute` runs three kernels
(candidates -> per-image top-K + first-match scan + counters) instead of the reference's Python triple loop.

Tie order: (confidence desc, append order asc) == `torch.argsort(stable=True)`; the reference's unstable sort leaves
the order of exactly-equal confidences (typically the -inf candidates) implementation-defined.
Out of scope here (LLM querying, OIv6 precision, visualisation dumps): `get_related_top_k_predictions*`,
`compute_precision`, `save_visualization_results` raise NotImplementedError pointing at the reference.
"""
import os
import warnings

import numpy as np
import torch

from . import ops, tables


def _dev(device=None):
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise RuntimeError("hiercom_b200: the evaluator kernels need a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


_WARNED = set()


def _load_keys(path, fallback):
    """Triplet set at `path` (the reference resolves these relative to its working directory, evaluator.py:37-39,77-81), else the
    packed copy of the reference's shipped file under scene_graph_commonsense_b200/data/.  A missing file is never silent: it
    raises under HC_STRICT_PATHS=1 (a run with regenerated sets must not quietly evaluate against the shipped ones) and warns
    once per path otherwise."""
    if path and os.path.exists(path):
        return tables.dict_to_keys(torch.load(path))
    if path:
        if os.environ.get("HC_STRICT_PATHS", "0") == "1":
            raise FileNotFoundError("hiercom_b200: %s not found (HC_STRICT_PATHS=1 forbids the shipped fallback)" % path)
        if path not in _WARNED:
            _WARNED.add(path)
            warnings.warn("hiercom_b200: %s not found in %s - using the packed copy of the reference's shipped set "
                          "(set HC_STRICT_PATHS=1 to make this an error)" % (path, os.getcwd()), RuntimeWarning, stacklevel=3)
    return fallback()


def _interleave(a, b):
    return torch.stack((a, b), dim=1).reshape((-1,) + tuple(a.shape[1:])).contiguous()


def _csr_from_sorted(sorted_ids, uniq):
    """offsets [len(uniq)+1] int32 of runs in a sorted id vector."""
    bounds = torch.searchsorted(sorted_ids, uniq, right=False)
    end = torch.tensor([sorted_ids.numel()], device=sorted_ids.device, dtype=bounds.dtype)
    return torch.cat((bounds, end)).to(torch.int32).contiguous()


class _CounterMixin:
    def _counters_to_attrs(self):
        raise NotImplementedError


class Evaluator(_CounterMixin):
    """evaluator.py:15-367.  R@k / mR@k (+ zero-shot) with three candidates per directed pair (hierarchical) or one (flat)."""

    def __init__(self, args, num_classes, iou_thresh, top_k, max_cache_size=10000, device=None):
        self.args = args
        self.hierar = args['models']['hierarchical_pred']
        self.top_k = top_k
        self.num_classes = num_classes
        self.iou_thresh = iou_thresh
        self.feature_size = args['models']['feature_size']
        self.run_mode = args['training']['run_mode']
        self.splits = (args['models']['num_geometric'], args['models']['num_possessive'], args['models']['num_semantic'])
        if num_classes != tables.NUM_PRED or len(top_k) != 3:
            raise RuntimeError("hiercom_b200: counters are laid out for 50 predicates and three cut-offs (evaluate.py:79)")
        self.device = _dev(device)
        self.is_vg = args['dataset']['dataset'] == 'vg'
        self._total = torch.zeros(tables.EV_SIZE, dtype=torch.int64, device=self.device)
        self.zs_bitmap = None
        if self.is_vg:
            zs = _load_keys(args['dataset'].get('zero_shot_triplets'), tables.zero_shot_keys)
            self.zs_bitmap = torch.from_numpy(tables.keys_to_bitmap(zs).view(np.int32)).to(self.device)
        self.pass_bitmap = None
        if self.run_mode in ('train_cs', 'eval_cs'):                       # evaluator.py:76-81
            sfx = '_gpt4v' if args['models'].get('llm_model') == 'gpt4v' else ''
            if sfx and not os.path.exists('triplets/commonsense_aligned_triplets%s.pt' % sfx):
                raise FileNotFoundError('triplets/commonsense_aligned_triplets_gpt4v.pt')
            al = _load_keys('triplets/commonsense_aligned_triplets%s.pt' % sfx, tables.commonsense_aligned_keys)
            vi = _load_keys('triplets/commonsense_violated_triplets%s.pt' % sfx, tables.commonsense_violated_keys)
            self.set_commonsense(al, vi)
        self.synonyms = torch.from_numpy(tables.object_synonym_matrix()).to(self.device)
        self.annotation_paths = None
        self.clear_data()
        self._counters_to_attrs()

    def set_commonsense(self, aligned_keys, violated_keys):
        """Install commonsense sets given as packed keys (or reference-format dicts)."""
        if hasattr(aligned_keys, "keys"):
            aligned_keys = tables.dict_to_keys(aligned_keys)
        if hasattr(violated_keys, "keys"):
            violated_keys = tables.dict_to_keys(violated_keys)
        bm = ops.cs_bitmap_build(aligned_keys, violated_keys)
        self.pass_bitmap = torch.from_numpy(bm.view(np.int32)).to(self.device)

    # -------------------------------------------------------------------------------------------- recording
    def _t(self, x, dtype=None):
        if x is None:
            return None
        t = torch.as_tensor(x)
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        return t.to(self.device)

    def accumulate(self, which_in_batch, relation_pred, relation_target, super_relation_pred, connectivity,
                   subject_cat_pred, object_cat_pred, subject_cat_target, object_cat_target,
                   subject_bbox_pred, object_bbox_pred, subject_bbox_target, object_bbox_target, iou_mask,
                   predcls=True, cat_subject_confidence=None, cat_object_confidence=None, height=None, width=None):
        call = dict(which=self._t(which_in_batch, torch.int64), rel=self._t(relation_pred, torch.float32),
                    conn=self._t(connectivity, torch.float32), ov=self._t(iou_mask).to(torch.uint8),
                    cs=self._t(subject_cat_pred, torch.int32), co=self._t(object_cat_pred, torch.int32),
                    bs=self._t(subject_bbox_pred).to(torch.int32).view(-1, 4), bo=self._t(object_bbox_pred).to(torch.int32).view(-1, 4))
        if not predcls:
            call["fs"] = self._t(cat_subject_confidence, torch.float32)
            call["fo"] = self._t(cat_object_confidence, torch.float32)
        else:
            call["t"] = self._t(relation_target, torch.int32)
            call["tcs"] = self._t(subject_cat_target, torch.int32)
            call["tco"] = self._t(object_cat_target, torch.int32)
            call["tbs"] = self._t(subject_bbox_target).to(torch.int32).view(-1, 4)
            call["tbo"] = self._t(object_bbox_target).to(torch.int32).view(-1, 4)
        self._calls.append(call)
        self.relation_pred = True        # reference code tests `self.relation_pred is None` to detect an empty buffer

    def accumulate_target(self, relation_target, subject_cat_target, object_cat_target, subject_bbox_target, object_bbox_target):
        """evaluator.py:272-277 (SGDET/SGCLS): per-image lists, `None` for images without GT."""
        self._sgd_targets = (relation_target, subject_cat_target, object_cat_target, subject_bbox_target, object_bbox_target)

    def load_annotation_paths(self, annot_path):
        self.annotation_paths = annot_path

    def clear_data(self):
        self._calls = []
        self._sgd_targets = None
        self.relation_pred = None

    def clear_gpt_cache(self):
        self.cache = {}

    # -------------------------------------------------------------------------------------------- compute
    def _gather_rows(self):
        cat = lambda k: torch.cat([c[k] for c in self._calls])
        rows = {k: cat(k) for k in self._calls[0].keys()}
        sizes = torch.tensor([c["which"].numel() for c in self._calls], device=self.device)
        base = torch.cumsum(sizes, 0) - sizes
        rows["call_base"] = torch.repeat_interleave(base, sizes)
        rows["call_size"] = torch.repeat_interleave(sizes, sizes)
        return rows

    def compute(self, per_class=False, predcls=True):
        if not self._calls:
            return self._metrics()
        dev = self.device
        r = self._gather_rows()
        n = r["which"].numel()
        k = 3 if self.hierar else 1
        row_sub = torch.arange(0, 2 * n, 2, dtype=torch.int32, device=dev)
        row_obj = row_sub + 1
        pred_cat = _interleave(r["cs"], r["co"])
        pred_box = _interleave(r["bs"], r["bo"])
        conf_sub = r.get("fs") if not predcls else None
        conf_obj = r.get("fo") if not predcls else None
        cand_conf, cand_label, _, _ = ops.candidates(r["rel"].contiguous(), self.splits, self.hierar, r["ov"].contiguous(),
                                                     r["conn"].contiguous(), row_sub, row_obj, pred_cat, self.pass_bitmap, None,
                                                     conf_sub=conf_sub, conf_obj=conf_obj, layout=0)
        # reference append order: per call [k=0 rows | k=1 rows | k=2 rows] (evaluator.py:157-179,231-246); per image the
        # boolean mask keeps that global order (:295,303).
        rows_idx = torch.arange(n, device=dev).repeat_interleave(k)
        kk = torch.arange(k, device=dev).repeat(n)
        ref_idx = k * r["call_base"][rows_idx] + kk * r["call_size"][rows_idx] + (rows_idx - r["call_base"][rows_idx])
        img_of = r["which"][rows_idx]
        order = torch.argsort(img_of * (k * n + 1) + ref_idx)
        uniq = torch.unique(r["which"])                                   # sorted (evaluator.py:294)
        cand_offsets = _csr_from_sorted(img_of[order], uniq)
        cand_conf = cand_conf[order].contiguous()
        cand_label = cand_label[order].contiguous()
        cand_row = rows_idx[order].to(torch.int32).contiguous()
        if self._sgd_targets is None:
            t_order = torch.argsort(r["which"], stable=True)
            gt_offsets = _csr_from_sorted(r["which"][t_order], uniq)
            gt_label = r["t"][t_order].contiguous()
            gt_sub = (2 * t_order).to(torch.int32).contiguous()
            gt_obj = gt_sub + 1
            gt_cat = _interleave(r["tcs"], r["tco"])
            gt_box = _interleave(r["tbs"], r["tbo"])
        else:
            gt_offsets, gt_label, gt_sub, gt_obj, gt_cat, gt_box = self._sgd_tables(uniq)
        delta = torch.zeros(tables.EV_SIZE, dtype=torch.int64, device=dev)
        ops.topk_match(cand_offsets, cand_conf, cand_label, k, row_sub, row_obj, pred_cat, pred_box, gt_offsets, gt_label, gt_sub,
                       gt_obj, gt_cat, gt_box, delta, cand_row=cand_row, synonyms=None if predcls else self.synonyms,
                       zs_bitmap=self.zs_bitmap, mode=0, feature_size=self.feature_size, iou_thresh=self.iou_thresh,
                       top_k=self.top_k)
        if not per_class:                                                 # evaluator.py:336-337,344-345
            for blk in (0, tables.EV_BLOCK):
                delta[blk + tables.EV_HITS_PC:blk + tables.EV_NGT] = 0
        self._total += delta
        self._counters_to_attrs()
        return self._metrics()

    def _sgd_tables(self, uniq):
        rel_t, cs_t, co_t, bs_t, bo_t = self._sgd_targets
        offsets, label, cats, boxes = [0], [], [], []
        for img in uniq.tolist():
            rt = rel_t[int(img)]
            if rt is not None:                                            # evaluator.py:298-299
                m = int(torch.as_tensor(rt).numel())
                label.append(torch.as_tensor(rt).reshape(-1).to(torch.int32).cpu())
                cats.append(torch.stack((torch.as_tensor(cs_t[int(img)]).reshape(-1).cpu(),
                                         torch.as_tensor(co_t[int(img)]).reshape(-1).cpu()), 1).reshape(-1).to(torch.int32))
                boxes.append(torch.stack((torch.as_tensor(bs_t[int(img)]).reshape(-1, 4).cpu().to(torch.int32),
                                          torch.as_tensor(bo_t[int(img)]).reshape(-1, 4).cpu().to(torch.int32)), 1).reshape(-1, 4))
                offsets.append(offsets[-1] + m)
            else:
                offsets.append(offsets[-1])
        g = offsets[-1]
        dev = self.device
        i32 = lambda x: x.to(torch.int32).to(dev).contiguous()
        if g == 0:
            z = torch.zeros(1, dtype=torch.int32, device=dev)
            return i32(torch.tensor(offsets)), z[:0], z[:0], z[:0], torch.zeros(2, dtype=torch.int32, device=dev), \
                torch.zeros(2, 4, dtype=torch.int32, device=dev)
        sub = torch.arange(0, 2 * g, 2)
        return i32(torch.tensor(offsets)), i32(torch.cat(label)), i32(sub), i32(sub + 1), i32(torch.cat(cats)), i32(torch.cat(boxes))

    # -------------------------------------------------------------------------------------------- results
    def _counters_to_attrs(self):
        c = self._total.cpu().numpy()
        K, NP = self.top_k, tables.NUM_PRED

        def block(b):
            hits = {k: float(c[b + tables.EV_HITS + i]) for i, k in enumerate(K)}
            pc = {k: torch.as_tensor(c[b + tables.EV_HITS_PC + i * NP:b + tables.EV_HITS_PC + (i + 1) * NP].astype(np.float32))
                  for i, k in enumerate(K)}
            n = float(c[b + tables.EV_NGT])
            n_pc = torch.as_tensor(c[b + tables.EV_NGT_PC:b + tables.EV_NGT_PC + NP].astype(np.float32))
            return hits, pc, n, n_pc
        self.result_dict, self.result_per_class, self.num_connected_target, self.num_conn_target_per_class = block(0)
        if self.is_vg:
            (self.result_dict_zs, self.result_per_class_zs, self.num_connected_target_zs,
             self.num_conn_target_per_class_zs) = block(tables.EV_BLOCK)

    def counters(self):
        """int64 [408] device tensor in the tables.EV_* layout (what the multi-GPU all-reduce sums)."""
        return self._total

    def load_counters(self, counters):
        self._total.copy_(counters.to(self.device))
        self._counters_to_attrs()

    def _metrics(self):
        """evaluator.py:358-367."""
        recall_k = [self.result_dict[k] / max(self.num_connected_target, 1e-3) for k in self.top_k]
        recall_k_per_class = [self.result_per_class[k] / self.num_conn_target_per_class for k in self.top_k]
        mean_recall_k = [torch.nanmean(r) for r in recall_k_per_class]
        recall_k_zs, recall_k_per_class_zs, mean_recall_k_zs = None, None, None
        if self.is_vg:
            recall_k_zs = [self.result_dict_zs[k] / max(self.num_connected_target_zs, 1e-3) for k in self.top_k]
            recall_k_per_class_zs = [self.result_per_class_zs[k] / self.num_conn_target_per_class_zs for k in self.top_k]
            mean_recall_k_zs = [torch.nanmean(r) for r in recall_k_per_class_zs]
        return recall_k, recall_k_per_class, mean_recall_k, recall_k_zs, recall_k_per_class_zs, mean_recall_k_zs

    # -------------------------------------------------------------------------------------------- out of scope
    def get_related_top_k_predictions_parallel(self, top_k):
        raise NotImplementedError("LLM commonsense collection stays reference code (evaluator.py:375-462)")

    def compute_precision(self):
        raise NotImplementedError("OpenImages precision metric stays reference code (evaluator.py:522-566)")

    def save_visualization_results(self, *a, **k):
        raise NotImplementedError("visualisation dumps stay reference code (evaluator.py:465-519)")


class Evaluator_Top3(_CounterMixin):
    """evaluator.py:589-773.  R@k* / mR@k*: a GT counts if ANY of the three per-head argmaxes equals it."""

    def __init__(self, args, num_classes, iou_thresh, top_k, device=None):
        self.args = args
        self.top_k = top_k
        self.num_classes = num_classes
        self.iou_thresh = iou_thresh
        self.feature_size = args['models']['feature_size']
        self.splits = (args['models']['num_geometric'], args['models']['num_possessive'], args['models']['num_semantic'])
        if num_classes != tables.NUM_PRED or len(top_k) != 3:
            raise RuntimeError("hiercom_b200: counters are laid out for 50 predicates and three cut-offs")
        self.device = _dev(device)
        self._total = torch.zeros(tables.T3_SIZE, dtype=torch.int64, device=self.device)
        self.clear_data()
        self._counters_to_attrs()

    _t = Evaluator._t

    def accumulate(self, which_in_batch, relation_pred, relation_target, super_relation_pred, connectivity,
                   subject_cat_pred, object_cat_pred, subject_cat_target, object_cat_target,
                   subject_bbox_pred, object_bbox_pred, subject_bbox_target, object_bbox_target, iou_mask):
        self._calls.append(dict(
            which=self._t(which_in_batch, torch.int64), rel=self._t(relation_pred, torch.float32),
            sup=self._t(super_relation_pred, torch.float32), conn=self._t(connectivity, torch.float32),
            ov=self._t(iou_mask).to(torch.uint8), t=self._t(relation_target, torch.int32),
            cs=self._t(subject_cat_pred, torch.int32), co=self._t(object_cat_pred, torch.int32),
            tcs=self._t(subject_cat_target, torch.int32), tco=self._t(object_cat_target, torch.int32),
            bs=self._t(subject_bbox_pred).to(torch.int32).view(-1, 4), bo=self._t(object_bbox_pred).to(torch.int32).view(-1, 4),
            tbs=self._t(subject_bbox_target).to(torch.int32).view(-1, 4), tbo=self._t(object_bbox_target).to(torch.int32).view(-1, 4)))
        self.relation_pred = True

    def clear_data(self):
        self._calls = []
        self.relation_pred = None

    def compute(self, per_class=False):
        if not self._calls:
            return self._metrics()
        dev = self.device
        r = {k: torch.cat([c[k] for c in self._calls]) for k in self._calls[0].keys()}
        n = r["which"].numel()
        order = torch.argsort(r["which"], stable=True)                    # per image, append order (evaluator.py:705-709)
        uniq = torch.unique(r["which"])
        offsets = _csr_from_sorted(r["which"][order], uniq)
        row_sub = torch.arange(0, 2 * n, 2, dtype=torch.int32, device=dev)
        row_obj = row_sub + 1
        pred_cat = _interleave(r["cs"], r["co"])
        pred_box = _interleave(r["bs"], r["bo"])
        _, labels3, t3_conf, t3_super = ops.candidates(r["rel"].contiguous(), self.splits, True, r["ov"].contiguous(),
                                                       r["conn"].contiguous(), row_sub, row_obj, pred_cat, None,
                                                       r["sup"].contiguous(), want_top3=True)
        cand_row = order.to(torch.int32).contiguous()
        delta = torch.zeros(tables.T3_SIZE, dtype=torch.int64, device=dev)
        ops.topk_match(offsets, t3_conf[order].contiguous(), None, 1, row_sub, row_obj, pred_cat, pred_box, offsets,
                       r["t"][order].contiguous(), (2 * order).to(torch.int32).contiguous(), (2 * order + 1).to(torch.int32).contiguous(),
                       _interleave(r["tcs"], r["tco"]), _interleave(r["tbs"], r["tbo"]), delta, cand_row=cand_row, mode=1,
                       t3_labels=labels3, t3_super=t3_super, feature_size=self.feature_size, iou_thresh=self.iou_thresh,
                       top_k=self.top_k)
        if not per_class:
            delta[tables.T3_HITS_PC:tables.T3_TOP1] = 0
            delta[tables.T3_TOP1_PC:tables.T3_NGT] = 0
        self._total += delta
        self._counters_to_attrs()
        return self._metrics()

    def _counters_to_attrs(self):
        c = self._total.cpu().numpy()
        K, NP = self.top_k, tables.NUM_PRED
        f32 = lambda a: torch.as_tensor(a.astype(np.float32))
        self.result_dict = {k: float(c[tables.T3_HITS + i]) for i, k in enumerate(K)}
        self.result_per_class = {k: f32(c[tables.T3_HITS_PC + i * NP:tables.T3_HITS_PC + (i + 1) * NP]) for i, k in enumerate(K)}
        self.result_dict_top1 = {k: float(c[tables.T3_TOP1 + i]) for i, k in enumerate(K)}
        self.result_per_class_top1 = {k: f32(c[tables.T3_TOP1_PC + i * NP:tables.T3_TOP1_PC + (i + 1) * NP]) for i, k in enumerate(K)}
        self.num_connected_target = float(c[tables.T3_NGT])
        self.num_conn_target_per_class = f32(c[tables.T3_NGT_PC:tables.T3_NGT_PC + NP])

    def counters(self):
        return self._total

    def load_counters(self, counters):
        self._total.copy_(counters.to(self.device))
        self._counters_to_attrs()

    def _metrics(self):
        """evaluator.py:768-773."""
        recall_k = [self.result_dict[k] / max(self.num_connected_target, 1e-3) for k in self.top_k]
        recall_k_per_class = [self.result_per_class[k] / self.num_conn_target_per_class for k in self.top_k]
        mean_recall_k = [torch.nanmean(r) for r in recall_k_per_class]
        return recall_k, recall_k_per_class, mean_recall_k
