"""Drop-in twins of the hierarchical relation predictors HIERCOM adds to Scene-Graph-Benchmark
(SGB = scenegraph_benchmark/Scene-Graph-Benchmark.pytorch/maskrcnn_benchmark/modeling/roi_heads/relation_head):

  MotifHierarchicalPredictor   roi_relation_predictors.py:324-469   LSTM context,        rel_compress = BayesHead, frequency bias
  TransformerHierPredictor     roi_relation_predictors.py:135-253   transformer context, rel_compress + ctx_compress (two BayesHeads)
  VCTreeHierPredictor          roi_relation_predictors.py:588-702   VCTree context,      ctx_compress = BayesHeadProb

Same registry names, same constructor `(config, in_channels)`, same `forward(proposals, rel_pair_idxs, rel_labels, rel_binarys,
roi_features, union_features, logger=None)` and the same 6-tuple `(obj_dists, relation1_dist, relation2_dist, relation3_dist,
superrelation_dist, add_losses)`; same parameter names (`post_emb`, `post_cat`, `rel_compress.fc3_1` ..., `ctx_compress`,
`up_dim`, `freq_bias.obj_baseline`) so an SGB checkpoint loads with `load_state_dict`.

What is ours: everything after the context layer - the pair gather, `post_cat`, `* union_features`, the Bayes heads, the frequency
bias with its log-sum-exp super bias and the hierarchical log-softmax - i.e. `sgb.hierarchical_relation_tail` (tcgen05 GEMMs + the
SGB kernels of csrc/sgb.cu).  What stays SGB code (SURVEY §2b marks it outside the hot path): the context encoders (`LSTMContext`,
`TransformerContext`, `VCTreeLSTMContext`), `FrequencyBias`' statistics loading and the training-only binary loss.  They are imported
lazily from an installed `maskrcnn_benchmark`, or handed in (`context_layer=`, `statistics=`), which is how the tests run without SGB.

`register(registry)` puts the three classes into SGB's `registry.ROI_RELATION_PREDICTOR` under the reference's names.
CausalAnalysisHierPredictor (roi_relation_predictors.py:1093-1476) keeps its own fusion / counterfactual logic; only its last four
lines are this path (`hier_log_softmax`).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, sgb

_SGB_PKG = "maskrcnn_benchmark.modeling.roi_heads.relation_head"


def _sgb_import(module, name):
    import importlib
    try:
        return getattr(importlib.import_module(_SGB_PKG + "." + module), name)
    except Exception as e:          # noqa: BLE001 - SGB's own import chain (native _C extension, yacs ...) can fail in many ways
        raise RuntimeError("hiercom_b200: %s.%s.%s is not importable (%s); pass the object to the predictor's constructor instead"
                           % (_SGB_PKG, module, name, e)) from e


class FrequencyBias(nn.Module):
    """model_motifs.py:12-51 - P(predicate | subject class, object class) as an embedding over class pairs; the tail kernel gathers
    rows of `obj_baseline.weight` itself, `index_with_labels` is kept for callers outside the path."""

    def __init__(self, cfg, statistics, eps=1e-3):
        super().__init__()
        pred_dist = statistics["pred_dist"].float()
        assert pred_dist.size(0) == pred_dist.size(1)
        self.num_objs, self.num_rels = pred_dist.size(0), pred_dist.size(2)
        self.obj_baseline = nn.Embedding(self.num_objs * self.num_objs, self.num_rels)
        with torch.no_grad():
            self.obj_baseline.weight.copy_(pred_dist.view(-1, self.num_rels))

    def index_with_labels(self, labels):
        return self.obj_baseline(labels[:, 0] * self.num_objs + labels[:, 1])

    forward = index_with_labels


def _layer_init(layer, init_para=0.1, normal=False):
    """utils_relation.py:81-90."""
    if normal:
        nn.init.normal_(layer.weight, mean=0, std=init_para)
    else:
        nn.init.xavier_normal_(layer.weight, gain=1.0)
    nn.init.constant_(layer.bias, 0)


def hier_log_softmax(rel1_logits, rel2_logits, rel3_logits, super_logits):
    """The four lines every hierarchical predictor ends with (roi_relation_predictors.py:241-244, 456-459, 1455-1458):
    super = log_softmax(super_logits); relation_k = log_softmax(rel_k_logits) + super[:, k] (super slot 0 = background).
    One launch of the SGB softmax kernel; returns (relation_1, relation_2, relation_3, super_relation)."""
    g, p, s = rel1_logits.shape[1], rel2_logits.shape[1], rel3_logits.shape[1]
    z = torch.cat((rel1_logits, rel2_logits, rel3_logits, super_logits), dim=1).float().contiguous()
    rel, sup = ops.sgb_hier_softmax(z, (g, p, s))
    return rel[:, :g], rel[:, g:g + p], rel[:, g + p:], sup


class _HierPredictorBase(nn.Module):
    context_module, context_class = None, None

    def __init__(self, config, in_channels, context_layer=None, statistics=None):
        super().__init__()
        m = config.MODEL
        self.attribute_on = m.ATTRIBUTE_ON
        self.num_obj_cls = m.ROI_BOX_HEAD.NUM_CLASSES
        self.num_att_cls = m.ROI_ATTRIBUTE_HEAD.NUM_ATTRIBUTES
        self.num_rel_cls = m.ROI_RELATION_HEAD.NUM_CLASSES
        assert in_channels is not None
        self.use_vision = m.ROI_RELATION_HEAD.PREDICT_USE_VISION
        self.use_bias = m.ROI_RELATION_HEAD.PREDICT_USE_BIAS
        if statistics is None:
            statistics = self._dataset_statistics(config)
        self._check_statistics(statistics)
        self.context_layer = context_layer if context_layer is not None else self._build_context(config, statistics, in_channels)
        self.hidden_dim = m.ROI_RELATION_HEAD.CONTEXT_HIDDEN_DIM
        self.pooling_dim = m.ROI_RELATION_HEAD.CONTEXT_POOLING_DIM
        self.post_emb = nn.Linear(self.hidden_dim, self.hidden_dim * 2)
        self.post_cat = nn.Linear(self.hidden_dim * 2, self.pooling_dim)
        _layer_init(self.post_emb, 10.0 * (1.0 / self.hidden_dim) ** 0.5, normal=True)
        _layer_init(self.post_cat)
        if self.pooling_dim != m.ROI_BOX_HEAD.MLP_HEAD_DIM:
            self.union_single_not_match = True
            self.up_dim = nn.Linear(m.ROI_BOX_HEAD.MLP_HEAD_DIM, self.pooling_dim)
            _layer_init(self.up_dim)
        else:
            self.union_single_not_match = False
        self._statistics = statistics

    @staticmethod
    def _dataset_statistics(config):
        import importlib
        try:
            return importlib.import_module("maskrcnn_benchmark.data").get_dataset_statistics(config)
        except Exception as e:      # noqa: BLE001
            raise RuntimeError("hiercom_b200: maskrcnn_benchmark.data.get_dataset_statistics is not importable (%s); pass "
                               "statistics= (obj_classes, rel_classes, att_classes, pred_dist) to the predictor" % e) from e

    def _check_statistics(self, st):
        assert self.num_obj_cls == len(st["obj_classes"])
        assert self.num_att_cls == len(st["att_classes"])
        assert self.num_rel_cls == len(st["rel_classes"])

    def _build_context(self, config, statistics, in_channels):
        raise NotImplementedError

    def _context(self, roi_features, proposals, rel_pair_idxs, logger):
        """-> (obj_dists, obj_preds, edge_ctx, binary_preds or None)"""
        raise NotImplementedError

    def _union(self, union_features):
        # up_dim only exists when CONTEXT_POOLING_DIM != MLP_HEAD_DIM (not config 5): a plain library GEMM, off the measured path
        return self.up_dim(union_features) if self.union_single_not_match else union_features

    @staticmethod
    def _split(tensors, num_rels):
        return tuple(t if isinstance(t, tuple) else t.split(num_rels, dim=0) for t in tensors)


class MotifHierarchicalPredictor(_HierPredictorBase):
    """roi_relation_predictors.py:324-469."""

    def __init__(self, config, in_channels, context_layer=None, statistics=None):
        super().__init__(config, in_channels, context_layer, statistics)
        self.rel_compress = sgb.BayesHead(input_dim=self.pooling_dim)
        self.rel_compress.layer_init()
        if self.use_bias:
            self.freq_bias = FrequencyBias(config, self._statistics)
        del self._statistics

    def _build_context(self, config, statistics, in_channels):
        if self.attribute_on:
            return _sgb_import("model_motifs_with_attribute", "AttributeLSTMContext")(config, statistics["obj_classes"], statistics["att_classes"],
                                                                                      statistics["rel_classes"], in_channels)
        return _sgb_import("model_motifs", "LSTMContext")(config, statistics["obj_classes"], statistics["rel_classes"], in_channels)

    def forward(self, proposals, rel_pair_idxs, rel_labels, rel_binarys, roi_features, union_features, logger=None):
        if self.attribute_on:
            obj_dists, obj_preds, att_dists, edge_ctx = self.context_layer(roi_features, proposals, logger)
        else:
            obj_dists, obj_preds, edge_ctx, _ = self.context_layer(roi_features, proposals, logger)
        edge_rep = self.post_emb(edge_ctx)                               # :399 (per OBJECT: upstream of the per-pair path)
        num_objs = [len(b) for b in proposals]
        assert len(rel_pair_idxs) == len(num_objs)
        r1, r2, r3, sup = sgb.hierarchical_relation_tail(
            edge_rep, rel_pair_idxs, num_objs, obj_preds, self._union(union_features), self.post_cat, self.rel_compress,
            self.freq_bias.obj_baseline.weight if self.use_bias else None, use_vision=self.use_vision)
        return obj_dists.split(num_objs, dim=0), r1, r2, r3, sup, {}


class TransformerHierPredictor(_HierPredictorBase):
    """roi_relation_predictors.py:135-253: logits = rel_compress(post_cat(prod) * union) + ctx_compress(prod); no frequency bias in
    the forward (the reference builds `freq_bias` but never reads it, :186-187)."""

    def __init__(self, config, in_channels, context_layer=None, statistics=None):
        super().__init__(config, in_channels, context_layer, statistics)
        self.rel_compress = sgb.BayesHead(self.pooling_dim)
        self.ctx_compress = sgb.BayesHead(self.hidden_dim * 2)
        self.rel_compress.layer_init()
        self.ctx_compress.layer_init()
        if self.use_bias:
            self.freq_bias = FrequencyBias(config, self._statistics)
        del self._statistics

    def _build_context(self, config, statistics, in_channels):
        return _sgb_import("model_transformer", "TransformerContext")(config, statistics["obj_classes"], statistics["rel_classes"], in_channels)

    def forward(self, proposals, rel_pair_idxs, rel_labels, rel_binarys, roi_features, union_features, logger=None):
        if self.attribute_on:
            obj_dists, obj_preds, att_dists, edge_ctx = self.context_layer(roi_features, proposals, logger)
        else:
            obj_dists, obj_preds, edge_ctx = self.context_layer(roi_features, proposals, logger)
        edge_rep = self.post_emb(edge_ctx)
        num_objs = [len(b) for b in proposals]
        assert len(rel_pair_idxs) == len(num_objs)
        if not self.use_vision:
            raise NotImplementedError("TransformerHierPredictor reads visual_rep unconditionally (roi_relation_predictors.py:231)")
        r1, r2, r3, sup = sgb.hierarchical_relation_tail(edge_rep, rel_pair_idxs, num_objs, obj_preds, self._union(union_features),
                                                         self.post_cat, self.rel_compress, None, ctx_compress=self.ctx_compress)
        return obj_dists.split(num_objs, dim=0), r1, r2, r3, sup, {}


class VCTreeHierPredictor(_HierPredictorBase):
    """roi_relation_predictors.py:588-702: edge_rep = relu(post_emb(edge_ctx)); ctx_compress = BayesHeadProb over
    post_cat(prod) * union; the frequency bias is constructed but not applied (:637, 682-685)."""

    def __init__(self, config, in_channels, context_layer=None, statistics=None):
        super().__init__(config, in_channels, context_layer, statistics)
        self.ctx_compress = sgb.BayesHeadProb(self.pooling_dim)
        self.ctx_compress.layer_init()
        self.freq_bias = FrequencyBias(config, self._statistics)
        del self._statistics

    def _build_context(self, config, statistics, in_channels):
        return _sgb_import("model_vctree", "VCTreeLSTMContext")(config, statistics["obj_classes"], statistics["rel_classes"], statistics, in_channels)

    def forward(self, proposals, rel_pair_idxs, rel_labels, rel_binarys, roi_features, union_features, logger=None):
        obj_dists, obj_preds, edge_ctx, binary_preds = self.context_layer(roi_features, proposals, rel_pair_idxs, logger)
        edge_rep = F.relu(self.post_emb(edge_ctx))
        num_objs = [len(b) for b in proposals]
        assert len(rel_pair_idxs) == len(num_objs)
        r1, r2, r3, sup = sgb.hierarchical_relation_tail(edge_rep, rel_pair_idxs, num_objs, obj_preds, self._union(union_features),
                                                         self.post_cat, self.ctx_compress, None)
        add_losses = {}
        if self.training:                                                  # :694-700 (training only; plain torch)
            binary_loss = [F.binary_cross_entropy_with_logits(bp, (bg > 0).float()) for bg, bp in zip(rel_binarys, binary_preds)]
            add_losses["binary_loss"] = sum(binary_loss) / len(binary_loss)
        return obj_dists.split(num_objs, dim=0), r1, r2, r3, sup, add_losses


PREDICTORS = {"MotifHierarchicalPredictor": MotifHierarchicalPredictor, "TransformerHierPredictor": TransformerHierPredictor,
              "VCTreeHierPredictor": VCTreeHierPredictor}


def register(registry=None):
    """Registers the three classes in SGB's `registry.ROI_RELATION_PREDICTOR` under the reference's own names, replacing the stock
    entries: `make_roi_relation_predictor(cfg, in_channels)` (roi_relation_predictors.py:1479-1481) then builds ours."""
    if registry is None:
        import importlib
        registry = importlib.import_module("maskrcnn_benchmark.modeling.registry").ROI_RELATION_PREDICTOR
    for name, cls in PREDICTORS.items():
        registry[name] = cls
    return registry
