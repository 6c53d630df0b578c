"""Drop-in hierarchical predictors of the SGB twin (scene_graph_commonsense_b200/sgb_predictors.py): interface checks on the CPU
(constructor, parameter names of the reference so SGB checkpoints load, registry names) and, on a GPU, the forward of each class
against goldens produced by the REAL SGB predictor classes (oracle/make_golden_sgb.py)."""
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch
import torch.nn as nn

from scene_graph_commonsense_b200 import synthetic
from tests import helpers
from tests.golden_cases import SGB_CASES, SGB_VARIANT_CASE

DEV = "cuda"


def _config(pooling=4096, mlp_head=4096, use_bias=True):
    rel = NS(NUM_CLASSES=51, PREDICT_USE_VISION=True, PREDICT_USE_BIAS=use_bias, CONTEXT_HIDDEN_DIM=512, CONTEXT_POOLING_DIM=pooling)
    return NS(MODEL=NS(ATTRIBUTE_ON=False, ROI_BOX_HEAD=NS(NUM_CLASSES=151, MLP_HEAD_DIM=mlp_head),
                       ROI_ATTRIBUTE_HEAD=NS(NUM_ATTRIBUTES=201), ROI_RELATION_HEAD=rel))


def _statistics(freq=None):
    pd = torch.zeros(151, 151, 51) if freq is None else freq.view(151, 151, 51)
    return dict(obj_classes=["o%d" % i for i in range(151)], rel_classes=["r%d" % i for i in range(51)],
                att_classes=["a%d" % i for i in range(201)], pred_dist=pd)


class _Ctx(nn.Module):
    """Stands in for LSTMContext / TransformerContext / VCTreeLSTMContext (outside the path): returns the seeded context."""

    def __init__(self, edge_ctx, obj_preds, kind):
        super().__init__()
        self.edge_ctx, self.obj_preds, self.kind = edge_ctx, obj_preds, kind

    def forward(self, roi_features, proposals, *rest):
        dists = torch.zeros(self.edge_ctx.shape[0], 151, device=self.edge_ctx.device)
        if self.kind == "transformer":
            return dists, self.obj_preds, self.edge_ctx
        return dists, self.obj_preds, self.edge_ctx, None


def _copy_head(head, sd):
    with torch.no_grad():
        for n in ("fc3_1", "fc3_2", "fc3_3", "fc5"):
            getattr(head, n).weight.copy_(sd[n + ".weight"]); getattr(head, n).bias.copy_(sd[n + ".bias"])


def _build(kind, sd, batch, device, sd_ctx=None):
    from scene_graph_commonsense_b200 import sgb_predictors as SP
    cls = {"motif": SP.MotifHierarchicalPredictor, "transformer": SP.TransformerHierPredictor, "vctree": SP.VCTreeHierPredictor}[kind]
    ctx = _Ctx(batch["edge_ctx"].to(device), batch["obj_labels"].to(device), kind)
    p = cls(_config(), 4096, context_layer=ctx, statistics=_statistics(sd["freq_bias"]))
    with torch.no_grad():
        p.post_emb.weight.copy_(sd["post_emb.weight"]); p.post_emb.bias.copy_(sd["post_emb.bias"])
        p.post_cat.weight.copy_(sd["post_cat.weight"]); p.post_cat.bias.copy_(sd["post_cat.bias"])
    if kind == "motif":
        _copy_head(p.rel_compress, sd)
    elif kind == "transformer":
        _copy_head(p.rel_compress, sd)
        _copy_head(p.ctx_compress, sd_ctx)
    else:
        _copy_head(p.ctx_compress, sd)
    return p.to(device).eval()


def test_predictor_classes_keep_the_reference_interface():
    from scene_graph_commonsense_b200 import sgb_predictors as SP
    batch = synthetic.make_sgb_batch([3, 2], seed=5)
    sd = synthetic.sgb_state_dict(seed=5)
    want = {
        "motif": {"post_emb", "post_cat", "rel_compress.fc3_1", "rel_compress.fc3_2", "rel_compress.fc3_3", "rel_compress.fc5", "freq_bias.obj_baseline"},
        "transformer": {"post_emb", "post_cat", "rel_compress.fc3_1", "rel_compress.fc5", "ctx_compress.fc3_1", "ctx_compress.fc5", "freq_bias.obj_baseline"},
        "vctree": {"post_emb", "post_cat", "ctx_compress.fc3_1", "ctx_compress.fc3_3", "ctx_compress.fc5", "freq_bias.obj_baseline"},
    }
    for kind, names in want.items():
        p = _build(kind, sd, batch, "cpu", sd_ctx=synthetic.sgb_state_dict(seed=6, pooling=1024))
        keys = {k.rsplit(".", 1)[0] for k in p.state_dict()}
        assert names <= keys, (kind, names - keys)
        assert torch.equal(p.freq_bias.obj_baseline.weight, sd["freq_bias"])          # statistics['pred_dist'] is copied in (model_motifs.py:27-29)
    assert SP.TransformerHierPredictor(_config(), 4096, context_layer=nn.Identity(), statistics=_statistics()).ctx_compress.fc3_1.in_features == 1024
    up = SP.MotifHierarchicalPredictor(_config(pooling=4096, mlp_head=2048), 2048, context_layer=nn.Identity(), statistics=_statistics())
    assert up.union_single_not_match and up.up_dim.in_features == 2048 and up.up_dim.out_features == 4096
    registry = {}
    assert SP.register(registry) is registry
    assert set(registry) == {"MotifHierarchicalPredictor", "TransformerHierPredictor", "VCTreeHierPredictor"}    # roi_relation_predictors.py:135,324,588
    with pytest.raises(RuntimeError):          # no installed maskrcnn_benchmark here: the context layer has to be handed in
        SP.MotifHierarchicalPredictor(_config(), 4096, statistics=_statistics())


def _pairs(num_objs):
    return [torch.nonzero(torch.ones(n, n) - torch.eye(n)).view(-1, 2) for n in num_objs]


def _worst(got, g, fmt, nimg):
    worst = 0.0
    for i in range(nimg):
        for t, key in zip(got, fmt):
            ref = g[key % i]
            worst = max(worst, float(np.abs(np.exp(t[i].cpu().numpy().astype(np.float64)) - np.exp(ref.astype(np.float64))).max()))
    return worst


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SGB_CASES))
def test_motif_hierarchical_predictor_forward_matches_real_sgb(name):
    c = SGB_CASES[name]
    g = helpers.golden(name)
    batch, sd = synthetic.make_sgb_batch(c["num_objs"], seed=c["seed"]), synthetic.sgb_state_dict(seed=c["seed"])
    p = _build("motif", sd, batch, DEV)
    proposals = [[0] * n for n in batch["num_objs"]]                       # forward only takes len(); BoxLists in SGB
    out = p(proposals, [x.to(DEV) for x in _pairs(batch["num_objs"])], None, None, None, batch["union_features"].to(DEV), None)
    assert len(out) == 6 and out[5] == {} and len(out[0]) == len(batch["num_objs"])
    assert _worst(out[1:5], g, ("rel1_%d", "rel2_%d", "rel3_%d", "sup_%d"), len(batch["num_objs"])) <= 2e-3


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["transformer", "vctree"])
def test_transformer_and_vctree_hier_predictors_match_real_sgb(kind):
    c = SGB_VARIANT_CASE
    g = helpers.golden("sgb_variants")
    batch, sd = synthetic.make_sgb_batch(c["num_objs"], seed=c["seed"]), synthetic.sgb_state_dict(seed=c["seed"])
    p = _build(kind, sd, batch, DEV, sd_ctx=synthetic.sgb_state_dict(seed=c["seed"] + 100, pooling=1024))
    proposals = [[0] * n for n in batch["num_objs"]]
    out = p(proposals, [x.to(DEV) for x in _pairs(batch["num_objs"])], None, None, None, batch["union_features"].to(DEV), None)
    worst = 0.0
    for i in range(len(batch["num_objs"])):
        rel = torch.cat((out[1][i], out[2][i], out[3][i]), dim=1).cpu().numpy().astype(np.float64)
        worst = max(worst, float(np.abs(np.exp(rel) - np.exp(g["%s_rel_%d" % (kind, i)].astype(np.float64))).max()),
                    float(np.abs(np.exp(out[4][i].cpu().numpy().astype(np.float64)) - np.exp(g["%s_sup_%d" % (kind, i)].astype(np.float64))).max()))
    assert worst <= 2e-3, worst
