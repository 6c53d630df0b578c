"""Golden vectors for the SGDET / SGCLS proposal front-end, produced by the UNMODIFIED reference.

Run in the build container only:   python oracle/make_golden_frontend.py      -> tests/golden/fe_*.npz (committed)

The front-end is inline code inside `evaluate.eval_sgd` / `eval_sgc` (which also build DDP, DETR and a DataLoader), so
it cannot be called.  Instead of restating it, this script SLICES the reference's own source lines out of
/root/reference/evaluate.py at run time (from the first `logits_pred = torch.argmax(F.softmax(...` of eval_sgd up to
the `PREPARE TARGETS` docstring: evaluate.py:309-368, super-category lookup included), dedents them and executes them
with synthetic `out_dict` tensors on the CPU - reference source is neither modified nor copied into the repo.
`utils.match_object_categories` and `utils.match_target_sgd` are imported and called directly.
Tie normalisation (H1 twin): `torch.topk` is patched to (value desc, index asc) while the reference computes.
"""
import os
import sys
import textwrap
import types

import numpy as np
import torch
import torch.nn.functional as F
import torchvision

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("HIERCOM_REFERENCE", "/root/reference")
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
sys.modules.setdefault("torchmetrics", types.ModuleType("torchmetrics"))
os.chdir(REF)

import dataset_utils as ref_du       # noqa: E402
import utils as ref_utils            # noqa: E402

from scene_graph_commonsense_b200 import synthetic  # noqa: E402
from tests.golden_cases import FRONTEND_CASES       # noqa: E402

_orig_topk = torch.topk


def _stable_topk(x, k, dim=-1, largest=True, sorted=True):
    v, i = torch.sort(x, dim=dim, descending=largest, stable=True)
    return torch.return_types.topk((v.narrow(dim, 0, k), i.narrow(dim, 0, k)))


def reference_frontend_source():
    lines = open(os.path.join(REF, "evaluate.py")).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.strip().startswith("logits_pred = torch.argmax(F.softmax(out_dict['pred_logits']"))
    end = next(i for i in range(start, len(lines)) if "PREPARE TARGETS" in lines[i]) - 1     # the opening triple quote
    return textwrap.dedent("\n".join(lines[start:end])), (start + 1, end)


def run_reference_frontend(pred_logits, pred_boxes):
    src, span = reference_frontend_source()
    args = {"models": {"num_classes": 150, "topk_cat": 2, "feature_size": 32, "nms": 0.5}}
    ns = dict(torch=torch, F=F, torchvision=torchvision, args=args, rank="cpu",
              out_dict={"pred_logits": pred_logits.clone(), "pred_boxes": pred_boxes.clone()},
              object_class_alp2fre_dict=ref_du.object_class_alp2fre(),
              sub2super_cat_dict=torch.load("datasets/vg_scene_graph_annot/sub2super_cat_dict.pt"))
    torch.topk = _stable_topk
    try:
        exec(compile(src, "evaluate.py[%d:%d]" % span, "exec"), ns)
    finally:
        torch.topk = _orig_topk
    return ns, span


def ragged(arrs, dtype):
    arrs = [np.asarray(a).astype(dtype) for a in arrs]
    off = np.concatenate(([0], np.cumsum([len(a) for a in arrs]))).astype(np.int64)
    flat = np.concatenate(arrs) if arrs else np.zeros(0, dtype)
    return flat, off


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, case in FRONTEND_CASES.items():
        samples = synthetic.make_batch(case["ids"], case["n_gt"], with_maps=False, p_rel=0.5)
        logits, boxes = synthetic.make_detr_outputs(samples, num_queries=case["queries"], **case.get("kw", {}))
        ns, span = run_reference_frontend(logits, boxes)
        # the reference drops images without object queries from its lists; the cases are built so that none is dropped
        assert len(ns["categories_pred"]) == len(samples), "case must keep every image (see frontend_oracle.detr_proposals)"
        out = dict(pred_logits=logits.numpy(), pred_boxes=boxes.numpy())
        out["cats"], out["offsets"] = ragged([c.numpy() for c in ns["categories_pred"]], np.int64)
        out["conf"], _ = ragged([c.numpy() for c in ns["cat_pred_confidence"]], np.float32)
        out["bbox"] = np.concatenate([b.numpy().reshape(-1, 4) for b in ns["bbox_pred"]]).astype(np.float32)
        out["masks_sum"] = np.concatenate([m.numpy().reshape(m.shape[0], -1).sum(1) for m in ns["masks_pred"]]).astype(np.int64)
        sup = -np.ones((len(out["cats"]), 4), dtype=np.int8)
        r = 0
        for img in ns["super_categories_pred"]:
            for sc in img:
                v = sc.numpy().reshape(-1)
                sup[r, :len(v)] = v
                r += 1
        out["supers"] = sup
        # SGCLS: utils.match_object_categories on the same proposals against the GT boxes (evaluate.py:605)
        bbox_target = [s.bbox.clone() for s in samples]
        torch.topk = _stable_topk
        try:
            m_cat, m_conf, m_box = ref_utils.match_object_categories(ns["categories_pred"], ns["cat_pred_confidence"],
                                                                     ns["bbox_pred"], bbox_target)
        finally:
            torch.topk = _orig_topk
        out["moc_none"] = np.array(m_cat is None)
        if m_cat is not None:
            out["moc_cats"], out["moc_offsets"] = ragged([[int(c) for c in img] for img in m_cat], np.int64)
            out["moc_conf"], _ = ragged([[float(c) for c in img] for img in m_conf], np.float32)
            out["moc_box"] = np.concatenate([b.numpy().reshape(-1, 4) for b in m_box]).astype(np.int32)
        # utils.match_target_sgd (evaluate.py:376)
        cs, co, bs_, bo_, rel = ref_utils.match_target_sgd("cpu", [s.relationships for s in samples], [s.subj_or_obj for s in samples],
                                                           [s.categories for s in samples], [s.bbox for s in samples])
        t_off = [0]
        t_rel, t_cs, t_co, t_bs, t_bo = [], [], [], [], []
        for i in range(len(samples)):
            if rel[i] is None:
                t_off.append(t_off[-1])
                continue
            t_rel.append(rel[i].numpy().reshape(-1)); t_cs.append(cs[i].numpy().reshape(-1)); t_co.append(co[i].numpy().reshape(-1))
            t_bs.append(bs_[i].numpy().reshape(-1, 4)); t_bo.append(bo_[i].numpy().reshape(-1, 4))
            t_off.append(t_off[-1] + len(t_rel[-1]))
        cat = lambda l, w=None: (np.concatenate(l) if l else np.zeros((0,) if w is None else (0, w))).astype(np.int32)
        out.update(tgt_offsets=np.asarray(t_off, np.int64), tgt_rel=cat(t_rel), tgt_cat_sub=cat(t_cs), tgt_cat_obj=cat(t_co),
                   tgt_box_sub=cat(t_bs, 4), tgt_box_obj=cat(t_bo, 4))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, "evaluate.py lines %d-%d executed;" % span, "proposals/img", np.diff(out["offsets"]).tolist(),
              "matched/img", (np.diff(out["moc_offsets"]).tolist() if m_cat is not None else None),
              "targets/img", np.diff(out["tgt_offsets"]).tolist())


if __name__ == "__main__":
    main()
