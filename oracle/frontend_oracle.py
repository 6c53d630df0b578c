"""ORACLE (test infrastructure, never imported by the product): CPU restatement of the SGDET / SGCLS proposal
front-end that feeds the relation path.

Reference code restated (bowen-upenn/scene_graph_commonsense @ 3388036f):
  evaluate.py:309-365 (eval_sgd) == evaluate.py:545-591 (eval_sgc)   DETR logits/boxes -> labelled, NMS-ed proposals
  evaluate.py:368                                                      super-category lookup
  utils.py:58-74      iou                                              rasterised 32x32 IoU of two (x1,x2,y1,y2) boxes
  utils.py:376-422    match_object_categories                         SGCLS: label every GT box from its best proposals
  utils.py:294-352    match_target_sgd                                 (restated in hiercom_oracle.match_target_sgd)
Third-party arithmetic the reference reaches here: `F.softmax`, `torch.topk` (PyTorch) and `torchvision.ops.nms`
(torchvision 0.15.2 pinned by requirements.txt:158-160; csrc/ops/cpu/nms_kernel.cpp).  The NMS below restates that
published greedy algorithm in float32.

PINNING: oracle/make_golden_frontend.py executes the reference's own source lines (sliced out of evaluate.py at run
time, not copied) and utils.match_object_categories on seeded inputs; tests/test_frontend_oracle_golden.py replays the
committed outputs through this file.  Tie normalisation (the same decision as H1): `torch.topk` has no defined order
among equal values, so the golden generator patches it to (value desc, index asc), which is the order used here.
"""
import numpy as np
import torch

F32 = np.float32


def detr_proposals(pred_logits, pred_boxes, alp2fre, num_classes=150, topk_cat=2, feature_size=32, nms_thresh=0.5):
    """evaluate.py:309-365.  pred_logits [B,Q,num_classes+1] f32, pred_boxes [B,Q,4] f32 (cx,cy,w,h in 0..1).
    Returns one dict per image (categories int64 [n], conf f32 [n], bbox f32 [n,4] = (x1,x2,y1,y2) on the grid), in
    the reference's output order (class ascending, then NMS keep order = confidence descending, stable).
    Deviation (documented in DESIGN.md): the reference DROPS images without any object query from its lists
    (`... for i in range(B) if torch.sum(has_object_pred[i]) > 0`), silently misaligning them with the per-image
    targets; here such an image stays in place with n = 0."""
    logits = torch.as_tensor(pred_logits, dtype=torch.float32)
    prob = torch.softmax(logits, dim=2)                                   # :309-314 (three identical softmax calls)
    has_obj = (torch.argmax(prob, dim=2) < num_classes).numpy()           # :309-310
    top_v, top_i = torch.topk(prob, k=topk_cat, dim=2)                    # :311-313
    top_v, top_i = top_v.numpy(), top_i.numpy()
    boxes = np.asarray(pred_boxes, dtype=F32)
    alp2fre = np.asarray(alp2fre)
    out = []
    for i in range(logits.shape[0]):
        q = np.nonzero(has_obj[i])[0]
        cats = alp2fre[top_i[i, q, :].reshape(-1)].astype(np.int64)       # :315-322
        conf = top_v[i, q, :].reshape(-1).astype(F32)                     # :314
        c = boxes[i, q]                                                   # :325-332
        half_w, half_h = c[:, 2] / F32(2), c[:, 3] / F32(2)
        bb = np.stack((c[:, 0] - half_w, c[:, 0] + half_w, c[:, 1] - half_h, c[:, 1] + half_h), axis=1).astype(F32)
        bb = np.clip(bb, F32(0), F32(1)) * F32(feature_size)
        bb = np.repeat(bb, topk_cat, axis=0)
        keep = cats != num_classes                                        # :323,341-345
        cats, conf, bb = cats[keep], conf[keep], bb[keep]
        order = []                                                        # :348-365 per-class NMS
        for cls in np.unique(cats):
            idx = np.nonzero(cats == cls)[0]
            k = nms_xyxy(bb[idx][:, [0, 2, 1, 3]], conf[idx], nms_thresh)
            order.extend(idx[k].tolist())
        order = np.asarray(order, dtype=np.int64)
        out.append(dict(categories=cats[order], conf=conf[order], bbox=bb[order].reshape(-1, 4)))
    return out


def nms_xyxy(boxes, scores, thresh):
    """torchvision.ops.nms (CPU kernel) in float32: visit boxes by descending score (stable), keep a box unless an
    earlier kept box overlaps it with inter / (area_i + area_j - inter) > thresh.  Returns kept indices in visit order."""
    x1, y1, x2, y2 = (boxes[:, j].astype(F32) for j in range(4))
    areas = ((x2 - x1) * (y2 - y1)).astype(F32)
    order = np.argsort(-scores.astype(F32), kind="stable")
    suppressed = np.zeros(len(scores), dtype=bool)
    keep = []
    for a, i in enumerate(order):
        if suppressed[i]:
            continue
        keep.append(i)
        for j in order[a + 1:]:
            if suppressed[j]:
                continue
            w = max(F32(0), F32(min(x2[i], x2[j]) - max(x1[i], x1[j])))
            h = max(F32(0), F32(min(y2[i], y2[j]) - max(y1[i], y1[j])))
            inter = F32(w * h)
            with np.errstate(invalid="ignore", divide="ignore"):
                ovr = F32(inter / F32(F32(areas[i] + areas[j]) - inter))
            if ovr > thresh:
                suppressed[j] = True
    return np.asarray(keep, dtype=np.int64)


def _clamp_slice(lo, hi, size):
    """Python slice semantics of mask[int(lo):int(hi)] on an axis of `size` cells (negative = from the end)."""
    lo, hi = int(lo), int(hi)
    lo = max(size + lo, 0) if lo < 0 else min(lo, size)
    hi = max(size + hi, 0) if hi < 0 else min(hi, size)
    return lo, max(hi, lo)


def grid_iou(box_t, box_p, size=32):
    """utils.py:58-74: IoU of the rasterised rectangles, as a Python float (double)."""
    ax0, ax1 = _clamp_slice(box_t[0], box_t[1], size)
    ay0, ay1 = _clamp_slice(box_t[2], box_t[3], size)
    bx0, bx1 = _clamp_slice(box_p[0], box_p[1], size)
    by0, by1 = _clamp_slice(box_p[2], box_p[3], size)
    a, b = (ax1 - ax0) * (ay1 - ay0), (bx1 - bx0) * (by1 - by0)
    inter = max(0, min(ax1, bx1) - max(ax0, bx0)) * max(0, min(ay1, by1) - max(ay0, by0))
    union = a + b - inter
    return 0.0 if union == 0 else float(inter) / float(union)


def match_object_categories(categories_pred, cat_pred_confidence, bbox_pred, bbox_target):
    """utils.py:376-422.  Per image i: for every GT box take the two proposals of largest grid IoU (float32 values,
    order = IoU desc then proposal index asc); if the two IoUs are EQUAL emit both labels and repeat the GT box,
    else emit the best one; confidence = proposal confidence * IoU (float32).  Returns (cats, confs, boxes) as lists of
    arrays, or (None, None, None) if the batch sizes differ or any image has fewer than two proposals (the reference
    returns from inside the loop over GT boxes, so an image with zero GT boxes never triggers it)."""
    if len(bbox_target) != len(bbox_pred):
        return None, None, None
    cats_out, conf_out, box_out = [], [], []
    for i in range(len(bbox_target)):
        cp, fp = np.asarray(categories_pred[i]), np.asarray(cat_pred_confidence[i], dtype=F32)
        bp, bt = np.asarray(bbox_pred[i]), np.asarray(bbox_target[i])
        cats, conf, boxes = [], [], []
        for k in range(len(bt)):
            if len(bp) < 2:
                return None, None, None
            ious = np.array([grid_iou(bt[k], bp[j]) for j in range(len(bp))], dtype=np.float64).astype(F32)
            order = np.argsort(-ious, kind="stable")[:2]
            if ious[order[0]] == ious[order[1]]:
                for j in order:
                    cats.append(cp[j]); conf.append(F32(fp[j] * ious[j])); boxes.append(bt[k])
            else:
                j = order[0]
                cats.append(cp[j]); conf.append(F32(fp[j] * ious[j])); boxes.append(bt[k])
        cats_out.append(np.asarray(cats, dtype=np.int64))
        conf_out.append(np.asarray(conf, dtype=F32))
        box_out.append(np.asarray(boxes).reshape(-1, 4).astype(bt.dtype))
    return cats_out, conf_out, box_out
