"""SGDET / SGCLS proposal front-end: the step immediately before the relation path.

The reference has this inline in `evaluate.eval_sgd` / `eval_sgc` (evaluate.py:311-370 == :545-591) and in
`utils.match_object_categories` (utils.py:376-422) / `utils.match_target_sgd` (utils.py:294-352), as Python loops with
a device sync per element (`.item()` per label at :321, `int()` per box coordinate at :339, an O(N_gt * N_prop) loop of
32x32 mask rasterisations at utils.py:398-401).  Here they are sm_100a kernels over CSR image segments (csrc/frontend.cu)
behind the C ABI; this module is the host-side mirror:

  detr_proposals(out_dict, args)                     the inline block, as a function (returns a `Proposals` CSR view)
  match_object_categories(...)                       reference name / argument lists / (None, None, None) behaviour
  match_target_sgd(rank, relationships, ...)         reference name / argument lists / per-image None behaviour
  sgdet_batch(...) / sgcls_batch(...)                device window (pipeline.DeviceBatch) ready for RelationPipeline.step

No CPU fallback: everything raises off-GPU.
"""
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import ops, tables
from .pipeline import DeviceBatch


def _const(name, fn, device, dtype):
    cache = _const.cache.setdefault(str(device), {})
    if name not in cache:
        cache[name] = torch.from_numpy(np.ascontiguousarray(fn())).to(dtype).to(device)
    return cache[name]


_const.cache = {}


@dataclass
class Proposals:
    """CSR view of the proposals of one window (device tensors; `offsets_host` is the one D2H read)."""
    n: int
    offsets: torch.Tensor            # int32 [B+1]
    offsets_host: np.ndarray
    cats: torch.Tensor               # int32 [n]   frequency-ordered object labels
    conf: torch.Tensor               # f32   [n]   label confidences
    box_f: Optional[torch.Tensor]    # f32   [n,4] (x1,x2,y1,y2) on the feature grid (None for SGCLS matches)
    box_i: torch.Tensor              # int32 [n,4] int()-truncated boxes (the form the pair/match kernels take)
    supers: Optional[torch.Tensor]   # int8  [n,4] super-categories, -1 padded
    box_img: torch.Tensor            # int32 [n]
    src: Optional[torch.Tensor] = None   # SGCLS: GT box index each row was matched for

    def _split(self, t):
        o = self.offsets_host
        return [t[int(o[i]):int(o[i + 1])] for i in range(len(o) - 1)]

    def to_lists(self):
        """The reference's variables after evaluate.py:370: (categories_pred, cat_pred_confidence, bbox_pred,
        super_categories_pred) as per-image lists."""
        sup = None
        if self.supers is not None:
            sup = [[row[row >= 0].long() for row in img] for img in self._split(self.supers)]
        boxes = self._split(self.box_f if self.box_f is not None else self.box_i)
        return [c.long() for c in self._split(self.cats)], self._split(self.conf), boxes, sup


def detr_proposals(out_dict, args=None, num_classes=None, topk_cat=None, feature_size=None, nms=None) -> Proposals:
    """evaluate.py:311-370: `out_dict['pred_logits']` [B,Q,num_classes+1], `out_dict['pred_boxes']` [B,Q,4] (cxcywh, 0..1)
    -> labelled, NMS-ed proposals.  Parameters default to args['models'][...] (config.yaml:32,40,41) or VG's values."""
    m = (args or {}).get("models", {}) if args is not None else {}
    num_classes = int(num_classes if num_classes is not None else m.get("num_classes", tables.NUM_OBJ))
    topk_cat = int(topk_cat if topk_cat is not None else m.get("topk_cat", 2))
    feature_size = int(feature_size if feature_size is not None else m.get("feature_size", 32))
    nms = float(nms if nms is not None else m.get("nms", 0.5))
    logits, boxes = out_dict["pred_logits"], out_dict["pred_boxes"]
    if not logits.is_cuda:
        raise RuntimeError("hiercom_b200: detr_proposals needs CUDA tensors (no CPU fallback)")
    dev = logits.device
    label_map = _const("alp2fre", tables.alp2fre, dev, torch.int32)
    s2s = _const("sub2super", tables.sub2super_table, dev, torch.int8)
    d = ops.detr_proposals(logits, boxes, label_map, s2s, num_classes, topk_cat, feature_size, nms)
    return Proposals(d["n"], d["offsets"], d["offsets_host"], d["cats"], d["conf"], d["box_f"], d["box_i"], d["supers"], d["box_img"])


def _csr(list_of_tensors, dtype, device, width=None):
    counts = [int(t.shape[0]) if hasattr(t, "shape") else len(t) for t in list_of_tensors]
    off = torch.as_tensor(np.concatenate(([0], np.cumsum(counts))), dtype=torch.int32)
    parts = [torch.as_tensor(t).reshape((-1,) if width is None else (-1, width)) for t in list_of_tensors]
    flat = torch.cat(parts) if parts else torch.zeros((0,) if width is None else (0, width))
    if flat.is_floating_point() and dtype in (torch.int32, torch.int64):
        flat = flat.trunc()
    return flat.to(device=device, dtype=dtype).contiguous(), off.to(device)


def match_object_categories(categories_pred, cat_pred_confidence, bbox_pred, bbox_target, device=None):
    """utils.py:376-422, reference argument/return shapes: per-image lists in, `(categories_pred_matched,
    cat_pred_confidence_matched, bbox_target_matched)` per-image lists out (rows of one image stay on the device),
    or `(None, None, None)` when the batch sizes differ or an image with GT boxes has fewer than two proposals."""
    if len(bbox_target) != len(bbox_pred):
        return None, None, None
    m = match_object_categories_csr(categories_pred, cat_pred_confidence, bbox_pred, bbox_target, device)
    if m is None:
        return None, None, None
    cats = [list(c.long()) for c in m._split(m.cats)]
    conf = [list(c) for c in m._split(m.conf)]
    return cats, conf, m._split(m.box_i)


def match_object_categories_csr(categories_pred, cat_pred_confidence, bbox_pred, bbox_target, device=None) -> Optional[Proposals]:
    if isinstance(categories_pred, Proposals):
        p = categories_pred
        dev = p.cats.device
        prop = (p.cats, p.conf, p.box_f, p.offsets)
    else:
        dev = torch.device(device) if device is not None else next((torch.as_tensor(t).device for t in bbox_pred if torch.is_tensor(t) and t.is_cuda),
                                                                   torch.device("cuda", torch.cuda.current_device()))
        cats, off = _csr(categories_pred, torch.int32, dev)
        conf, _ = _csr(cat_pred_confidence, torch.float32, dev)
        box, _ = _csr(bbox_pred, torch.float32, dev, 4)
        prop = (cats, conf, box, off)
    gt_box, gt_off = (bbox_target if isinstance(bbox_target, tuple) else _csr(bbox_target, torch.int32, dev, 4))
    s2s = _const("sub2super", tables.sub2super_table, dev, torch.int8)
    d = ops.match_object_categories(prop[0], prop[1], prop[2], prop[3], gt_box, gt_off, s2s)
    if d is None:
        return None
    return Proposals(d["n"], d["offsets"], d["offsets_host"], d["cats"], d["conf"], None, d["box_i"], d["supers"], d["box_img"], d["src"])


def pack_relationships(relationships, subj_or_obj, device):
    """Per-image lists of per-row tensors (dataloader.py:159-165) -> packed triangle arrays (rel_tri int32, dir_tri int8,
    tri_offsets int32 [B+1]) on the device; t = g(g-1)/2 + e."""
    rel = [torch.cat([torch.as_tensor(r).reshape(-1) for r in img]) if len(img) else torch.zeros(0) for img in relationships]
    dr = [torch.cat([torch.as_tensor(r).reshape(-1) for r in img]) if len(img) else torch.zeros(0) for img in subj_or_obj]
    off = np.concatenate(([0], np.cumsum([int(r.numel()) for r in rel]))).astype(np.int32)
    rel_tri = (torch.cat(rel) if rel else torch.zeros(0)).to(torch.int32).to(device)
    dir_tri = (torch.cat(dr) if dr else torch.zeros(0)).to(torch.int8).to(device)
    return rel_tri, dir_tri, torch.from_numpy(off).to(device)


def match_target_sgd(rank, relationships, subj_or_obj, categories_target, bbox_target):
    """utils.py:294-352, reference argument/return shapes: five per-image lists `(cat_subject_target, cat_object_target,
    bbox_subject_target, bbox_object_target, relation_target)`; an image without targets gets None entries."""
    dev = torch.device("cuda", rank) if isinstance(rank, int) else torch.device(rank)
    rel_tri, dir_tri, tri_off = pack_relationships(relationships, subj_or_obj, dev)
    counts = [int(torch.as_tensor(c).shape[0]) for c in categories_target]
    box_off = torch.as_tensor(np.concatenate(([0], np.cumsum(counts))), dtype=torch.int32, device=dev)
    cats = torch.cat([torch.as_tensor(c).reshape(-1) for c in categories_target]).to(dev)
    boxes = torch.cat([torch.as_tensor(b).reshape(-1, 4) for b in bbox_target]).to(dev)
    gt_off, label, sub, obj = ops.targets_flat(dir_tri, rel_tri, tri_off, box_off)
    o = gt_off.cpu().numpy()
    out = ([], [], [], [], [])
    for i in range(len(counts)):
        a, b = int(o[i]), int(o[i + 1])
        if a == b:
            for lst in out:
                lst.append(None)
            continue
        s, t = sub[a:b].long(), obj[a:b].long()
        for lst, v in zip(out, (cats[s], cats[t], boxes[s].view(-1, 4), boxes[t].view(-1, 4), label[a:b].to(torch.as_tensor(relationships[i][0]).dtype))):
            lst.append(v)
    return out


# ---------------------------------------------------------------------------------------------------- device windows
def _window(prop: Proposals, feat, depth, gt_cats, gt_boxes, gt_box_off, rel_tri, dir_tri, tri_off, skip_mode, group_size):
    dev = prop.cats.device
    counts = np.diff(prop.offsets_host).astype(np.int64)
    tri = counts * (counts - 1) // 2
    n_img = len(counts)
    group_id, n_groups = None, 0
    if skip_mode == "batch":
        gs = group_size or n_img
        gid = (np.arange(n_img) // gs).astype(np.int32)
        n_groups = int(gid.max()) + 1
        group_id = torch.from_numpy(gid).to(dev)
    elif skip_mode != "per_image":
        raise ValueError("skip_mode must be 'batch' or 'per_image'")
    gt_off, label, sub, obj = ops.targets_flat(dir_tri, rel_tri, tri_off, gt_box_off)
    gt = dict(offsets=gt_off, label=label, sub=sub, obj=obj, cat=gt_cats, box=gt_boxes)
    tri_offsets = torch.from_numpy(np.concatenate(([0], np.cumsum(tri))).astype(np.int32)).to(dev)
    return DeviceBatch(feat, depth, prop.box_i, prop.offsets, prop.box_img, prop.cats, prop.supers, tri_offsets, None, None, group_id,
                       n_groups, int(tri.max()) if n_img else 0, int((counts * (counts - 1)).sum()), 0, conf=prop.conf, gt=gt,
                       box_offsets_host=prop.offsets_host)


def sgdet_batch(out_dict, feat, depth, gt_cats, gt_boxes, gt_box_offsets, rel_tri, dir_tri, tri_offsets, args=None,
                skip_mode="per_image", group_size=None) -> DeviceBatch:
    """SGDET window (evaluate.py:304-446 minus the head): DETR outputs + GT tables already on the device ->
    `DeviceBatch` whose boxes/labels/confidences are the NMS-ed proposals and whose `gt` is match_target_sgd's table.
    gt_cats int32 [n_gt], gt_boxes int32 [n_gt,4], gt_box_offsets int32 [B+1]; rel_tri/dir_tri/tri_offsets as in
    `pack_relationships`."""
    prop = detr_proposals(out_dict, args)
    return _window(prop, feat, depth, gt_cats, gt_boxes, gt_box_offsets, rel_tri, dir_tri, tri_offsets, skip_mode, group_size)


def sgcls_batch(out_dict, feat, depth, gt_cats, gt_boxes, gt_box_offsets, rel_tri, dir_tri, tri_offsets, args=None,
                skip_mode="per_image", group_size=None) -> Optional[DeviceBatch]:
    """SGCLS window (evaluate.py:538-694 minus the head): GT boxes labelled from the proposals by
    match_object_categories; None when the reference would `continue` (evaluate.py:606-607)."""
    prop = detr_proposals(out_dict, args)
    m = match_object_categories_csr(prop, None, None, (gt_boxes, gt_box_offsets))
    if m is None:
        return None
    return _window(m, feat, depth, gt_cats, gt_boxes, gt_box_offsets, rel_tri, dir_tri, tri_offsets, skip_mode, group_size)
