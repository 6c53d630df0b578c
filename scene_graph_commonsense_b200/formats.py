"""On-disk formats either side of the path.

* per-image annotation files of the reference (`<image>_annotations.pkl`, written with torch.save by
  dataset_utils.py:186-200) -> the transforms the reference's Dataset applies before the hot path sees them
  (dataloader.py:117-149) -> one packed CSR tensor file per evaluation window that `pipeline.HostBatch` loads directly
  (no per-image Python objects on the hot path);
* commonsense dict `.pt` files <-> packed key arrays / bitmaps (see tables.dict_to_keys, ops.cs_bitmap_build);
* checkpoints: `module.` prefix stripping (utils.py:207-214) and the save/load file-name mismatch of the reference
  (saved as `..._motif2_0.pth`, train_test.py:311-315, but loaded as `..._motif_2_0.pth`, train_test.py:84-88,
  evaluate.py:66-69).
"""
import os
from types import SimpleNamespace

import numpy as np
import torch

from . import tables
from .model import strip_module_prefix

# dataset_utils.py:647-650 relation_class_freq2scat: frequency-ordered predicate id -> super-category-ordered id
RELATION_FREQ2SCAT = [11, 18, 8, 20, 23, 10, 25, 0, 34, 6, 14, 44, 24, 45, 9, 26, 5, 33, 13, 16,
                      42, 27, 30, 48, 41, 29, 35, 3, 49, 4, 7, 15, 39, 2, 36, 17, 40, 22, 19, 28,
                      38, 43, 21, 1, 31, 46, 12, 37, 32, 47, -1]


def sample_from_annotation(annot, image_id=0, feat=None, rel_reorder=None, max_objects=20):
    """dataloader.py:117-149 on one loaded annotation dict: object-count gate (2..20 objects), `bbox.int()`, the
    wears->wearing merge (`rel[rel == 12] = 4`) and the predicate reorder (motif clustering by default; pass another
    50+1 entry table for gpt2 / bert / clip, dataset_utils.py:760-788).  Returns None for images the reference skips."""
    cats = torch.as_tensor(annot['categories'])
    if cats.shape[0] <= 1 or cats.shape[0] > max_objects:                     # dataloader.py:119-120
        return None
    table = torch.as_tensor(RELATION_FREQ2SCAT if rel_reorder is None else rel_reorder, dtype=torch.int64)
    rels = []
    for rel in annot['relationships']:
        rel = torch.as_tensor(rel, dtype=torch.int64).clone()
        rel[rel == 12] = 4                                                    # dataloader.py:144
        rels.append(table[rel])                                               # index -1 hits the trailing -1 entry
    depth = annot.get('image_depth')
    return SimpleNamespace(image_id=image_id, feat=feat, depth=depth, bbox=torch.as_tensor(annot['bbox']).int(),
                           categories=cats.to(torch.int64), super_categories=[torch.as_tensor(s, dtype=torch.int64) for s in annot['super_categories']],
                           relationships=rels, subj_or_obj=[torch.as_tensor(s, dtype=torch.float32) for s in annot['subj_or_obj']])


def load_annotation_file(path, **kw):
    return sample_from_annotation(torch.load(path, weights_only=False), **kw)


def pack_window(samples):
    """List of samples (as above; `feat` may be None) -> dict of flat numpy arrays: the CSR layout of pipeline.HostBatch
    (box_offsets, tri_offsets, boxes, cats, supers, box_img, rel_tri, dir_tri) plus depth maps."""
    counts = np.array([len(s.categories) for s in samples], dtype=np.int64)
    tri = counts * (counts - 1) // 2
    supers = -np.ones((int(counts.sum()), 4), dtype=np.int8)
    r = 0
    for s in samples:
        for sc in s.super_categories:
            v = np.asarray(sc, dtype=np.int64)[:4]
            supers[r, :len(v)] = v
            r += 1
    cat1 = lambda parts, dt: (np.concatenate(parts) if parts else np.zeros(0)).astype(dt)
    out = dict(
        box_offsets=np.concatenate(([0], np.cumsum(counts))).astype(np.int32),
        tri_offsets=np.concatenate(([0], np.cumsum(tri))).astype(np.int32),
        boxes=cat1([np.asarray(s.bbox).reshape(-1, 4) for s in samples], np.int32),
        cats=cat1([np.asarray(s.categories) for s in samples], np.int32),
        supers=supers,
        box_img=np.repeat(np.arange(len(samples), dtype=np.int32), counts),
        rel_tri=cat1([np.asarray(r_) for s in samples for r_ in s.relationships], np.int32),
        dir_tri=cat1([np.asarray(r_) for s in samples for r_ in s.subj_or_obj], np.int8),
        image_ids=np.array([s.image_id for s in samples], dtype=np.int64))
    if all(getattr(s, "depth", None) is not None for s in samples):
        out["depth"] = np.stack([np.asarray(s.depth, dtype=np.float32).reshape(1, 32, 32) for s in samples])
    return out


def save_window(path, packed):
    np.savez(path, **packed)


def load_window(path):
    return dict(np.load(path))


def host_batch_from_packed(packed, feat=None, skip_mode="batch", group_size=None, pinned=True):
    """Packed window (+ the DETR feature maps `feat` [B,256,32,32], produced by reference code) -> pipeline.HostBatch."""
    from .pipeline import HostBatch, footprint_cover_fraction, host_pair_offsets
    arrays = {k: packed[k] for k in ("box_offsets", "tri_offsets", "boxes", "cats", "supers", "box_img", "rel_tri", "dir_tri")}
    counts = np.diff(packed["box_offsets"]).astype(np.int64)
    n_img = len(counts)
    n_groups = 0
    if skip_mode == "batch":
        gs = group_size or n_img
        arrays["group_id"] = (np.arange(n_img) // gs).astype(np.int32)
        n_groups = int(arrays["group_id"].max()) + 1
    if feat is not None:
        arrays["feat"] = torch.as_tensor(feat, dtype=torch.float32)
        arrays["depth"] = torch.as_tensor(packed["depth"], dtype=torch.float32)
    tri = counts * (counts - 1) // 2
    meta = dict(n_groups=n_groups, max_tri=int(tri.max()) if n_img else 0, p_max=int((counts * (counts - 1)).sum()),
                cover_fraction=footprint_cover_fraction(arrays["boxes"], arrays["box_offsets"]),
                pair_offsets=host_pair_offsets(arrays["boxes"], arrays["box_offsets"], arrays.get("group_id")))
    return HostBatch(arrays, meta, pinned)


# ---------------------------------------------------------------------------------------------------- commonsense sets
def commonsense_pt_to_keys(path):
    """`triplets/commonsense_*_triplets.pt` (dict {(s,p,o): count}) or zero-shot list of 's_p_o' strings -> packed int64 keys."""
    return tables.dict_to_keys(torch.load(path, weights_only=False))


def keys_to_commonsense_dict(keys):
    """Packed keys -> the reference's dict format {(s,p,o): 1} (for writing a `.pt` the reference can load)."""
    keys = np.asarray(keys, dtype=np.int64)
    return {(int(k) // (tables.NUM_PRED * tables.NUM_OBJ), (int(k) // tables.NUM_OBJ) % tables.NUM_PRED, int(k) % tables.NUM_OBJ): 1 for k in keys}


# ---------------------------------------------------------------------------------------------------- checkpoints
def checkpoint_candidates(args, epoch, rank=0):
    """File names under which the reference may have stored / expects the relation-head checkpoint of `epoch`:
    the loader's spelling first (evaluate.py:66-69), then the trainer's (train_test.py:311-315, missing '_')."""
    hier = args['models']['hierarchical_pred']
    cs = args['training']['run_mode'] in ('train_cs', 'eval_cs')
    base = ('HierRelationModel' if hier else 'FlatRelationModel') + ('_CS' if cs else '_Baseline')
    clus = args['dataset']['supcat_clustering']
    root = args['training']['checkpoint_path']
    return [root + '%s_%s_%d_%d.pth' % (base, clus, epoch, rank), root + '%s_%s%d_%d.pth' % (base, clus, epoch, rank)]


def load_checkpoint(module, args=None, epoch=None, path=None, map_location="cpu"):
    """Load a reference checkpoint into a drop-in module, tolerating DDP's `module.` prefix and both file spellings."""
    paths = [path] if path is not None else checkpoint_candidates(args, epoch)
    for p in paths:
        if os.path.exists(p):
            sd = torch.load(p, map_location=map_location)
            module.load_state_dict(strip_module_prefix(sd))
            return p
    raise FileNotFoundError("no checkpoint found among: " + ", ".join(paths))
