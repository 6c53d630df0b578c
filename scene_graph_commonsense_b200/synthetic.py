"""Synthetic Visual-Genome-shaped inputs for the relation path.

Each image is generated from `torch.Generator(seed = base_seed * 1_000_003 + image_id)`, so any sharding of
image ids over ranks sees identical data.  Shapes and value ranges follow the reference's data pipeline:

  feat   f32 [256,32,32]  ~ N(0,1)           DETR encoder map, train_utils.py:9-18
  depth  f32 [1,32,32]    ~ U(0,1)           range-normalised MiDaS depth, dataset_utils.py:108
  bbox   int32 [N,4] (xmin,xmax,ymin,ymax) on the 32-grid, area-descending (dataset_utils.py:117,124;
         dataloader.py:121,129)
  categories int64 [N] in [0,150); super_categories: list of int64 tensors from sub2super_cat_dict.pt
  relationships[g-1][e], subj_or_obj[g-1][e] for e < g  (dataset_utils.py:159-184): predicate id in [0,50)
         or -1; direction 1 (g is subject), 0 (e is subject), -1 (no relation)

Nothing here touches the oracle or the reference; tests, bench and smoke share it.
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch

from . import tables

FEATURE_SIZE = 32
NUM_IMG_FEATURE = 256


@dataclass
class ImageSample:
    image_id: int
    feat: torch.Tensor
    depth: torch.Tensor
    bbox: torch.Tensor
    categories: torch.Tensor
    super_categories: List[torch.Tensor]
    relationships: List[torch.Tensor]
    subj_or_obj: List[torch.Tensor]
    # SGDET/SGCLS-style extras (None for PredCLS samples)
    bbox_pred: Optional[torch.Tensor] = None        # f32 [M,4] (x1,x2,y1,y2) on the 32-grid
    categories_pred: Optional[torch.Tensor] = None  # int64 [M]
    cat_conf_pred: Optional[torch.Tensor] = None    # f32 [M]
    super_categories_pred: Optional[List[torch.Tensor]] = field(default=None)


def _gen(base_seed, image_id):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(base_seed) * 1_000_003 + int(image_id))
    return g


def _randint(g, lo, hi, n):
    return torch.randint(lo, hi + 1, (n,), generator=g, dtype=torch.int64)


BOX_MODES = ("small", "vg", "full")


def make_boxes(g, n, mode="small"):
    """Boxes on the 32-grid, largest first (the reference's loader sorts by area).  mode "small": SURVEY §8d's distribution (side
    4..15, 9.8 % of the pooled conv3_1 cells shared by a pair); "vg": sides U{8..32} placed uniformly inside the grid (the wide
    spread of Visual Genome boxes: parts next to whole-scene regions); "full": every box is the whole grid (nothing to skip)."""
    if mode == "full":
        return torch.tensor([[0, FEATURE_SIZE, 0, FEATURE_SIZE]] * n, dtype=torch.int32).reshape(n, 4)
    if mode == "vg":
        w = _randint(g, 8, FEATURE_SIZE, n)
        h = _randint(g, 8, FEATURE_SIZE, n)
        x0 = (torch.rand(n, generator=g) * (FEATURE_SIZE - w + 1).to(torch.float32)).to(torch.int64)
        y0 = (torch.rand(n, generator=g) * (FEATURE_SIZE - h + 1).to(torch.float32)).to(torch.int64)
    elif mode == "small":
        x0 = _randint(g, 0, 23, n)
        y0 = _randint(g, 0, 23, n)
        w = _randint(g, 4, 15, n)
        h = _randint(g, 4, 15, n)
    else:
        raise ValueError("box mode must be one of %s" % (BOX_MODES,))
    x1 = torch.clamp(x0 + w, max=FEATURE_SIZE)
    y1 = torch.clamp(y0 + h, max=FEATURE_SIZE)
    area = (x1 - x0) * (y1 - y0)
    order = torch.sort(area, descending=True, stable=True)[1]
    box = torch.stack((x0, x1, y0, y1), dim=1)[order]
    return box.to(torch.int32)


def favoured_pred(image_id, sub, obj, num_pred=tables.NUM_PRED):
    """A per-directed-pair predicate that `pair_scores` boosts and `make_image` prefers as the GT label, so that
    table-driven tests see real hits instead of chance-level matches."""
    return (int(image_id) * 31 + int(sub) * 17 + int(obj) * 7 + 3) % num_pred


def make_image(image_id, num_boxes, base_seed=0, p_rel=0.3, with_maps=True, p_fav=0.6, box_mode="small"):
    g = _gen(base_seed, image_id)
    if with_maps:
        feat = torch.randn(NUM_IMG_FEATURE, FEATURE_SIZE, FEATURE_SIZE, generator=g)
        depth = torch.rand(1, FEATURE_SIZE, FEATURE_SIZE, generator=g)
    else:
        feat = torch.zeros(0)
        depth = torch.zeros(0)
    bbox = make_boxes(g, num_boxes, box_mode)
    cats = _randint(g, 0, tables.NUM_OBJ - 1, num_boxes)
    s2s = tables.sub2super_table()
    supers = [torch.as_tensor([int(v) for v in s2s[int(c)] if v >= 0], dtype=torch.int64) for c in cats]
    probs = torch.as_tensor(tables.vg_predicate_counts(), dtype=torch.float64)
    probs = probs / probs.sum()
    rels, dirs = [], []
    for gi in range(1, num_boxes):
        has = torch.rand(gi, generator=g) < p_rel
        pred = torch.multinomial(probs, gi, replacement=True, generator=g)
        d = (torch.rand(gi, generator=g) < 0.5).to(torch.float32)
        use_fav = torch.rand(gi, generator=g) < p_fav
        fav = torch.as_tensor([favoured_pred(image_id, gi, e) if d[e] == 1 else favoured_pred(image_id, e, gi)
                               for e in range(gi)], dtype=torch.int64)
        pred = torch.where(use_fav, fav, pred)
        rels.append(torch.where(has, pred, torch.full_like(pred, -1)))
        dirs.append(torch.where(has, d, torch.full_like(d, -1.0)))
    return ImageSample(image_id, feat, depth, bbox, cats, supers, rels, dirs)


def with_categories(sample, cats):
    """The same sample with its object categories (and the super-category lists that follow from them) replaced: lets a test
    force classes with 1, 2 and 3 super-classes (`sub2super_cat_dict.pt`: 118 / 19 / 13 of the 150) into a small image."""
    s2s = tables.sub2super_table()
    sample.categories = torch.as_tensor([int(c) for c in cats], dtype=torch.int64)
    sample.super_categories = [torch.as_tensor([int(v) for v in s2s[int(c)] if v >= 0], dtype=torch.int64) for c in cats]
    return sample


def assign_gt_from_scores(sample, relation_of, splits=(15, 11, 24), p_model=0.8, p_top=0.5, base_seed=0):
    """Re-labels the GT relations of `sample` from a model's own scores so that Recall@K is discriminating (hits exist, and a wrong
    score or ranking changes them): a related unordered pair keeps its direction; with probability `p_model` its predicate
    becomes one of the three per-super-category argmaxes of `relation_of(sub, obj)` (f32[sum(splits)] log-joints of the DIRECTED
    pair) - the best of the three with probability `p_top`, else a uniformly chosen one - otherwise it keeps its random label.
    Deterministic in (base_seed, image_id); modifies and returns the sample."""
    g = _gen(base_seed + 104729, sample.image_id)
    offs = np.concatenate(([0], np.cumsum(splits)))
    for gi in range(1, len(sample.categories)):
        for e in range(gi):
            d = int(sample.subj_or_obj[gi - 1][e])
            u = torch.rand(3, generator=g)
            if d < 0 or float(u[0]) >= p_model:
                continue
            sub, obj = (gi, e) if d == 1 else (e, gi)
            r = np.asarray(relation_of(sub, obj), dtype=np.float32)
            arg = [int(offs[k] + np.argmax(r[offs[k]:offs[k + 1]])) for k in range(len(splits))]
            if float(u[1]) < p_top:
                lab = max(arg, key=lambda a: (r[a], -a))
            else:
                lab = arg[min(int(float(u[2]) * len(arg)), len(arg) - 1)]
            sample.relationships[gi - 1][e] = lab
    return sample


def assign_gt_from_ranking(sample, ranked, p_keep=0.9, thin=0.1, base_seed=0):
    """Re-draws the GT relations of `sample` around a model's own RANKED triplets so that Recall@K lands mid-range (a wrong score,
    label, filter decision or rank anywhere in the path moves it): the relations `make_image` drew are thinned (each kept with
    probability `thin`: p_rel 0.3 -> 0.03), then every ranked triplet `(sub, obj, label)` (image-local box rows, best first)
    becomes GT with probability `p_keep` unless its unordered pair was already taken by a better rank (the reference's GT holds
    one relation per unordered pair, dataset_utils.py:159-184).  Deterministic in (base_seed, image_id); modifies and returns the sample."""
    g = _gen(base_seed + 104729, sample.image_id)
    for gi in range(1, len(sample.categories)):
        drop = torch.rand(gi, generator=g) >= thin
        sample.relationships[gi - 1] = torch.where(drop, torch.full_like(sample.relationships[gi - 1], -1), sample.relationships[gi - 1])
        sample.subj_or_obj[gi - 1] = torch.where(drop, torch.full_like(sample.subj_or_obj[gi - 1], -1.0), sample.subj_or_obj[gi - 1])
    taken = set()
    u = torch.rand(max(len(ranked), 1), generator=g)
    for j, (sub, obj, lab) in enumerate(ranked):
        hi, lo = (sub, obj) if sub > obj else (obj, sub)
        if hi == lo or (hi, lo) in taken or float(u[j]) >= p_keep:
            continue
        taken.add((hi, lo))
        sample.relationships[hi - 1][lo] = int(lab)
        sample.subj_or_obj[hi - 1][lo] = 1.0 if sub == hi else 0.0
    return sample


WEIGHT_PRESETS = {
    # name: (trunk_gain, logit_gain).  "init" = nn default init; "trained" = the round-1 trained-scale variant (head layers
    # only: logit std 1.1, `pred` <= 0.09); "sharp" = He-gain trunk (`pred` O(1), max 1.8) and logit std 3.3:
    # the top joint probability has median 0.46 and bf16 operand rounding reaches the probabilities at full scale.
    "init": (1.0, 1.0), "trained": (1.0, 40.0), "sharp": (6.0 ** 0.5, 16.0),
}


def preset_state_dict(name, seed=0, **kw):
    tg, lg = WEIGHT_PRESETS[name]
    return head_state_dict(seed=seed, logit_gain=lg, trunk_gain=tg, **kw)


def make_sgdet_image(image_id, num_gt, num_prop, base_seed=0, p_rel=0.3, with_maps=True, box_mode="small"):
    """SGDET/SGCLS-shaped sample: `num_prop` float proposals; the first `num_gt` are
    jittered copies of the GT boxes with the GT label (so matches exist), the rest are random."""
    s = make_image(image_id, num_gt, base_seed, p_rel, with_maps, box_mode=box_mode)
    g = _gen(base_seed + 7919, image_id)
    extra = make_boxes(g, max(num_prop - num_gt, 0), box_mode).to(torch.float32)
    base = torch.cat((s.bbox.to(torch.float32), extra), dim=0)[:num_prop]
    jitter = torch.rand(base.shape, generator=g)
    s.bbox_pred = torch.clamp(base + jitter, 0.0, float(FEATURE_SIZE))
    cat_extra = _randint(g, 0, tables.NUM_OBJ - 1, max(num_prop - num_gt, 0))
    s.categories_pred = torch.cat((s.categories, cat_extra))[:num_prop]
    s.cat_conf_pred = torch.rand(num_prop, generator=g)
    s2s = tables.sub2super_table()
    s.super_categories_pred = [torch.as_tensor([int(v) for v in s2s[int(c)] if v >= 0], dtype=torch.int64)
                               for c in s.categories_pred]
    return s


def make_detr_outputs(samples, num_queries=100, base_seed=0, p_noobj=0.35, p_dup=0.3, p_second_noobj=0.25,
                      num_classes=tables.NUM_OBJ, logit_gain=6.0):
    """Synthetic DETR head outputs for a list of GT samples (the inputs of the SGDET/SGCLS proposal front-end,
    evaluate.py:304-332): `pred_logits` f32 [B,Q,num_classes+1] in DETR's ALPHABETICAL label space (last = no object) and
    `pred_boxes` f32 [B,Q,4] = (cx,cy,w,h) in 0..1.  Queries are jittered copies of the image's GT boxes with the GT
    label (mapped back to the alphabetical id) and a random runner-up label, duplicates of an earlier query (so the
    per-class NMS has work), or "no object" queries; the runner-up of some object queries is "no object" (exercises the
    label != 150 mask after the top-2 expansion, evaluate.py:323,341-345)."""
    fre2alp = np.argsort(tables.alp2fre())          # inverse permutation: frequency id -> alphabetical id
    logits = torch.empty(len(samples), num_queries, num_classes + 1)
    boxes = torch.empty(len(samples), num_queries, 4)
    for b, s in enumerate(samples):
        g = _gen(base_seed + 104729, s.image_id)
        n_gt = s.bbox.shape[0]
        lg = torch.randn(num_queries, num_classes + 1, generator=g)
        bx = torch.empty(num_queries, 4)
        for q in range(num_queries):
            r = float(torch.rand(1, generator=g))
            if q > 0 and r < p_dup:                                   # near-duplicate of an earlier query
                src = int(torch.randint(0, q, (1,), generator=g))
                bx[q] = torch.clamp(bx[src] + 0.01 * torch.randn(4, generator=g), 0.02, 0.98)
                lg[q] = lg[src] + 0.3 * torch.randn(num_classes + 1, generator=g)
                continue
            if n_gt > 0 and r < 1.0 - p_noobj:
                k = int(torch.randint(0, n_gt, (1,), generator=g))
                x0, x1, y0, y1 = [float(v) for v in s.bbox[k]]
                box = torch.tensor([(x0 + x1) / 2, (y0 + y1) / 2, x1 - x0, y1 - y0]) / FEATURE_SIZE
                top = int(fre2alp[int(s.categories[k])])
            else:
                box = torch.rand(4, generator=g) * torch.tensor([1.0, 1.0, 0.5, 0.5])
                top = num_classes if r >= 1.0 - p_noobj else int(torch.randint(0, num_classes, (1,), generator=g))
            bx[q] = torch.clamp(box + 0.02 * torch.randn(4, generator=g), 0.0, 1.0)
            lg[q, top] += logit_gain
            second = num_classes if float(torch.rand(1, generator=g)) < p_second_noobj else int(
                torch.randint(0, num_classes, (1,), generator=g))
            if second != top:
                lg[q, second] += 0.6 * logit_gain
        logits[b], boxes[b] = lg, bx
    return logits, boxes


def make_batch(image_ids, num_boxes, base_seed=0, p_rel=0.3, with_maps=True, box_mode="small"):
    """`num_boxes` may be an int or a per-image sequence (ragged batches)."""
    if isinstance(num_boxes, int):
        num_boxes = [num_boxes] * len(image_ids)
    return [make_image(i, n, base_seed, p_rel, with_maps, box_mode=box_mode) for i, n in zip(image_ids, num_boxes)]


# ----------------------------------------------------------------------------------------------
# deterministic weights for the relation head (state-dict layout of model.py:105-133)


def head_state_dict(seed=0, input_dim=128, feature_size=32, num_classes=150, num_super_classes=17,
                    splits=(15, 11, 24), logit_gain=1.0, vg=True, flat=False, dtype=torch.float32, trunk_gain=1.0):
    """Weights drawn like nn.Conv2d/nn.Linear default init (U(-1/sqrt(fan_in), 1/sqrt(fan_in))) from one seeded
    generator, in a fixed key order, so the reference module, the oracle and the CUDA path all load the same
    values.  `logit_gain` scales the final classification layers ("trained-scale" variant).  `trunk_gain`
    scales the weights of conv2_1 / conv3_1 / fc1 / fc2 (default init shrinks the signal by ~2x per ReLU layer, leaving
    `pred` at 0.08; sqrt(6) is the variance-preserving He gain, so `pred` is O(1) and bf16 rounding in the trunk reaches
    the logits at full scale).  The random draws are identical for every gain."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1_000_000_007 + int(seed))

    def u(shape, fan_in, gain=1.0):
        b = gain / (fan_in ** 0.5)
        return ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * b).to(dtype)

    c = input_dim
    cin = 2 * c + 1
    fc1_in = 8 * c * (feature_size // 4) ** 2
    fc2_in = 4096 + 2 * (num_classes + (num_super_classes if vg else 0))
    sd = {}
    for name in ("conv1_1", "conv1_2"):
        sd[name + ".weight"] = u((c, cin, 1, 1), cin)
        sd[name + ".bias"] = u((c,), cin)
    sd["conv2_1.weight"] = u((4 * c, 2 * c, 3, 3), 2 * c * 9, trunk_gain)
    sd["conv2_1.bias"] = u((4 * c,), 2 * c * 9)
    sd["conv3_1.weight"] = u((8 * c, 4 * c, 3, 3), 4 * c * 9, trunk_gain)
    sd["conv3_1.bias"] = u((8 * c,), 4 * c * 9)
    sd["fc1.weight"] = u((4096, fc1_in), fc1_in, trunk_gain)
    sd["fc1.bias"] = u((4096,), fc1_in)
    sd["fc2.weight"] = u((512, fc2_in), fc2_in, trunk_gain)
    sd["fc2.bias"] = u((512,), fc2_in)
    if flat:
        sd["fc3.weight"] = u((sum(splits), 512), 512, logit_gain)
        sd["fc3.bias"] = u((sum(splits),), 512, logit_gain)
    else:
        for i, n in enumerate(splits):
            sd["fc3_%d.weight" % (i + 1)] = u((n, 512), 512, logit_gain)
            sd["fc3_%d.bias" % (i + 1)] = u((n,), 512, logit_gain)
    sd["fc4.weight"] = u((1, 512), 512, logit_gain)
    sd["fc4.bias"] = u((1,), 512, logit_gain)
    if not flat:
        sd["fc5.weight"] = u((3, 512), 512, logit_gain)
        sd["fc5.bias"] = u((3,), 512, logit_gain)
    return sd


def reference_args(run_mode="eval_cs", hierar=True, splits=(15, 11, 24), dataset="vg"):
    """The slice of the reference's `args` dict (config.yaml + main.py:49-85 overrides) the path reads."""
    return {
        "dataset": {"dataset": dataset, "supcat_clustering": "motif",
                    "train_triplets": "datasets/vg_scene_graph_annot/train_triplets.pt",
                    "test_triplets": "datasets/vg_scene_graph_annot/test_triplets.pt",
                    "zero_shot_triplets": "datasets/vg_scene_graph_annot/zero_shot_triplets.pt",
                    "sub2super_cat_dict": "datasets/vg_scene_graph_annot/sub2super_cat_dict.pt"},
        "models": {"hierarchical_pred": hierar, "feature_size": FEATURE_SIZE, "image_size": 1024,
                   "num_img_feature": NUM_IMG_FEATURE, "hidden_dim": 128, "num_classes": tables.NUM_OBJ,
                   "num_relations": sum(splits), "num_super_classes": tables.NUM_SUPER_OBJ,
                   "num_geometric": splits[0], "num_possessive": splits[1], "num_semantic": splits[2],
                   "llm_model": "gpt3.5", "topk_cat": 2, "nms": 0.5, "use_depth": True},
        "training": {"run_mode": run_mode, "eval_mode": "pc", "batch_size": 12, "eval_freq_test": 1,
                     "print_freq_test": 20, "save_vis_results": False, "result_path": "results/",
                     "checkpoint_path": "checkpoints/", "test_epoch": 2},
    }


# ----------------------------------------------------------------------------------------------
# table-driven scores for integer-stage tests and the evaluation micro-benchmark


def pair_scores(image_id, sub, obj, splits=(15, 11, 24), base_seed=0, gain=3.0, tie_quantum=None):
    """Deterministic fake head output for directed pair (image_id, sub -> obj):
    `(relation f32[sum(splits)], super f32[3], connectivity f32[1])` with the reference's hierarchical log-joint
    structure (model.py:176-184).  `tie_quantum` rounds the logits to a grid so ties are frequent (H1 tests)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(((int(base_seed) * 7919 + int(image_id)) * 1009 + int(sub)) * 1013 + int(obj) + 17)
    n = sum(splits)
    logits = torch.randn(n + 4, generator=g) * gain
    logits[favoured_pred(image_id, sub, obj, n)] += 2.0 * gain
    if tie_quantum:
        logits = torch.round(logits / tie_quantum) * tie_quantum
    sup = torch.log_softmax(logits[n:n + 3], dim=0)
    parts, a = [], 0
    for k, w in enumerate(splits):
        parts.append(torch.log_softmax(logits[a:a + w], dim=0) + sup[k])
        a += w
    return torch.cat(parts), sup, logits[n + 3:n + 4]


def batch_score_fn(batch, splits=(15, 11, 24), base_seed=0, gain=3.0, tie_quantum=None):
    """Adapter with the oracle replay's `head_fn` signature: ctx = (keep, sub_idx, obj_idx)."""
    def fn(h_sub, h_obj, c1, c2, s1, s2, ctx):
        keep, sub, obj = ctx
        rows = [pair_scores(batch[int(i)].image_id, sub, obj, splits, base_seed, gain, tie_quantum) for i in keep]
        return (torch.stack([r[0] for r in rows]), torch.stack([r[1] for r in rows]),
                torch.stack([r[2] for r in rows]))
    return fn


def synthetic_cs_keys(seed=0, frac_aligned=0.5, frac_violated=0.1):
    """Dense stand-ins for the commonsense sets (packed keys), so the filter is exercised on both outcomes at test
    sizes; the shipped sets cover only 1.9 % / 0.14 % of the key space.  The two sets overlap on purpose (the real
    ones share 403 keys)."""
    rng = np.random.default_rng(seed)
    r = rng.random(tables.TRIPLET_SPACE)
    aligned = np.nonzero(r < frac_aligned)[0].astype(np.int64)
    r2 = rng.random(tables.TRIPLET_SPACE)
    violated = np.nonzero(r2 < frac_violated)[0].astype(np.int64)
    return aligned, violated


# ----------------------------------------------------------------------------------------------
# Scene-Graph-Benchmark-shaped inputs (config 5: PredCLS Motifs tail, 512-d context, 4096-d union features, 151/51 classes)


def sgb_state_dict(seed=0, hidden=512, pooling=4096, num_obj=151, num_rel=51, logit_gain=4.0):
    g = torch.Generator(device="cpu")
    g.manual_seed(2_000_003 + int(seed))

    def n(shape, std):
        return torch.randn(shape, generator=g) * std
    sd = {"post_emb.weight": n((2 * hidden, hidden), 10.0 * (1.0 / hidden) ** 0.5 / 8), "post_emb.bias": torch.zeros(2 * hidden),
          "post_cat.weight": n((pooling, 2 * hidden), (2.0 / (pooling + 2 * hidden)) ** 0.5), "post_cat.bias": n((pooling,), 0.01)}
    for name, rows in (("fc3_1", 15), ("fc3_2", 11), ("fc3_3", 24), ("fc5", 4)):
        sd[name + ".weight"] = n((rows, pooling), logit_gain * (2.0 / (pooling + rows)) ** 0.5)
        sd[name + ".bias"] = n((rows,), 0.1)
    sd["freq_bias"] = torch.log_softmax(n((num_obj * num_obj, num_rel), 1.5), dim=1)
    return sd


def make_sgb_batch(num_objs, seed=0, hidden=512, pooling=4096, num_obj_cls=151):
    g = torch.Generator(device="cpu")
    g.manual_seed(3_000_017 + int(seed))
    total = int(sum(num_objs))
    pairs = int(sum(n * (n - 1) for n in num_objs))
    boxes = []
    for n in num_objs:
        xy = torch.rand(n, 2, generator=g) * 400
        wh = torch.rand(n, 2, generator=g) * 200 + 20
        boxes.append(torch.cat((xy, xy + wh), dim=1).float())
    return dict(num_objs=list(num_objs), edge_ctx=torch.randn(total, hidden, generator=g),
                obj_labels=torch.randint(1, num_obj_cls, (total,), generator=g),
                obj_logits=torch.randn(total, num_obj_cls, generator=g) * 2.0,
                union_features=torch.relu(torch.randn(pairs, pooling, generator=g)), boxes=boxes)


def sgb_validator(combined_obj_label, rel_labels, image=None, boxlist=None):
    """Deterministic stand-in for CommonsenseValidator.query: reject (-1) when (subject + predicate + object) % 3 == 0."""
    s = combined_obj_label[:, 0].long() + combined_obj_label[:, 1].long() + rel_labels.long()
    return torch.where(s % 3 == 0, torch.full_like(s, -1), torch.ones_like(s))
