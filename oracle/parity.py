"""TEST INFRASTRUCTURE - float-parity checker for the relation head at benchmark scale.  Not product code: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg only.

  stratified_pair_sample   picks directed pairs of a batch evenly from geometric strata (how many pooled conv3_1 cells both
                           boxes reach; boxes on the image border; degenerate boxes) so every code path of the shared-footprint
                           formulation (per-box maps only / one shared cell / many shared cells / zero-filled halo) is covered
  oracle_scores            the fp32 reference formulation (oracle.bayesian_relation_classifier == model.py:170-186) on those pairs
  operand_rounded_scores   the SAME formulation with every GEMM / convolution operand (weights and layer inputs) rounded to a
                           16-bit type, fp32 accumulation - the best any "bf16 in, fp32 accumulate" implementation can do.  It is
                           NOT the reference; it separates "error of the bf16 operand format" from "error of our kernels"
  sgb_tail_parity          config 5: a per-image random sample of the SGB tail's pairs through the fp32 restatement (sgb_oracle.predictor_tail)
  parity_stats             max |dP| on joint probabilities (north_star's 2e-3 bar), max relative log-prob error, and the fraction
                           of per-super-category argmax labels that differ
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import hiercom_oracle as O

SPLITS = (15, 11, 24)


def _cells(lo, hi):
    """Interval of pooled conv3_1 cells (8 per axis) whose value can depend on a box [lo, hi) of the 32-grid: dilate by the
    3x3 conv2_1, 2x2 pool, dilate by the 3x3 conv3_1, 2x2 pool (model.py:143-146)."""
    if hi <= lo:
        return 0, 0
    qlo, qhi = max(0, (lo - 1) >> 1), min(15, hi >> 1)
    qlo, qhi = max(0, qlo - 1), min(15, qhi + 1)
    return qlo >> 1, (qhi >> 1) + 1


def _clip_box(b, fs=32):
    x0, x1 = O._slice_bounds(b[0], b[1], fs)
    y0, y1 = O._slice_bounds(b[2], b[3], fs)
    return x0, x1, y0, y1


def shared_cell_count(box_s, box_o, fs=32):
    xs0, xs1, ys0, ys1 = _clip_box(box_s, fs)
    xo0, xo1, yo0, yo1 = _clip_box(box_o, fs)
    if xs1 <= xs0 or ys1 <= ys0 or xo1 <= xo0 or yo1 <= yo0:
        return 0
    ax, bx = _cells(xs0, xs1)
    cx, dx = _cells(xo0, xo1)
    ay, by = _cells(ys0, ys1)
    cy, dy = _cells(yo0, yo1)
    return max(0, min(bx, dx) - max(ax, cx)) * max(0, min(by, dy) - max(ay, cy))


def pair_stratum(box_s, box_o, fs=32):
    """0: no shared cell, 1: one or two shared cells, 2: 3..15, 3: >= 16 shared cells, 4: a box touches the image border
    (TMA zero-fill halo), 5: a degenerate (empty) box."""
    for b in (box_s, box_o):
        x0, x1, y0, y1 = _clip_box(b, fs)
        if x1 <= x0 or y1 <= y0:
            return 5
    n = shared_cell_count(box_s, box_o, fs)
    for b in (box_s, box_o):
        x0, x1, y0, y1 = _clip_box(b, fs)
        if n > 0 and (x0 == 0 or y0 == 0 or x1 == fs or y1 == fs):
            return 4
    return 0 if n == 0 else 1 if n <= 2 else 2 if n < 16 else 3


def stratified_pair_sample(boxes, sub, obj, n, seed=0, fs=32):
    """boxes int [nbox,4] (xmin,xmax,ymin,ymax); sub/obj int [P] global box rows of the directed pairs -> sorted indices of
    ~n pairs, drawn evenly from the non-empty strata (a short stratum is taken whole and its share goes to the others)."""
    boxes, sub, obj = np.asarray(boxes), np.asarray(sub), np.asarray(obj)
    strata = np.array([pair_stratum(boxes[s], boxes[o], fs) for s, o in zip(sub, obj)])
    rng = np.random.default_rng(seed)
    groups = [np.nonzero(strata == k)[0] for k in range(6)]
    groups = [g for g in groups if len(g)]
    picked, left = [], n
    for i, g in enumerate(sorted(groups, key=len)):
        share = left // (len(groups) - i)
        take = g if len(g) <= share else rng.choice(g, size=share, replace=False)
        picked.append(take)
        left -= len(take)
    idx = np.sort(np.concatenate(picked)) if picked else np.zeros(0, dtype=np.int64)
    return idx, strata[idx]


def _inputs(samples, pair_list, sgdet):
    hs, ho, c1, c2, s1, s2 = [], [], [], [], [], []
    for img, a, b in pair_list:
        s = samples[img]
        bb = s.bbox_pred if sgdet else s.bbox
        cats = s.categories_pred if sgdet else s.categories
        sup = s.super_categories_pred if sgdet else s.super_categories
        hs.append(O._masked_input(s, bb[a]))
        ho.append(O._masked_input(s, bb[b]))
        c1.append(cats[a]); c2.append(cats[b]); s1.append(sup[a]); s2.append(sup[b])
    return torch.stack(hs), torch.stack(ho), torch.stack(c1), torch.stack(c2), s1, s2


def _run(fn, samples, sd, pair_list, sgdet, batch):
    rel, sup, conn = [], [], []
    for k in range(0, len(pair_list), batch):
        r1, r2, r3, s, c = fn(sd, *_inputs(samples, pair_list[k:k + batch], sgdet))[:5]
        rel.append(torch.cat((r1, r2, r3), 1)); sup.append(s); conn.append(c)
    return torch.cat(rel).numpy(), torch.cat(sup).numpy(), torch.cat(conn).numpy()


def oracle_scores(samples, sd, pair_list, sgdet=False, batch=12):
    """fp32 reference formulation on `pair_list` = [(image index in samples, subject box, object box)], in calls of `batch`
    rows (config.yaml:53 batch_size 12) -> (relation [n,50] log-joints, super [n,3], connectivity [n,1])."""
    return _run(O.bayesian_relation_classifier, samples, sd, pair_list, sgdet, batch)


def _rounded_classifier(dtype):
    rd = lambda t: t.to(dtype).to(torch.float32)

    def fn(sd, h_sub, h_obj, c1, c2, s1, s2):
        with torch.no_grad():
            w = {k: rd(v) for k, v in sd.items() if k.endswith("weight") and not k.startswith(("fc3", "fc4", "fc5"))}
            a = rd(torch.tanh(F.conv2d(rd(h_sub), w["conv1_1.weight"], sd["conv1_1.bias"])))
            b = rd(torch.tanh(F.conv2d(rd(h_obj), w["conv1_2.weight"], sd["conv1_2.bias"])))
            n_half = a.shape[1]
            # conv2_1 is evaluated as a subject half and an object half (linear before the ReLU); each half is a stored operand
            u = rd(F.conv2d(a, w["conv2_1.weight"][:, :n_half], None, padding=1))
            v = rd(F.conv2d(b, w["conv2_1.weight"][:, n_half:], sd["conv2_1.bias"], padding=1))
            h = rd(F.max_pool2d(F.relu(u + v), 2, 2))
            h = rd(F.max_pool2d(F.relu(F.conv2d(h, w["conv3_1.weight"], sd["conv3_1.bias"], padding=1)), 2, 2))
            h = rd(F.relu(F.linear(h.reshape(h.shape[0], -1), w["fc1.weight"], sd["fc1.bias"])))
            w2 = sd["fc2.weight"].clone()
            w2[:, :4096] = w["fc2.weight"][:, :4096]            # the label columns are fp32 embedding rows on the device
            pred = F.relu(F.linear(O.concat_labels(h, c1, c2, s1, s2), w2, sd["fc2.bias"]))
            return O.hier_head(sd, pred) + (pred,)
    return fn


def operand_rounded_scores(samples, sd, pair_list, sgdet=False, batch=12, dtype=torch.bfloat16):
    return _run(_rounded_classifier(dtype), samples, sd, pair_list, sgdet, batch)


def parity_stats(rel_got, rel_ref, splits=SPLITS):
    """rel_*: [n, sum(splits)] log joint probabilities.  -> dict for a test assertion or the bench line."""
    g, r = np.asarray(rel_got, dtype=np.float64), np.asarray(rel_ref, dtype=np.float64)
    dp = np.abs(np.exp(g) - np.exp(r))
    offs = np.concatenate(([0], np.cumsum(splits)))
    flips = sum(int((g[:, offs[k]:offs[k + 1]].argmax(1) != r[:, offs[k]:offs[k + 1]].argmax(1)).sum()) for k in range(len(splits)))
    top = np.exp(r).max(1)
    return dict(n=int(g.shape[0]), max_abs_dp=float(dp.max()), mean_abs_dp=float(dp.max(1).mean()),
                max_rel_logp_err=float((np.abs(g - r) / np.maximum(np.abs(r), 1e-6)).max()),
                argmax_flip_rate=flips / float(len(splits) * g.shape[0]), argmax_flips=flips,
                top_joint_prob_median=float(np.median(top)), top1_flip_rate=float((g.argmax(1) != r.argmax(1)).mean()))


def sgb_tail_parity(sd, batch, pairs, num_objs, rel_got, per_image=9, seed=3):
    """Config 5 (SGB plug-and-play tail): `per_image` random directed pairs of every image through the fp32 restatement of
    roi_relation_predictors.py:399-459 (sgb_oracle.predictor_tail) against `rel_got` = the device's [P, 50] log joint
    probabilities (rows in the concatenated per-image pair order).  -> parity_stats dict."""
    from . import sgb_oracle as SO
    rng = np.random.default_rng(seed)
    base = np.concatenate(([0], np.cumsum([int(p.shape[0]) for p in pairs])))
    pick = [np.sort(rng.choice(int(p.shape[0]), size=min(per_image, int(p.shape[0])), replace=False)) for p in pairs]
    sub_pairs = [pairs[i].cpu()[torch.from_numpy(pick[i])] for i in range(len(pairs))]
    rows = np.concatenate([pick[i] + base[i] for i in range(len(pairs))])
    o1, o2, o3, _ = SO.predictor_tail(sd, batch["edge_ctx"], sub_pairs, num_objs, batch["obj_labels"], batch["union_features"][torch.from_numpy(rows)])
    rel_ref = torch.cat((torch.cat(list(o1)), torch.cat(list(o2)), torch.cat(list(o3))), dim=1).numpy()
    got = rel_got[torch.from_numpy(rows).to(rel_got.device)].cpu().numpy() if isinstance(rel_got, torch.Tensor) else np.asarray(rel_got)[rows]
    return parity_stats(got, rel_ref)
