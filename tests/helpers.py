"""Shared test plumbing: oracle construction from the packed tables, case replay, golden loading."""
import os

import numpy as np

from oracle import hiercom_oracle as O
from scene_graph_commonsense_b200 import synthetic, tables

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SPLITS = (15, 11, 24)


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def cs_key_arrays(run_mode, cs):
    """(aligned, violated) packed-key arrays for a case, or (None, None) when the filter is off."""
    if run_mode not in ("eval_cs", "train_cs"):
        return None, None
    if cs is None:
        return tables.commonsense_aligned_keys(), tables.commonsense_violated_keys()
    return synthetic.synthetic_cs_keys(*cs)


def oracle_evaluators(case, hierar=True):
    al, vi = cs_key_arrays(case["run_mode"], case.get("cs"))
    ev = O.OracleEvaluator(SPLITS, hierar=hierar,
                           aligned=None if al is None else set(al.tolist()),
                           violated=None if vi is None else set(vi.tolist()),
                           zero_shot=set(tables.zero_shot_keys().tolist()))
    t3 = O.OracleEvaluatorTop3(SPLITS) if hierar else None
    return ev, t3


def flat_metrics(m):
    out = []
    for x in m:
        if x is None:
            continue
        for v in x:
            out.append(np.atleast_1d(np.asarray(v, dtype=np.float64)))
    return np.concatenate(out)


def replay_predcls_case(case):
    samples = synthetic.make_batch(case["ids"], case["n"], with_maps=False, p_rel=0.5)
    ev, t3 = oracle_evaluators(case, case["hierar"])
    stats = dict(num_not_connected=0, num_connected=0, num_connected_pred=0, connectivity_precision=0,
                 connectivity_recall=0)
    m = m3 = None
    for w in case.get("windows", [list(range(len(samples)))]):
        batch = [samples[i] for i in w]
        fn = synthetic.batch_score_fn(batch, SPLITS, **case["kw"])
        O.replay_predcls(batch, fn, ev, t3, stats=stats, features=False)
        m = ev.compute(per_class=True)
        ev.clear_data()
        if t3 is not None:
            m3 = t3.compute(per_class=True)
            t3.clear_data()
    return ev, t3, m, m3, stats, samples


def train_case_inputs(case):
    """Seed-reproducible inputs of a TRAIN_CASES entry: samples, head weights, hidden vectors, per-row tables, calls.
    Row order: image-major, then t = g(g-1)/2 + e, then direction (oracle/train_oracle.training_groups)."""
    import torch
    from oracle import train_oracle as TO
    samples = synthetic.make_batch(case["ids"], case["n"], with_maps=False, p_rel=case.get("p_rel", 0.5))
    sd = synthetic.head_state_dict(seed=7, logit_gain=case.get("gain", 1.0), flat=not case["hierar"])
    sd = {k: v for k, v in sd.items() if k.startswith(("fc3", "fc4", "fc5"))}
    counts = [s.bbox.shape[0] for s in samples]
    row_img, row_g, row_e, row_dir, groups = TO.training_groups(counts)
    g = torch.Generator(device="cpu")
    g.manual_seed(4242 + case["ids"][0])
    pred = torch.relu(torch.randn(len(row_img), 512, generator=g))       # post-ReLU hidden vectors (model.py:170)
    target = np.full(len(row_img), -1, dtype=np.int64)
    cat_sub = np.zeros(len(row_img), dtype=np.int64)
    cat_obj = np.zeros(len(row_img), dtype=np.int64)
    for r in range(len(row_img)):
        s = samples[row_img[r]]
        gg, ee, d = int(row_g[r]), int(row_e[r]), int(row_dir[r])
        rel = int(s.relationships[gg - 1][ee])
        direc = int(s.subj_or_obj[gg - 1][ee])
        if (d == 0 and direc == 1) or (d == 1 and direc == 0):           # train_utils.py:62-73
            target[r] = rel
        sub, obj = (gg, ee) if d == 0 else (ee, gg)
        cat_sub[r], cat_obj[r] = int(s.categories[sub]), int(s.categories[obj])
    box_base = np.concatenate(([0], np.cumsum(counts)))
    row_sub = np.array([box_base[row_img[r]] + (row_g[r] if row_dir[r] == 0 else row_e[r]) for r in range(len(row_img))], dtype=np.int32)
    row_obj = np.array([box_base[row_img[r]] + (row_e[r] if row_dir[r] == 0 else row_g[r]) for r in range(len(row_img))], dtype=np.int32)
    box_cat = np.concatenate([s.categories.numpy() for s in samples]).astype(np.int32)
    return dict(samples=samples, sd=sd, pred=pred, counts=counts, groups=groups, target=target, cat_sub=cat_sub, cat_obj=cat_obj,
                row_sub=row_sub, row_obj=row_obj, box_cat=box_cat)
