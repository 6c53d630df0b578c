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
