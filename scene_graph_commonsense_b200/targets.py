"""Host-side target preparation for SGDET / SGCLS windows (reference utils.py:294-352 `match_target_sgd`).

The reference rebuilds, per image, a flat list of GT triplets in (g, e) loop order with subject/object assigned from
`subj_or_obj`.  Here the same list is emitted once per window as CSR tables the top-K/match kernel reads directly.
Reference quirk kept (it changes the recall denominator, so parity requires it): the outer loop is
`range(len(relationships[image]))` and `relationships` has N-1 rows, so relations of the LAST box (g = N-1) never
become targets.
"""
import numpy as np


def flat_targets_sgd(samples):
    offsets, label, sub, obj = [0], [], [], []
    cats, boxes = [], []
    base = 0
    for s in samples:
        n = len(s.categories)
        for g in range(1, len(s.relationships)):
            rel = s.relationships[g - 1]
            d = s.subj_or_obj[g - 1]
            for e in range(g):
                dv = float(d[e])
                if dv == 1:
                    a, b = g, e
                elif dv == 0:
                    a, b = e, g
                else:
                    continue
                label.append(int(rel[e]))
                sub.append(base + a)
                obj.append(base + b)
        offsets.append(len(label))
        cats.append(np.asarray(s.categories, dtype=np.int32))
        boxes.append(np.asarray(s.bbox, dtype=np.int32).reshape(-1, 4))
        base += n
    i32 = lambda x: np.asarray(x, dtype=np.int32)
    return dict(offsets=i32(offsets), label=i32(label) if label else np.zeros(1, np.int32)[:0], sub=i32(sub), obj=i32(obj),
                cat=np.concatenate(cats).astype(np.int32), box=np.concatenate(boxes).astype(np.int32))
