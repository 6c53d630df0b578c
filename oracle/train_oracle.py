"""TEST INFRASTRUCTURE ONLY - CPU restatement of the training-side losses on the hierarchical head.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file; the product
(scene_graph_commonsense_b200/) never does.  Pinned by oracle/make_golden_train.py, which drives the UNMODIFIED reference
`train_utils.train_one_direction` / `calculate_losses_on_relationships` through the loop body of train_test.py:187-258 with the
reference `model.BayesianHead` as the head and lets torch autograd produce the gradients (tests/golden/train_*.npz,
replayed by tests/test_train_oracle_golden.py).

Formulation.  The reference evaluates the losses once per CALL = (graph_iter g, edge_iter e, direction) over the images of the
lock-step batch that own box g (train_test.py:189-258); here every directed pair is a ROW and a call is a GROUP of rows:

  connected rows      directed target != -1                                        train_utils.py:62-73
  loss_connectivity   connected rows exist: BCEWithLogits(conn[connected], 1)      train_utils.py:88-90 (OVERWRITES the
                      else lambda_not_connected * BCEWithLogits(conn[not connected], 0), NaN (no row) -> 0      :67-68 value)
  loss_relationship   hierarchical: NLL(super[connected], super target) + sum_k weighted-NLL(rel_k[connected_k], t - off_k)
                      flat: weighted CrossEntropy(relation[connected], t)          train_utils.py:116-157
  loss_commonsense    run_mode == 'train_cs': p = max softmax(rel_k) per (row, k) [hier: 3 per row, flat: 1], triplet
                      (cat_sub, argmax + off_k, cat_obj); lambda_cs_weak * mean(p[not in aligned]) +
                      lambda_cs_strong * mean(p[in violated])                      train_utils.py:36-60
  step loss           train_test.py:219-230 adds the RUNNING sums after every call:
                      losses += loss_relationship + l_conn * loss_connectivity + l_cs * loss_commonsense (all cumulative), i.e.
                      call m of M (0-based) enters the step loss with weight (M - m).
"""
import torch
import torch.nn.functional as F


def head_outputs(pred, sd, splits=(15, 11, 24), temps=(1.0, 1.0, 1.0), hier=True):
    """model.py:170-184 (hierarchical) / model.py:97-101 (flat) from the 512-d hidden vector."""
    conn = F.linear(pred, sd["fc4.weight"], sd["fc4.bias"])[:, 0]
    if not hier:
        return F.linear(pred, sd["fc3.weight"], sd["fc3.bias"]), None, conn
    sup = F.log_softmax(F.linear(pred, sd["fc5.weight"], sd["fc5.bias"]), dim=1)
    rel = []
    for k in range(3):
        z = F.linear(pred, sd["fc3_%d.weight" % (k + 1)], sd["fc3_%d.bias" % (k + 1)])
        rel.append(F.log_softmax(z / temps[k], dim=1) + sup[:, k].view(-1, 1))
    return torch.cat(rel, dim=1), sup, conn


def call_losses(relation, sup, conn, target, cat_sub, cat_obj, class_weight, lam, aligned=None, violated=None,
                splits=(15, 11, 24), hier=True):
    """Losses of ONE call over its rows (train_utils.py:36-100,116-157).  `target` = directed labels (-1: not connected).
    aligned / violated: sets of (s, p, o) tuples or None (run_mode != 'train_cs').  Returns (rel, conn, cs) tensors/floats."""
    G, Pn = splits[0], splits[1]
    offs = (0, G, G + Pn)
    ends = (G, G + Pn, relation.shape[1])
    n = relation.shape[0]
    zero = relation.new_zeros(())
    loss_cs = zero
    if aligned is not None:
        if hier:
            probs = torch.hstack([torch.max(F.softmax(relation[:, offs[k]:ends[k]], dim=1), dim=1)[0] for k in range(3)])
            pred = torch.hstack([torch.argmax(relation[:, offs[k]:ends[k]], dim=1) + offs[k] for k in range(3)])
            cs, co = cat_sub.repeat(3), cat_obj.repeat(3)
        else:
            probs = torch.max(F.softmax(relation, dim=1), dim=1)[0]
            pred = torch.argmax(relation, dim=1)
            cs, co = cat_sub, cat_obj
        trip = [(int(cs[i]), int(pred[i]), int(co[i])) for i in range(len(pred))]
        not_yes = torch.tensor([t not in aligned for t in trip], dtype=torch.bool)
        in_no = torch.tensor([t in violated for t in trip], dtype=torch.bool)
        if int(not_yes.sum()) > 0:
            loss_cs = loss_cs + lam["cs_weak"] * probs[not_yes].mean()
        if int(in_no.sum()) > 0:
            loss_cs = loss_cs + lam["cs_strong"] * probs[in_no].mean()
    connected = torch.nonzero(target != -1).flatten()
    not_connected = torch.nonzero(target == -1).flatten()
    loss_conn = zero
    if len(not_connected) > 0:
        loss_conn = lam["not_connected"] * F.binary_cross_entropy_with_logits(conn[not_connected], torch.zeros(len(not_connected)))
    loss_rel = zero
    if len(connected) > 0:
        loss_conn = F.binary_cross_entropy_with_logits(conn[connected], torch.ones(len(connected)))
        t = target[connected].long()
        if hier:
            st = (t >= G).long() + (t >= G + Pn).long()
            loss_rel = F.nll_loss(sup[connected], st)
            for k in range(3):
                sel = torch.nonzero(st == k).flatten()
                if len(sel) > 0:
                    loss_rel = loss_rel + F.nll_loss(relation[connected][sel][:, offs[k]:ends[k]], t[sel] - offs[k],
                                                     weight=class_weight[offs[k]:ends[k]])
        else:
            loss_rel = F.cross_entropy(relation[connected], t, weight=class_weight)
    return loss_rel, loss_conn, loss_cs


def step_losses(relation, sup, conn, target, cat_sub, cat_obj, groups, class_weight, lam, aligned=None, violated=None,
                splits=(15, 11, 24), hier=True, group_weight=None):
    """All calls of one training step.  `groups`: list of int64 row-index tensors in call order (empty groups are legal and
    contribute nothing but still count in M).  Returns (per_call [M,3] detached, step loss with the reference's running-sum
    weights (train_test.py:219-230): sum_m (M - m) (rel_m + l_conn conn_m + l_cs cs_m))."""
    M = len(groups)
    per_call = torch.zeros(M, 3)
    total = relation.new_zeros(())
    for m, rows in enumerate(groups):
        if len(rows) == 0:
            continue
        lr, lc, ls = call_losses(relation[rows], None if sup is None else sup[rows], conn[rows], target[rows], cat_sub[rows],
                                 cat_obj[rows], class_weight, lam, aligned, violated, splits, hier)
        per_call[m] = torch.stack([lr.detach(), lc.detach(), ls.detach()])
        w = float(M - m) if group_weight is None else float(group_weight[m])
        total = total + w * (lr + lam["connectivity"] * lc + lam["commonsense"] * ls)
    return per_call, total


def training_groups(counts):
    """Rows and calls of one lock-step training batch (train_test.py:189-258: no overlap skip in training).
    Row order: image-major, then t = g(g-1)/2 + e, then direction (0: sub=g obj=e, 1: swapped).  Call m = 2 t + dir holds the
    rows of every image with N_i > g, in image order.  Returns (row_img, row_g, row_e, row_dir, groups)."""
    import numpy as np
    counts = [int(c) for c in counts]
    row_img, row_g, row_e, row_dir = [], [], [], []
    base = [0]
    for i, n in enumerate(counts):
        for g in range(n):
            for e in range(g):
                for d in (0, 1):
                    row_img.append(i); row_g.append(g); row_e.append(e); row_dir.append(d)
        base.append(len(row_img))
    nmax = max(counts) if counts else 0
    groups = []
    for g in range(nmax):
        for e in range(g):
            t = g * (g - 1) // 2 + e
            for d in (0, 1):
                groups.append(torch.tensor([base[i] + 2 * t + d for i, n in enumerate(counts) if n > g], dtype=torch.int64))
    return (np.array(row_img), np.array(row_g), np.array(row_e), np.array(row_dir), groups)
