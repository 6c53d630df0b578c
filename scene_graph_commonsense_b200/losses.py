"""Training-side losses on the fused hierarchical head - drop-in for the loss part of the reference's
`train_utils.train_one_direction` (train_utils.py:23-113) and `calculate_losses_on_relationships` (:116-157), evaluated for ALL calls
of a training step (train_test.py:187-258) in three kernel launches, with a hand-written backward for the heads.

The reference computes, per call (graph_iter g, edge_iter e, direction) over the images of the lock-step batch that own box g:
the commonsense penalty (run_mode 'train_cs'), the connectivity BCE and the hierarchical (or flat) NLL, and accumulates them with
running sums (`losses += loss_relationship + ...` with cumulative operands, train_test.py:219-230), so call m of M enters the step
loss with weight M - m.  Here every directed pair is a row, a call is a group of rows (`training_rows`), and

    loss = RelationLoss(args, device)(pred, head, rows)["losses"];  loss.backward()

returns the same step loss and leaves the same gradients on `pred` (the 512-d hidden vectors, model.py:170) and on the head
parameters fc3_x / fc4 / fc5 as the reference's autograd does (tests/test_gpu_train.py, goldens from the unmodified reference).
There is no CPU implementation: the operators are CUDA-only (ops.py).  The supervised-contrastive term (train_test.py:261-272) is
outside this row.
"""
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
from torch import nn

from . import ops, tables


@dataclass
class TrainRows:
    """All directed pairs of one training window (no overlap skip in training, train_test.py:208) and their calls."""
    n_rows: int
    n_groups: int
    row_sub: torch.Tensor         # int32 [P] global box id of the subject
    row_obj: torch.Tensor         # int32 [P]
    row_target: torch.Tensor      # int32 [P] directed predicate label, -1 = not connected (train_utils.py:62-73)
    group_offsets: torch.Tensor   # int32 [M+1]
    group_rows: torch.Tensor      # int32 [P] rows of each call, image order
    group_weight: torch.Tensor    # f32 [M]  M_b - m within each lock-step batch (train_test.py:219-230)
    box_cat: torch.Tensor         # int32 [nbox]


def tri_decode(t):
    """t = g(g-1)/2 + e  ->  (g, e), vectorised."""
    g = ((1.0 + np.sqrt(1.0 + 8.0 * t.astype(np.float64))) / 2.0).astype(np.int64)
    g = np.where(g * (g - 1) // 2 > t, g - 1, g)
    g = np.where((g + 1) * g // 2 <= t, g + 1, g)
    return g, t - g * (g - 1) // 2


def training_rows_host(counts, rel_tri, dir_tri, group_size=None):
    """Host packing (vectorised numpy) of one window: `counts` boxes per image, `rel_tri` / `dir_tri` the packed
    relationships[g-1][e] / subj_or_obj[g-1][e] at t = g(g-1)/2 + e (the HostBatch layout).  Row order: image-major, then t, then
    direction (0: sub = g, obj = e; 1: swapped).  `group_size` images form one lock-step batch (default: the whole window)."""
    counts = np.asarray(counts, dtype=np.int64)
    n_img = len(counts)
    tri = counts * (counts - 1) // 2
    tri_off = np.concatenate(([0], np.cumsum(tri)))
    box_off = np.concatenate(([0], np.cumsum(counts)))
    T = int(tri_off[-1])
    img_of_t = np.repeat(np.arange(n_img), tri)
    t_local = np.arange(T) - tri_off[img_of_t]
    g, e = tri_decode(t_local)
    gb, eb = box_off[img_of_t] + g, box_off[img_of_t] + e
    rel = np.asarray(rel_tri, dtype=np.int64)
    dr = np.asarray(dir_tri, dtype=np.int64)
    row_sub = np.stack((gb, eb), 1).reshape(-1)
    row_obj = np.stack((eb, gb), 1).reshape(-1)
    row_target = np.stack((np.where(dr == 1, rel, -1), np.where(dr == 0, rel, -1)), 1).reshape(-1)
    gs = group_size or max(n_img, 1)
    batch_of_img = np.arange(n_img) // gs
    n_batches = int(batch_of_img.max()) + 1 if n_img else 0
    nmax = np.zeros(max(n_batches, 1), dtype=np.int64)
    np.maximum.at(nmax, batch_of_img, counts)
    calls = nmax * (nmax - 1)                                          # M_b = 2 * Nmax (Nmax - 1) / 2
    call_off = np.concatenate(([0], np.cumsum(calls)))
    b_of_row = np.repeat(batch_of_img[img_of_t], 2)
    call = call_off[b_of_row] + np.stack((2 * t_local, 2 * t_local + 1), 1).reshape(-1)
    order = np.argsort(call, kind="stable")                           # rows are image-major already: image order inside a call
    M = int(call_off[-1])
    group_offsets = np.concatenate(([0], np.cumsum(np.bincount(call, minlength=M))))
    m_local = np.arange(M) - np.repeat(call_off[:-1], calls)
    group_weight = (np.repeat(calls, calls) - m_local).astype(np.float32)
    return dict(row_sub=row_sub.astype(np.int32), row_obj=row_obj.astype(np.int32), row_target=row_target.astype(np.int32),
                group_offsets=group_offsets.astype(np.int32), group_rows=order.astype(np.int32), group_weight=group_weight)


def training_rows(samples, device, group_size=None):
    """`samples`: the dataloader's per-image records (bbox, categories, relationships, subj_or_obj; dataloader.py:159-165)."""
    counts = [int(s.bbox.shape[0]) for s in samples]
    cat = lambda lst, dt: (np.concatenate([np.concatenate([np.asarray(r_) for r_ in x]) if len(x) else np.zeros(0, dt) for x in lst])
                           if len(lst) else np.zeros(0, dt))
    h = training_rows_host(counts, cat([s.relationships for s in samples], np.int64), cat([s.subj_or_obj for s in samples], np.int64),
                           group_size)
    box_cat = np.concatenate([np.asarray(s.categories) for s in samples]).astype(np.int32)
    d = {k: torch.from_numpy(v).to(device) for k, v in h.items()}
    return TrainRows(len(h["row_sub"]), len(h["group_weight"]), d["row_sub"], d["row_obj"], d["row_target"], d["group_offsets"],
                     d["group_rows"], d["group_weight"], torch.from_numpy(box_cat).to(device))


class _HeadLossFn(torch.autograd.Function):
    """pred [P,512], w_heads [n_out,512], b_heads [n_out] -> step loss; backward = hc_hier_head_bwd on the stored d_logits."""

    @staticmethod
    def forward(ctx, pred, w_heads, b_heads, rows, cfg, out):
        pred_c = pred.detach().float().contiguous()
        w_c = w_heads.detach().float().contiguous()
        b_c = b_heads.detach().float().contiguous()
        relation, sup, conn, _, _ = ops.hier_head(pred_c, None, None, None, None, None, None, w_c, b_c, cfg["splits"],
                                                  flat=not cfg["hier"], temps=cfg["temps"])
        need_grad = any(ctx.needs_input_grad[:3])
        gl, total, d_logits = ops.hier_loss(relation, sup, conn, rows.row_target, rows.group_offsets, rows.group_rows, rows.group_weight,
                                            cfg["class_weight"], cfg["splits"], cfg["hier"], cfg["temps"], cfg["aligned"], cfg["violated"],
                                            rows.row_sub, rows.row_obj, rows.box_cat, cfg["lambdas"], want_grad=need_grad)
        out.update(group_loss=gl, total=total, relation=relation, super_relation=sup, connectivity=conn)
        if need_grad:
            ctx.save_for_backward(d_logits, pred_c, w_c)
        return total[0].clone()

    @staticmethod
    def backward(ctx, grad):
        d_logits, pred_c, w_c = ctx.saved_tensors
        want_pred = ctx.needs_input_grad[0]
        want_w = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        d_pred, d_w, d_b = ops.hier_head_bwd(d_logits, pred_c, w_c, scale=grad.detach().float().reshape(1).contiguous(),
                                             want_pred=want_pred, want_weights=want_w)
        return d_pred, (d_w if ctx.needs_input_grad[1] else None), (d_b if ctx.needs_input_grad[2] else None), None, None, None


class RelationLoss(nn.Module):
    """The criteria of train_test.py:105-117 + the loss part of train_one_direction for a whole step.

    args: the reference's config dict (`models.hierarchical_pred`, `models.num_geometric/possessive/semantic`, `training.run_mode`,
    `training.lambda_*`).  class_weight defaults to `1 - count / sum(count)` of utils.get_num_each_class_reordered (train_test.py:106).
    aligned_keys / violated_keys: packed triplet keys (tables.pack_key); default = the shipped commonsense sets when
    run_mode == 'train_cs' (train_test.py:128-133)."""

    def __init__(self, args, device, class_weight=None, aligned_keys=None, violated_keys=None, temps=(1.0, 1.0, 1.0)):
        super().__init__()
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("hiercom_b200: the training losses run on CUDA only")
        m, t = args["models"], args["training"]
        self.hier = bool(m["hierarchical_pred"])
        self.splits = (m["num_geometric"], m["num_possessive"], m["num_semantic"]) if self.hier else (m["num_relations"], 0, 0)
        if class_weight is None:
            cnt = torch.from_numpy(tables.vg_predicate_counts().astype(np.int64))
            class_weight = 1 - cnt / torch.sum(cnt)                     # train_test.py:105-106, same float ops
        self.register_buffer("class_weight", torch.as_tensor(class_weight, dtype=torch.float32).to(dev).contiguous())
        self.lambdas = (float(t.get("lambda_connectivity", 0.1)), float(t.get("lambda_not_connected", 1.0)),
                        float(t.get("lambda_commonsense", 1.0)), float(t.get("lambda_cs_weak", 0.1)), float(t.get("lambda_cs_strong", 10.0)))
        self.temps = tuple(float(x) for x in temps)
        self.aligned = self.violated = None
        if t["run_mode"] == "train_cs":
            al = tables.commonsense_aligned_keys() if aligned_keys is None else aligned_keys
            vi = tables.commonsense_violated_keys() if violated_keys is None else violated_keys
            none = np.zeros(0, dtype=np.int64)
            self.aligned = torch.from_numpy(ops.cs_bitmap_build(np.asarray(al, dtype=np.int64), none)).to(dev)
            self.violated = torch.from_numpy(ops.cs_bitmap_build(np.asarray(vi, dtype=np.int64), none)).to(dev)

    def head_parameters(self, head):
        """(w_heads, b_heads) in the kernels' row order from a module that owns fc3_x / fc4 / fc5 (or fc3 / fc4)."""
        names = ("fc3_1", "fc3_2", "fc3_3", "fc4", "fc5") if self.hier else ("fc3", "fc4")
        return (torch.cat([getattr(head, n).weight for n in names]), torch.cat([getattr(head, n).bias for n in names]))

    def forward(self, pred, head, rows: TrainRows, w_heads: Optional[torch.Tensor] = None, b_heads: Optional[torch.Tensor] = None):
        """pred: f32 [rows.n_rows, 512] hidden vectors (after fc2 + ReLU + dropout, model.py:170).  head: module with the head layers
        (or pass w_heads / b_heads).  Returns dict(losses = step loss (differentiable), loss_relationship / loss_connectivity /
        loss_commonsense = the weighted running sums the reference logs, per_call [M,3], relation, super_relation, connectivity)."""
        if w_heads is None:
            w_heads, b_heads = self.head_parameters(head)
        if pred.shape[0] != rows.n_rows:
            raise RuntimeError("hiercom_b200: pred has %d rows, the window has %d directed pairs" % (pred.shape[0], rows.n_rows))
        cfg = dict(splits=self.splits, hier=self.hier, temps=self.temps, class_weight=self.class_weight, aligned=self.aligned,
                   violated=self.violated, lambdas=self.lambdas)
        out = {}
        loss = _HeadLossFn.apply(pred, w_heads, b_heads, rows, cfg, out)
        total = out["total"]
        return dict(losses=loss, loss_relationship=total[1], loss_connectivity=total[2], loss_commonsense=total[3],
                    per_call=out["group_loss"], relation=out["relation"], super_relation=out["super_relation"],
                    connectivity=out["connectivity"])
