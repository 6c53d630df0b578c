"""Golden vectors for the training-side losses from the UNMODIFIED reference.

Run in the build container only:   python oracle/make_golden_train.py        -> tests/golden/tr_*.npz

What runs reference code: `train_utils.train_one_direction` (commonsense penalty :36-60, connectivity BCE :62-90,
`calculate_losses_on_relationships` :116-157 with `utils.super_relation_processing`), the criteria exactly as train_test.py:105-117
builds them (`utils.get_num_each_class_reordered`), the head `model.BayesianHead` (+ `fc4`, model.py:171) or the flat head's
`fc3`/`fc4` (model.py:100-101), and torch autograd for the gradients.  The driver loop below follows train_test.py:187-258
including its running-sum accumulation (`losses += loss_relationship + ...` with cumulative operands).  The recall evaluators are
not touched (batch_count is chosen so the `eval_freq` branch is not taken); the contrastive loss is outside N4.
"""
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("HIERCOM_REFERENCE", "/root/reference")
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
sys.modules.setdefault("torchmetrics", types.ModuleType("torchmetrics"))
os.chdir(REF)

import model as ref_model              # noqa: E402
import train_utils as ref_train_utils  # noqa: E402
import utils as ref_utils              # noqa: E402

from scene_graph_commonsense_b200 import synthetic  # noqa: E402
from tests.golden_cases import TRAIN_CASES  # noqa: E402
from tests.helpers import train_case_inputs  # noqa: E402

LAMBDAS = dict(lambda_connectivity=0.1, lambda_not_connected=1, lambda_commonsense=1, lambda_cs_weak=0.1, lambda_cs_strong=10,
               eval_freq=100)     # config.yaml:63-71


class TrainHead(torch.nn.Module):
    """The part of the reference classifier after the hidden vector, built from reference modules."""

    def __init__(self, sd, hierar, temps):
        super().__init__()
        self.hierar = hierar
        if hierar:
            self.head = ref_model.BayesianHead(512, 15, 11, 24, T1=temps[0], T2=temps[1], T3=temps[2])
            self.head.load_state_dict({k: v for k, v in sd.items() if k.split(".")[0] in ("fc3_1", "fc3_2", "fc3_3", "fc5")})
        else:
            self.fc3 = torch.nn.Linear(512, 50)
            self.fc3.load_state_dict({"weight": sd["fc3.weight"], "bias": sd["fc3.bias"]})
        self.fc4 = torch.nn.Linear(512, 1)
        self.fc4.load_state_dict({"weight": sd["fc4.weight"], "bias": sd["fc4.bias"]})
        self.pred = None
        self.rows = None

    def __call__(self, h_sub, h_obj, c1, c2, s1, s2, rank, h_sub_aug=None, h_obj_aug=None):
        pred = self.pred[self.rows]
        conn = self.fc4(pred)                                           # model.py:171 / :101
        if not self.hierar:
            return self.fc3(pred), conn, pred, pred                     # model.py:100-102
        r1, r2, r3, sup = self.head(pred)                               # model.py:24-34 == :172-184
        return r1, r2, r3, sup, conn, pred, pred


def run_case(name, c):
    args = synthetic.reference_args(run_mode=c["run_mode"], hierar=c["hierar"])
    args["training"].update(LAMBDAS)
    inp = train_case_inputs(c)
    samples, sd, pred0 = inp["samples"], inp["sd"], inp["pred"]
    head = TrainHead(sd, c["hierar"], c.get("temps", (1.0, 1.0, 1.0)))
    head.pred = pred0.clone().requires_grad_(True)
    aligned = violated = None
    if c["run_mode"] == "train_cs":
        al, vi = synthetic.synthetic_cs_keys(*c["cs"])
        unpack = lambda k: (int(k) // 7500, (int(k) // 150) % 50, int(k) % 150)
        aligned = {unpack(k): 1 for k in al}
        violated = {unpack(k): 1 for k in vi}
    # criteria exactly as train_test.py:105-117
    relation_count = ref_utils.get_num_each_class_reordered(args)
    class_weight = 1 - relation_count / torch.sum(relation_count)
    if c["hierar"]:
        crit = [torch.nn.NLLLoss(weight=class_weight[:15]), torch.nn.NLLLoss(weight=class_weight[15:26]),
                torch.nn.NLLLoss(weight=class_weight[26:]), torch.nn.NLLLoss()]
    else:
        crit = torch.nn.CrossEntropyLoss(weight=class_weight)
    crit_conn = torch.nn.BCEWithLogitsLoss()

    counts = [s.bbox.shape[0] for s in samples]
    base = np.concatenate(([0], np.cumsum([n * (n - 1) for n in counts])))
    relationships = [s.relationships for s in samples]
    subj_or_obj = [s.subj_or_obj for s in samples]
    relations_target, direction_target = [], []
    num_graph_iter = torch.as_tensor(counts) - 1                        # train_test.py:173-180
    for graph_iter in range(int(max(num_graph_iter))):
        keep = torch.nonzero(num_graph_iter > graph_iter).view(-1)
        relations_target.append(torch.vstack([relationships[i][graph_iter] for i in keep]).T)
        direction_target.append(torch.vstack([subj_or_obj[i][graph_iter] for i in keep]).T)

    batch_size = len(samples)
    hca = [[] for _ in range(batch_size)]
    hcl = [[] for _ in range(batch_size)]
    losses, loss_connectivity, loss_relationship, loss_commonsense = 0.0, 0.0, 0.0, 0.0   # train_test.py:187
    per_call = []
    num_graph_iter = torch.as_tensor(counts)
    for graph_iter in range(int(max(num_graph_iter))):
        keep = torch.nonzero(num_graph_iter > graph_iter).view(-1)
        cat_g = torch.tensor([samples[i].categories[graph_iter] for i in keep])
        bb_g = torch.stack([samples[i].bbox[graph_iter] for i in keep])
        for edge_iter in range(graph_iter):
            cat_e = torch.tensor([samples[i].categories[edge_iter] for i in keep])
            bb_e = torch.stack([samples[i].bbox[edge_iter] for i in keep])
            iou_mask = torch.ones(len(keep), dtype=torch.bool)          # train_test.py:208
            t = graph_iter * (graph_iter - 1) // 2 + edge_iter
            for first in (True, False):
                head.rows = torch.tensor([base[int(i)] + 2 * t + (0 if first else 1) for i in keep])
                a = (cat_g, cat_e, None, None, bb_g, bb_e) if first else (cat_e, cat_g, None, None, bb_e, bb_g)
                r = ref_train_utils.train_one_direction(head, args, None, None, *a, None, None, iou_mask, 'cpu', graph_iter, edge_iter,
                                                        keep, None, None, crit, crit_conn, relations_target, direction_target, 1,
                                                        hca, hcl, aligned, violated, 10, first_direction=first)
                cur_rel, cur_conn, cur_cs = r[0], r[1], r[2]
                per_call.append([float(torch.as_tensor(x).detach()) for x in (cur_rel, cur_conn, cur_cs)])
                loss_relationship += cur_rel                             # train_test.py:219-230 / :244-255
                loss_connectivity += cur_conn
                loss_commonsense += cur_cs
                losses += loss_relationship + args['training']['lambda_connectivity'] * loss_connectivity \
                    + args['training']['lambda_commonsense'] * loss_commonsense
    losses.backward()
    out = dict(per_call=np.array(per_call, dtype=np.float32), total=np.float32(float(losses)),
               grad_pred=head.pred.grad.numpy().astype(np.float32))
    names = ["fc3_1", "fc3_2", "fc3_3", "fc4", "fc5"] if c["hierar"] else ["fc3", "fc4"]
    mods = {"fc4": head.fc4}
    if c["hierar"]:
        mods.update({n: getattr(head.head, n) for n in ("fc3_1", "fc3_2", "fc3_3", "fc5")})
    else:
        mods["fc3"] = head.fc3
    out["grad_w"] = np.concatenate([mods[n].weight.grad.numpy() for n in names]).astype(np.float32)
    out["grad_b"] = np.concatenate([mods[n].bias.grad.numpy() for n in names]).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    pc = out["per_call"]
    print(name, "calls", len(pc), "total %.5f" % out["total"], "sum rel/conn/cs", pc.sum(0), "|grad_pred| %.4f" % np.abs(out["grad_pred"]).max())


if __name__ == "__main__":
    for name, c in TRAIN_CASES.items():
        run_case(name, c)
