"""N4: training-side losses on the hierarchical head - CUDA kernels (through torch.ops.hiercom.*) against goldens produced by the
UNMODIFIED reference `train_utils.train_one_direction` + autograd (oracle/make_golden_train.py), and against the oracle
restatement on a cfg2-shaped window.  Tolerances (fp32, different summation order / exp implementation than torch CPU):
per-call losses 2e-5 abs+rel, step loss 2e-5 rel, gradients 1e-4 rel + 5e-5 abs (gradient entries reach O(10))."""
import numpy as np
import pytest
import torch

from scene_graph_commonsense_b200 import synthetic, tables
from tests.golden_cases import TRAIN_CASES
from tests.helpers import SPLITS, cs_key_arrays, golden, train_case_inputs

pytestmark = pytest.mark.gpu
LAMBDAS = dict(lambda_connectivity=0.1, lambda_not_connected=1, lambda_commonsense=1, lambda_cs_weak=0.1, lambda_cs_strong=10)


class Head(torch.nn.Module):
    def __init__(self, sd, hier):
        super().__init__()
        names = ("fc3_1", "fc3_2", "fc3_3", "fc4", "fc5") if hier else ("fc3", "fc4")
        for n in names:
            lin = torch.nn.Linear(512, sd[n + ".weight"].shape[0])
            lin.load_state_dict({"weight": sd[n + ".weight"], "bias": sd[n + ".bias"]})
            setattr(self, n, lin)
        self.names = names


def make_loss(case, dev):
    from scene_graph_commonsense_b200 import losses
    args = synthetic.reference_args(run_mode=case["run_mode"], hierar=case["hierar"])
    args["training"].update(LAMBDAS)
    al, vi = cs_key_arrays(case["run_mode"], case.get("cs"))
    return losses, losses.RelationLoss(args, dev, aligned_keys=al, violated_keys=vi, temps=case.get("temps", (1.0, 1.0, 1.0)))


def run_ours(case, inp, dev):
    losses, crit = make_loss(case, dev)
    rows = losses.training_rows(inp["samples"], dev)
    head = Head(inp["sd"], case["hierar"]).to(dev)
    pred = inp["pred"].to(dev).requires_grad_(True)
    out = crit(pred, head, rows)
    out["losses"].backward()
    gw = torch.cat([getattr(head, n).weight.grad for n in head.names]).cpu().numpy()
    gb = torch.cat([getattr(head, n).bias.grad for n in head.names]).cpu().numpy()
    return out, pred.grad.cpu().numpy(), gw, gb


@pytest.mark.parametrize("name", sorted(TRAIN_CASES))
def test_losses_and_gradients_match_reference(name):
    case = TRAIN_CASES[name]
    g = golden(name)
    inp = train_case_inputs(case)
    out, gp, gw, gb = run_ours(case, inp, torch.device("cuda:0"))
    np.testing.assert_allclose(out["per_call"].cpu().numpy(), g["per_call"], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(float(out["losses"].detach()), float(g["total"]), rtol=2e-5)
    pc = g["per_call"].astype(np.float64)
    w = np.arange(len(pc), 0, -1, dtype=np.float64)
    np.testing.assert_allclose([float(out["loss_relationship"]), float(out["loss_connectivity"]), float(out["loss_commonsense"])],
                               (pc * w[:, None]).sum(0), rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose(gp, g["grad_pred"], rtol=1e-4, atol=5e-5)
    np.testing.assert_allclose(gw, g["grad_w"], rtol=1e-4, atol=5e-5)
    np.testing.assert_allclose(gb, g["grad_b"], rtol=1e-4, atol=5e-5)


def test_bit_reproducible_and_upstream_scale():
    case = TRAIN_CASES["tr_hier_cs"]
    inp = train_case_inputs(case)
    dev = torch.device("cuda:0")
    a = run_ours(case, inp, dev)
    b = run_ours(case, inp, dev)
    assert torch.equal(a[0]["per_call"], b[0]["per_call"]) and float(a[0]["losses"].detach()) == float(b[0]["losses"].detach())
    for x, y in zip(a[1:], b[1:]):
        assert np.array_equal(x, y)                              # no float atomics anywhere
    losses, crit = make_loss(case, dev)
    rows = losses.training_rows(inp["samples"], dev)
    head = Head(inp["sd"], True).to(dev)
    pred = inp["pred"].to(dev).requires_grad_(True)
    (crit(pred, head, rows)["losses"] * 0.25).backward()
    np.testing.assert_allclose(pred.grad.cpu().numpy(), a[1] * 0.25, rtol=1e-6, atol=1e-7)


def test_no_grad_path_and_shape_errors():
    case = TRAIN_CASES["tr_flat_plain"]
    inp = train_case_inputs(case)
    dev = torch.device("cuda:0")
    losses, crit = make_loss(case, dev)
    rows = losses.training_rows(inp["samples"], dev)
    head = Head(inp["sd"], False).to(dev)
    with torch.no_grad():
        out = crit(inp["pred"].to(dev), head, rows)
    np.testing.assert_allclose(float(out["losses"].detach()), float(golden("tr_flat_plain")["total"]), rtol=2e-5)
    with pytest.raises(RuntimeError):
        crit(inp["pred"][:-1].to(dev), head, rows)
    with pytest.raises((RuntimeError, NotImplementedError)):
        crit(inp["pred"], head.cpu(), rows)                      # CPU tensors reach no kernel


def test_cfg2_shaped_window_against_oracle():
    """64 images x 40 boxes (99 840 rows, 1 560 calls of 64 rows), train_cs with the shipped commonsense sets: per-call losses and the step
    loss against the loop oracle; gradients against torch autograd through the oracle formulation."""
    from oracle import train_oracle as TO
    from scene_graph_commonsense_b200 import losses
    dev = torch.device("cuda:0")
    case = dict(ids=list(range(300, 364)), n=[40] * 64, run_mode="train_cs", hierar=True, gain=3.0, p_rel=0.3)
    samples = synthetic.make_batch(case["ids"], case["n"], with_maps=False, p_rel=0.3)
    sd = {k: v for k, v in synthetic.head_state_dict(seed=7, logit_gain=3.0).items() if k.startswith(("fc3", "fc4", "fc5"))}
    args = synthetic.reference_args(run_mode="train_cs", hierar=True)
    args["training"].update(LAMBDAS)
    crit = losses.RelationLoss(args, dev)
    rows = losses.training_rows(samples, dev)
    assert rows.n_rows == 99840 and rows.n_groups == 1560
    g = torch.Generator(device="cpu"); g.manual_seed(99)
    pred0 = torch.relu(torch.randn(rows.n_rows, 512, generator=g))
    head = Head(sd, True).to(dev)
    pred = pred0.to(dev).requires_grad_(True)
    out = crit(pred, head, rows)
    out["losses"].backward()
    # oracle on the same rows
    sd64 = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    p64 = pred0.clone().requires_grad_(True)
    rel, sup, conn = TO.head_outputs(p64, sd64, SPLITS, (1.0, 1.0, 1.0), True)
    unpack = lambda k: (int(k) // 7500, (int(k) // 150) % 50, int(k) % 150)
    aligned = {unpack(k) for k in tables.commonsense_aligned_keys()}
    violated = {unpack(k) for k in tables.commonsense_violated_keys()}
    cnt = tables.vg_predicate_counts().astype(np.float64)
    cw = torch.from_numpy((1 - cnt / cnt.sum()).astype(np.float32))
    box_cat = rows.box_cat.cpu().long()
    cs, co = box_cat[rows.row_sub.cpu().long()], box_cat[rows.row_obj.cpu().long()]
    go, gr = rows.group_offsets.cpu().numpy(), rows.group_rows.cpu().long()
    groups = [gr[go[m]:go[m + 1]] for m in range(rows.n_groups)]
    lam = dict(connectivity=0.1, not_connected=1.0, commonsense=1.0, cs_weak=0.1, cs_strong=10.0)
    per_call, total = TO.step_losses(rel, sup, conn, rows.row_target.cpu().long(), cs, co, groups, cw, lam, aligned, violated, SPLITS, True)
    np.testing.assert_allclose(out["per_call"].cpu().numpy(), per_call.numpy(), rtol=3e-5, atol=3e-5)
    np.testing.assert_allclose(float(out["losses"].detach()), float(total.detach()), rtol=3e-5)
    total.backward()
    scale = float(np.abs(p64.grad.numpy()).max())
    np.testing.assert_allclose(pred.grad.cpu().numpy(), p64.grad.numpy(), rtol=2e-4, atol=2e-5 * scale)
    gw = torch.cat([getattr(head, n).weight.grad for n in head.names]).cpu().numpy()
    gw_o = np.concatenate([sd64[n + ".weight"].grad.numpy() for n in head.names])
    np.testing.assert_allclose(gw, gw_o, rtol=1e-3, atol=1e-4 * float(np.abs(gw_o).max()))
