"""The training-loss restatement (oracle/train_oracle.py) replayed against goldens produced by the UNMODIFIED reference
`train_utils.train_one_direction` + torch autograd (oracle/make_golden_train.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import train_oracle as TO
from scene_graph_commonsense_b200 import synthetic, tables
from tests.golden_cases import TRAIN_CASES
from tests.helpers import SPLITS, cs_key_arrays, golden, train_case_inputs

LAM = dict(connectivity=0.1, not_connected=1.0, commonsense=1.0, cs_weak=0.1, cs_strong=10.0)      # config.yaml:63-69


def oracle_step(case, inp, requires_grad=True):
    sd = {k: v.clone().requires_grad_(requires_grad) for k, v in inp["sd"].items()}
    pred = inp["pred"].clone().requires_grad_(requires_grad)
    rel, sup, conn = TO.head_outputs(pred, sd, SPLITS, case.get("temps", (1.0, 1.0, 1.0)), case["hierar"])
    al, vi = cs_key_arrays(case["run_mode"], case.get("cs"))
    unpack = lambda k: (int(k) // 7500, (int(k) // 150) % 50, int(k) % 150)
    aligned = None if al is None else {unpack(k) for k in al}
    violated = None if vi is None else {unpack(k) for k in vi}
    cnt = tables.vg_predicate_counts().astype(np.float64)
    cw = torch.from_numpy((1 - cnt / cnt.sum()).astype(np.float32))
    per_call, total = TO.step_losses(rel, sup, conn, torch.from_numpy(inp["target"]), torch.from_numpy(inp["cat_sub"]),
                                     torch.from_numpy(inp["cat_obj"]), inp["groups"], cw, LAM, aligned, violated, SPLITS,
                                     case["hierar"])
    return per_call, total, pred, sd


@pytest.mark.parametrize("name", sorted(TRAIN_CASES))
def test_train_oracle_matches_reference(name):
    case = TRAIN_CASES[name]
    g = golden(name)
    inp = train_case_inputs(case)
    per_call, total, pred, sd = oracle_step(case, inp)
    assert per_call.shape == g["per_call"].shape
    np.testing.assert_allclose(per_call.numpy(), g["per_call"], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(float(total.detach()), float(g["total"]), rtol=5e-6)
    total.backward()
    np.testing.assert_allclose(pred.grad.numpy(), g["grad_pred"], rtol=1e-4, atol=2e-5)
    names = ["fc3_1", "fc3_2", "fc3_3", "fc4", "fc5"] if case["hierar"] else ["fc3", "fc4"]
    gw = np.concatenate([sd[n + ".weight"].grad.numpy() for n in names])
    gb = np.concatenate([sd[n + ".bias"].grad.numpy() for n in names])
    np.testing.assert_allclose(gw, g["grad_w"], rtol=1e-4, atol=5e-5)
    np.testing.assert_allclose(gb, g["grad_b"], rtol=1e-4, atol=5e-5)


def test_training_groups_layout():
    row_img, row_g, row_e, row_dir, groups = TO.training_groups([3, 1, 4])
    assert len(row_img) == 3 * 2 + 0 + 4 * 3
    assert len(groups) == 4 * 3                       # M = Nmax (Nmax - 1) calls
    assert groups[0].tolist() == [0, 6]               # (g=1,e=0,dir 0): images 0 and 2
    assert groups[-1].tolist() == [6 + 2 * 5 + 1]     # (g=3,e=2,dir 1): image 2 only
    assert sum(len(x) for x in groups) == len(row_img)
