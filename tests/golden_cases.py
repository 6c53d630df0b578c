"""Case tables shared by oracle/make_golden.py (which runs the real reference) and the tests that replay them."""

PREDCLS_CASES = {
    # name: (image ids, boxes per image, run_mode, hierar, score kwargs, batches)
    "pc_cs_single20": dict(ids=[0], n=[20], run_mode="eval_cs", hierar=True, kw=dict(gain=3.0)),
    "pc_cs_ragged": dict(ids=[10, 11, 12, 13, 14, 15], n=[2, 16, 7, 3, 12, 9], run_mode="eval_cs", hierar=True, kw=dict(gain=3.0)),
    "pc_cs_dense": dict(ids=[16, 17, 18, 19], n=[12, 5, 17, 9], run_mode="eval_cs", hierar=True, kw=dict(gain=3.0), cs=(3, 0.5, 0.1)),
    "pc_plain_ragged": dict(ids=[20, 21, 22, 23], n=[5, 11, 8, 14], run_mode="eval", hierar=True, kw=dict(gain=3.0)),
    "pc_plain_ties": dict(ids=[30, 31, 32], n=[9, 13, 6], run_mode="eval", hierar=True, kw=dict(gain=1.0, tie_quantum=0.5)),
    "pc_cs_ties": dict(ids=[33, 34], n=[15, 10], run_mode="eval_cs", hierar=True, kw=dict(gain=1.0, tie_quantum=1.0), cs=(4, 0.6, 0.05)),
    "pc_flat": dict(ids=[40, 41, 42], n=[8, 12, 4], run_mode="eval", hierar=False, kw=dict(gain=3.0)),
    "pc_flat_cs": dict(ids=[43, 44], n=[10, 6], run_mode="eval_cs", hierar=False, kw=dict(gain=3.0), cs=(5, 0.5, 0.1)),
    "pc_two_windows": dict(ids=[50, 51, 52, 53, 54, 55], n=[6, 9, 12, 5, 10, 7], run_mode="eval_cs", hierar=True,
                           kw=dict(gain=3.0), windows=[[0, 1, 2], [3, 4, 5]], cs=(6, 0.5, 0.1)),
    "pc_cfg2_slice": dict(ids=[60, 61], n=[40, 40], run_mode="eval_cs", hierar=True, kw=dict(gain=3.0)),
}


SGDET_CASES = {
    "sgd_cs": dict(ids=[70, 71, 72], n_gt=[6, 9, 4], n_prop=[14, 20, 9], run_mode="eval_cs", cs=(7, 0.5, 0.1)),
    "sgd_plain": dict(ids=[73, 74], n_gt=[8, 5], n_prop=[18, 12], run_mode="eval"),
    "sgd_nogt": dict(ids=[75, 76], n_gt=[2, 7], n_prop=[6, 15], run_mode="eval", p_rel=[0.0, 0.6]),
}

SGB_CASES = {
    "sgb_case_a": dict(num_objs=[6, 9, 4, 12], seed=0),
    "sgb_case_b": dict(num_objs=[3, 14, 2, 8, 5], seed=1, empty_gt=2),
}

SGB_VARIANT_CASE = dict(num_objs=[5, 8, 3, 10], seed=2)      # TransformerHierPredictor / VCTreeHierPredictor through the real forward

FRONTEND_CASES = {
    # DETR outputs -> proposals (evaluate.py:309-368), match_object_categories, match_target_sgd.
    # The reference hard-codes 100 queries (`.view(-1, 100, topk_cat)`, evaluate.py:311), so raggedness comes from p_noobj.
    "fe_dense": dict(ids=[80, 81, 82], n_gt=[9, 14, 6], queries=100),
    "fe_ragged": dict(ids=[83, 84, 85, 86], n_gt=[3, 12, 7, 2], queries=100, kw=dict(p_noobj=0.8, p_dup=0.45)),
    "fe_sparse": dict(ids=[87, 88], n_gt=[5, 4], queries=100, kw=dict(p_noobj=0.96, p_dup=0.1, p_second_noobj=0.6)),
}

TRAIN_CASES = {
    # training-side losses on the hierarchical head: train_utils.train_one_direction driven by the loop body of
    # train_test.py:187-258.  Inputs are regenerated from seeds (synthetic.make_batch, synthetic.head_state_dict, seeded hidden
    # vectors); the golden files hold the reference's per-call losses, the step loss and its autograd gradients.
    "tr_hier_cs": dict(ids=[90, 91, 92], n=[5, 3, 6], run_mode="train_cs", hierar=True, cs=(8, 0.5, 0.1), gain=3.0),
    "tr_hier_plain": dict(ids=[93, 94, 95, 96], n=[4, 7, 2, 6], run_mode="train", hierar=True, gain=3.0),
    "tr_hier_temps": dict(ids=[97, 98, 99], n=[6, 6, 4], run_mode="train_cs", hierar=True, cs=(9, 0.3, 0.3), gain=2.0,
                          temps=(2.0, 0.5, 1.5), p_rel=0.8),
    "tr_hier_sparse": dict(ids=[100, 101], n=[5, 8], run_mode="train_cs", hierar=True, cs=(10, 0.9, 0.0), gain=3.0, p_rel=0.1),
    "tr_flat_cs": dict(ids=[102, 103, 104], n=[6, 4, 5], run_mode="train_cs", hierar=False, cs=(11, 0.5, 0.1), gain=3.0),
    "tr_flat_plain": dict(ids=[105, 106], n=[7, 3], run_mode="train", hierar=False, gain=1.0),
}


# Head / pipeline goldens through the REAL reference modules end to end (BayesianRelationClassifier -> evaluate_one_direction
# -> Evaluator / Evaluator_Top3), with object classes forced to cover 1, 2 and 3 super-classes (the reference's
# process_super_class adds only the first and the last entry of a list), an empty box, ragged sizes and the batch skip rule.
REAL_PIPELINE_CASE = dict(ids=[110, 111, 112], n=[5, 3, 4],
                          cats=[[7, 12, 0, 2, 20], [8, 3, 15], [25, 1, 101, 16]], run_mode="eval_cs", p_rel=0.7,
                          empty_box=(1, 2), presets=("trained", "sharp"))
HEAD3_CASE = dict(id=902, n=6, cats=[7, 12, 0, 2, 20, 3], pairs=[(0, 1), (1, 0), (2, 0), (0, 3), (3, 4), (5, 1), (4, 5), (2, 3)],
                  presets=("trained", "sharp"))


def real_pipeline_samples(case=None):
    """The samples of REAL_PIPELINE_CASE (shared by the generator and the tests that replay it)."""
    import torch
    from scene_graph_commonsense_b200 import synthetic
    case = case or REAL_PIPELINE_CASE
    samples = [synthetic.with_categories(synthetic.make_image(i, n, p_rel=case["p_rel"]), c)
               for i, n, c in zip(case["ids"], case["n"], case["cats"])]
    if case.get("empty_box"):
        i, j = case["empty_box"]
        samples[i].bbox[j] = torch.tensor([5, 5, 9, 12], dtype=samples[i].bbox.dtype)      # xmin == xmax: an all-zero mask
    return samples


def install_gt(samples, flat):
    """Writes a flat relationships array (image-major, rows g = 1..N-1, e < g) back into the samples."""
    import torch
    k = 0
    for s in samples:
        for gi in range(1, len(s.categories)):
            s.relationships[gi - 1] = torch.as_tensor(flat[k:k + gi], dtype=s.relationships[gi - 1].dtype)
            k += gi
    assert k == len(flat)
    return samples
