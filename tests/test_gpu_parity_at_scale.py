"""Float parity of the DEFAULT (shared-footprint) pipeline against the fp32 oracle AT THE BENCHMARKED CONFIGURATIONS
(VERDICT r1 weak 1/2): the exact cfg2 bench batch (64 images x 40 boxes, 99 840 directed pairs, batch skip rule), a cfg3
bench batch (8 images x 100 proposals) and cfg5 (SGB tail, 64 x 40), each on >= 512 directed pairs drawn evenly from
geometric strata (no shared cell / 1-2 / 3-15 / >= 16 shared cells / boxes on the image border / empty boxes) so the
sorted-row, k_masks, multi-chunk and per-box-map machinery is what is being compared - not the repo's own dense path.

Bars.  fp16 operands (`PackedHead(operand_dtype=torch.float16)`, the bench default): joint probabilities within 2e-3 ABSOLUTE of the
fp32 oracle, directly, on BOTH weight presets.  bf16 operands: "trained" weights (round-1 trained-scale head, logit std 1.1): joint probabilities within 2e-3 ABSOLUTE of the fp32
oracle, directly (north_star).  "sharp" weights (He-gain trunk, `pred` O(1), logit std 3.3): operand rounding to bf16 alone
moves a probability by up to ~1e-2 in ANY bf16-in / fp32-accumulate implementation (oracle.parity.operand_rounded_scores is the
reference formulation with only that rounding applied), so there the kernels are held to that implementation-independent model
(within 2e-3 of it) and the distance to fp32 is recorded and bounded by the model's own distance.  Every run prints
max |dP|, max relative log-prob error and the per-super argmax flip rate."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import parity as PA
from oracle import sgb_oracle as SO
from scene_graph_commonsense_b200 import synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda"
PROB_TOL = 2e-3
N_SAMPLE = int(os.environ.get("HC_PARITY_SAMPLE", "512"))
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _record(name, stats):
    print("PARITY %s %s" % (name, json.dumps(stats)))
    try:
        os.makedirs(OUT, exist_ok=True)
        with open(os.path.join(OUT, "parity_at_scale.jsonl"), "a") as f:
            f.write(json.dumps(dict(case=name, **stats)) + "\n")
    except OSError:
        pass


_HEADS = {}


def _head(preset, operands="bf16"):
    from scene_graph_commonsense_b200 import model
    if ("sd", preset) not in _HEADS:
        _HEADS.clear()                                             # one preset resident at a time (276.7 M parameters)
        _HEADS[("sd", preset)] = synthetic.preset_state_dict(preset)
    sd = _HEADS[("sd", preset)]
    if operands not in _HEADS:
        _HEADS[operands] = model.PackedHead(sd, DEV, operand_dtype=torch.float16 if operands == "fp16" else torch.bfloat16)
    return sd, _HEADS[operands]


def _run_relation_case(name, samples, sgdet, preset, chunk_pairs, min_strata=4, operands="bf16", dense_above=None, expect_path="shared",
                       bar=None):
    """bar "direct": 2e-3 absolute against the fp32 oracle; "model": no further from fp32 than the operand-rounding model of the
    reference formulation in the same 16-bit format.  Default: direct for fp16 operands and for the trained preset."""
    bar = bar or ("direct" if (preset == "trained" or operands == "fp16") else "model")
    from scene_graph_commonsense_b200 import pipeline
    sd, packed = _head(preset, operands)
    kw = {} if dense_above is None else dict(dense_above=dense_above)
    pipe = pipeline.RelationPipeline(packed, DEV, commonsense=True, chunk_pairs=chunk_pairs, predcls=not sgdet, **kw)   # bench defaults
    assert pipe.fc1_shared and pipe.conv3_shared and pipe.conv3_block_rows == 4 and pipe.conv3_block_cols == 4
    b = pipeline.host_batch_from_samples(samples, skip_mode="batch", sgdet=sgdet).to_device(DEV)
    pairs = pipe.enumerate_pairs(b)
    rel, sup, conn, logsig = pipe.forward_pairs(b, pairs)
    torch.cuda.synchronize()
    assert pipe.last_path == expect_path, (pipe.last_path, b.cover_fraction)
    sub, obj, img = pairs["sub"].cpu().numpy(), pairs["obj"].cpu().numpy(), pairs["img"].cpu().numpy()
    boxes = b.boxes.cpu().numpy()
    off = b.box_offsets.cpu().numpy()
    idx, strata = PA.stratified_pair_sample(boxes, sub, obj, N_SAMPLE, seed=1)
    assert len(idx) >= min(N_SAMPLE, pairs["n"]) * 0.95
    pair_list = [(int(img[p]), int(sub[p] - off[img[p]]), int(obj[p] - off[img[p]])) for p in idx]
    rel_g = rel[torch.from_numpy(idx).to(DEV)].cpu().numpy()
    sup_g = sup[torch.from_numpy(idx).to(DEV)].cpu().numpy()
    torch.set_num_threads(os.cpu_count() or 8)
    rel_ref, sup_ref, _ = PA.oracle_scores(samples, sd, pair_list, sgdet=sgdet)
    st = PA.parity_stats(rel_g, rel_ref)
    st["strata"] = {int(k): int((strata == k).sum()) for k in np.unique(strata)}
    st["super_max_abs_dp"] = float(np.abs(np.exp(sup_g.astype(np.float64)) - np.exp(sup_ref.astype(np.float64))).max())
    st["n_pairs_batch"], st["preset"], st["operands"] = int(pairs["n"]), preset, operands
    if bar == "direct":
        _record(name, st)
        assert len(st["strata"]) >= min_strata, st["strata"]
        assert st["max_abs_dp"] <= PROB_TOL, st
        assert st["super_max_abs_dp"] <= PROB_TOL, st
        assert st["argmax_flip_rate"] <= 0.01, st
        return
    rel_emu, _, _ = PA.operand_rounded_scores(samples, sd, pair_list, sgdet=sgdet,
                                              dtype=torch.float16 if operands == "fp16" else torch.bfloat16)
    model_vs_fp32 = PA.parity_stats(rel_emu, rel_ref)
    ours_vs_model = PA.parity_stats(rel_g, rel_emu)
    st[operands + "_operand_model_vs_fp32"] = {k: model_vs_fp32[k] for k in ("max_abs_dp", "mean_abs_dp", "argmax_flip_rate")}
    st["ours_vs_%s_operand_model" % operands] = {k: ours_vs_model[k] for k in ("max_abs_dp", "mean_abs_dp", "argmax_flip_rate")}
    _record(name, st)
    assert st["top_joint_prob_median"] >= 0.35, st              # the weights really are sharp
    # our kernels are no further from fp32 than the operand-rounding model of the reference formulation is (worst case and on
    # average); the two differ from EACH OTHER by about as much (independent rounding decisions: recorded, not asserted)
    assert st["max_abs_dp"] <= 1.25 * model_vs_fp32["max_abs_dp"] + 5e-4, st
    assert st["mean_abs_dp"] <= 1.3 * model_vs_fp32["mean_abs_dp"] + 1e-4, st
    assert st["argmax_flip_rate"] <= model_vs_fp32["argmax_flip_rate"] + 0.01, st


CASES = [("trained", "bf16"), ("trained", "fp16"), ("sharp", "bf16"), ("sharp", "fp16")]      # grouped by preset: one set of weights at a time


@pytest.mark.parametrize("preset,operands", CASES)
def test_cfg2_bench_batch_default_pipeline_vs_fp32_oracle(preset, operands):
    import bench
    samples = bench.make_samples(0)                                  # the exact batch rank 0 benchmarks: 64 images x 40 boxes
    _run_relation_case("cfg2_%s_%s" % (preset, operands), samples, False, preset, bench.WORKLOADS["cfg2"]["chunk_pairs"], operands=operands)


@pytest.mark.parametrize("preset,operands", CASES)
def test_cfg3_bench_batch_default_pipeline_vs_fp32_oracle(preset, operands):
    import bench
    wl = bench.WORKLOADS["cfg3"]
    samples = bench.make_samples(0, wl["images"], wl["boxes"], sgdet=True)
    _run_relation_case("cfg3_%s_%s" % (preset, operands), samples, True, preset, wl["chunk_pairs"], operands=operands)


@pytest.mark.parametrize("preset,path", [("trained", "shared"), ("trained", "dense"), ("sharp", "shared"), ("sharp", "dense")])
def test_full_grid_boxes_match_on_both_paths(preset, path):
    """The other end of the box-size distribution (VERDICT weak 4): every box covers the whole grid, so every cell is shared and
    nothing can be skipped.  By default such a window takes the DENSE kernels (`dense_above`, host estimate 1.0 > 0.85); forced
    through the shared-footprint machinery (work lists = the dense tiling, every cell through the difference operand) it gives the
    same scores.  fp16 operands: bf16 sits AT the 2e-3 bar here on the trained weights (1.8e-3 joint / 2.2e-3 super, r02g).  With the
    sharp weights whole-grid boxes push the top joint probability to a median of 0.82 and even fp16 operand rounding alone
    reaches 4e-3 (dense and shared path alike, r02h), so that case is held to the fp16 operand-rounding model instead."""
    import bench
    samples = bench.make_samples(0, 4, 24, boxes_mode="full")
    _run_relation_case("full_boxes_%s_fp16_%s" % (preset, path), samples, False, preset, 16384, min_strata=1, operands="fp16",
                       dense_above=2.0 if path == "shared" else None, expect_path=path, bar="model" if preset == "sharp" else "direct")


@pytest.mark.parametrize("precision", ["bf16x3", "fp16", "bf16"])
def test_cfg5_sgb_tail_vs_fp32_oracle_at_bench_size(precision):
    """64 images x 40 objects (99 840 pairs), 9 random pairs per image (576 in all) through the fp32 restatement of
    roi_relation_predictors.py:399-459.  bf16x3 (split operands) and plain fp16 hold the 2e-3 bar; plain bf16 does not (recorded)."""
    from scene_graph_commonsense_b200 import sgb
    n_img, n_obj = 64, 40
    num_objs = [n_obj] * n_img
    batch = synthetic.make_sgb_batch(num_objs, seed=0)
    sd = synthetic.sgb_state_dict(seed=0)
    post_cat = torch.nn.Linear(1024, 4096).to(DEV)
    head = sgb.BayesHead(input_dim=4096).to(DEV)
    with torch.no_grad():
        post_cat.weight.copy_(sd["post_cat.weight"]); post_cat.bias.copy_(sd["post_cat.bias"])
        for n in ("fc3_1", "fc3_2", "fc3_3", "fc5"):
            getattr(head, n).weight.copy_(sd[n + ".weight"]); getattr(head, n).bias.copy_(sd[n + ".bias"])
    edge_rep = torch.nn.functional.linear(batch["edge_ctx"], sd["post_emb.weight"], sd["post_emb.bias"])
    pairs = SO.prepare_test_pairs(num_objs)
    r1, r2, r3, sup = sgb.hierarchical_relation_tail(edge_rep.to(DEV), [p.to(DEV) for p in pairs], num_objs, batch["obj_labels"].to(DEV),
                                                     batch["union_features"].to(DEV), post_cat, head, sd["freq_bias"].to(DEV),
                                                     precision=precision)
    rel_g = torch.cat((torch.cat(list(r1)), torch.cat(list(r2)), torch.cat(list(r3))), dim=1)
    st = PA.sgb_tail_parity(sd, batch, pairs, num_objs, rel_g)
    st["precision"] = precision
    _record("cfg5_" + precision, st)
    if precision in ("bf16x3", "fp16"):
        assert st["max_abs_dp"] <= PROB_TOL, st
    else:       # plain bf16 operands: recorded; must at least be a faithful bf16 GEMM (no gross error)
        assert st["max_abs_dp"] <= 0.05 and st["argmax_flip_rate"] <= 0.05, st
