#!/usr/bin/env python
"""Config 5 (BASELINE.json): plug-and-play hierarchical head on the Scene-Graph-Benchmark Motifs predictor tail
(PredCLS, 512-d context, 4096-d union features, 51 classes) - throughput of the SGB twin on one B200.

A step = roi_relation_predictors.py:400-469 tail + inference.py:246-302 candidates/ranking + sgg_eval matching for one
window of 64 synthetic images x 40 objects (99 840 directed pairs).  Reports pairs/s with the inputs resident in HBM,
ALGORITHMIC GEMM TFLOP/s (8 830 976 FLOP/pair; the bf16x3 split executes 3x that on the tensor pipe) and the
CUDA-event time of each tagged launch.  One JSON object on stdout.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scene_graph_commonsense_b200 import ops, sgb, synthetic  # noqa: E402

FLOP_PAIR = 2 * 1024 * 4096 + 2 * 4096 * 54


def main():
    steps, warmup = int(os.environ.get("STEPS", "10")), 3
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    n_img, n_obj = 64, 40
    num_objs = [n_obj] * n_img
    batch = synthetic.make_sgb_batch(num_objs, seed=0)
    sd = synthetic.sgb_state_dict(seed=0)
    post_cat = torch.nn.Linear(1024, 4096).to(dev)
    head = sgb.BayesHead(input_dim=4096).to(dev)
    with torch.no_grad():
        post_cat.weight.copy_(sd["post_cat.weight"]); post_cat.bias.copy_(sd["post_cat.bias"])
        for n in ("fc3_1", "fc3_2", "fc3_3", "fc5"):
            getattr(head, n).weight.copy_(sd[n + ".weight"]); getattr(head, n).bias.copy_(sd[n + ".bias"])
    edge_rep = torch.nn.functional.linear(batch["edge_ctx"], sd["post_emb.weight"], sd["post_emb.bias"]).to(dev)   # upstream of the path
    pairs = [torch.nonzero(torch.ones(n, n) - torch.eye(n)).view(-1, 2).to(dev) for n in num_objs]
    obj_labels = batch["obj_labels"].to(dev)
    union = batch["union_features"].to(dev)
    freq = sd["freq_bias"].to(dev)
    logits = list(batch["obj_logits"].to(dev).split(num_objs))
    boxes = [b.to(dev) for b in batch["boxes"]]
    g = torch.Generator().manual_seed(7)
    gt_rels = []
    for n in num_objs:
        k = 12
        so = torch.stack([torch.randperm(n, generator=g)[:2] for _ in range(k)])
        gt_rels.append(torch.cat((so, torch.randint(1, 51, (k, 1), generator=g)), dim=1).to(dev))
    gt_classes = list(obj_labels.split(num_objs))
    post = sgb.HierarchPostProcessor(use_gt_box=True)
    n_pairs = sum(n * (n - 1) for n in num_objs)

    def step():
        r1, r2, r3, sup = sgb.hierarchical_relation_tail(edge_rep, pairs, num_objs, obj_labels, union, post_cat, head, freq)
        cand = post.candidates(r1, r2, r3, logits, pairs)
        rec = sgb.SGBRecall()
        rec.evaluate_batch(cand, gt_rels, gt_classes, boxes)
        return rec

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    ops.PROFILE["events"].clear()
    ops.PROFILE["on"] = True
    l0 = ops.LAUNCHES["n"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        rec = step()
    e1.record()
    torch.cuda.synchronize()
    ops.PROFILE["on"] = False
    ms = e0.elapsed_time(e1) / steps
    tags = {}
    for tag, a, b in ops.PROFILE["events"]:
        tags.setdefault(tag, []).append(a.elapsed_time(b))
    gemm_ms = sum(float(np.sum(v)) for t, v in tags.items() if t in ("post_cat", "bayes_head")) / steps
    res = rec.result()
    print(json.dumps({
        "workload": "cfg5: SGB Motifs PredCLS tail, %d images x %d objects (%d directed pairs/step), 4096-d union features, 51 classes" % (n_img, n_obj, n_pairs),
        "pairs_per_sec": n_pairs / (ms * 1e-3), "ms_per_step": ms, "steps": steps, "gpu_launches_per_step": (ops.LAUNCHES["n"] - l0) / steps,
        "gemm": {"ms_per_step": gemm_ms, "algorithmic_tflops": n_pairs * FLOP_PAIR / (gemm_ms * 1e-3) / 1e12,
                 "executed_tflops_bf16x3": 3 * n_pairs * FLOP_PAIR / (gemm_ms * 1e-3) / 1e12,
                 "note": "bf16x3 split operands (A_hi*W_hi + A_lo*W_hi + A_hi*W_lo) keep the 4096-long dot products at ~fp32 accuracy"},
        "kernel_breakdown_ms": {t: float(np.sum(v)) / steps for t, v in sorted(tags.items())},
        "recall": {str(k): float(v) for k, v in res["recall"].items()},
    }))


if __name__ == "__main__":
    main()
