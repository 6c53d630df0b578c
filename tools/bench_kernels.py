#!/usr/bin/env python
"""Per-kernel throughput of the HBM/latency-bound stages of the path (north_star: "pair gather, hierarchical softmax,
top-K and the triplet hash-filter ... reported as achieved HBM GB/s against peak").

For each kernel: ALGORITHMIC bytes (each input read once, each output written once - DESIGN.md §3) / CUDA-event time
(median of `--iters` launches after warm-up, events on the launching stream), at BASELINE size (cfg2: 64 images x 40
boxes, 99 840 pairs - a few MB, i.e. launch/latency-sized) and at a scaled size (`--scale-images`, default 2048 images)
where the kernels have enough bytes to approach bandwidth.  Prints one JSON object; run on a GPU box:

    python tools/bench_kernels.py > gpurun_out/kernels_rNN.json
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from scene_graph_commonsense_b200 import ops, pipeline, synthetic, tables  # noqa: E402


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def timeit(fn, iters, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()                      # > L2 (126 MB): evicts the kernel's inputs between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def window(n_images, n_boxes, dev):
    """Synthetic PredCLS window without feature maps, tiled from 64 distinct images (CSR arrays on the device)."""
    base = synthetic.make_batch(list(range(min(n_images, 64))), n_boxes, with_maps=False, p_rel=0.3)
    samples = [base[i % len(base)] for i in range(n_images)]
    return pipeline.batch_from_samples(samples, dev, skip_mode="batch", group_size=64, with_maps=False)


def run_size(n_images, n_boxes, dev, iters, peak, results, tag):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    b = window(n_images, n_boxes, dev)
    pipe = pipeline.RelationPipeline(None, dev, commonsense=True)
    n_box = b.boxes.shape[0]
    n_tri = int(b.tri_offsets[-1])

    def rec(name, bytes_alg, fn, note=""):
        med, best = timeit(fn, iters, flush)
        results.append({"kernel": name, "size": tag, "algorithmic_bytes": int(bytes_alg), "ms_median": med, "ms_min": best,
                        "gbs": bytes_alg / (med * 1e-3) / 1e9, "frac_of_hbm_peak": bytes_alg / (med * 1e-3) / 1e9 / peak, "note": note})

    # R1/R2/R4 pair enumeration: read boxes 16 B/box + rel/dir 5 B/triangle; write 21 B per directed pair (sub,obj,img,gt,rel,ov)
    pairs = pipe.enumerate_pairs(b)
    P = pairs["n"]
    rec("pairs_enumerate (4 kernels + 1 D2H of offsets)", n_box * 16 + n_tri * 5 + P * 21, lambda: pipe.enumerate_pairs(b),
        "includes the [B+1]-int D2H read that sizes the dense launches")

    # shared-footprint bookkeeping: sort keys, per-tile K-cell masks, conv3_1 work list of one chunk, operand zero fill
    out_fp = {}

    def keys():
        out_fp["keys"] = ops.pair_cell_keys(b.boxes, pairs["sub"], pairs["obj"])

    rec("pair_cell_keys_kernel", P * 12 + n_box * 16, keys, "cell rectangle both boxes reach -> sort key; read 2 box ids, write 1 key per pair")
    perm = torch.sort(out_fp["keys"], stable=True)[1]
    row_sub, row_obj = pairs["sub"][perm].contiguous(), pairs["obj"][perm].contiguous()

    def masks():
        out_fp["masks"] = ops.tile_cell_masks(b.boxes, row_sub, row_obj, 256)

    rec("tile_cell_masks_kernel", P * 8 + (P // 256 + 1) * 8 + n_box * 16, masks, "one warp per 256-row GEMM tile: OR of the rows' cell masks")
    cn = min(P, 12480)
    blk = torch.empty(cn * 16, dtype=torch.int32, device=dev)
    nb = torch.zeros(1, dtype=torch.int32, device=dev)
    ops.conv3_shared_blocks(b.boxes, pairs["sub"][:cn], pairs["obj"][:cn], 4, 32, blocks=blk, n_blocks=nb, block_cols=4)
    rec("conv3_blocks_kernel (shared footprint, 4x4-pixel blocks, one 12 480-pair chunk)", cn * 8 + int(nb.item()) * 4,
        lambda: ops.conv3_shared_blocks(b.boxes, pairs["sub"][:cn], pairs["obj"][:cn], 4, 32, blocks=blk, n_blocks=nb, block_cols=4),
        "ONE CTA (deterministic order, device-side count): latency-sized, runs on the pooling stream")
    if P <= 200000:
        d = torch.empty(P, 64, 1024, dtype=torch.bfloat16, device=dev)
        cells = int(np.unpackbits(out_fp["masks"].cpu().numpy().view(np.uint8)).sum())
        rec("cells_zero_kernel", cells * 256 * 2048, lambda: ops.cells_zero(out_fp["masks"], 256, P, d),
            "zero fill of the visited cells of the fc1 difference operand (%d cells x 256 rows x 2 KB)" % cells)
        del d
    del blk, row_sub, row_obj, perm

    # R6/R7 hierarchical head: read raw 2 KiB/pair + labels, write relation 200 + super 12 + conn 4 + logsig 4
    sd = synthetic.head_state_dict(seed=0, logit_gain=40.0)
    f32 = lambda t: t.detach().to(dev, torch.float32).contiguous()
    w_fc2 = f32(sd["fc2.weight"])
    emb = w_fc2[:, 4096:].t().contiguous()
    b_fc2 = f32(sd["fc2.bias"])
    w_heads = torch.cat([f32(sd[k + ".weight"]) for k in ("fc3_1", "fc3_2", "fc3_3", "fc4", "fc5")]).contiguous()
    b_heads = torch.cat([f32(sd[k + ".bias"]) for k in ("fc3_1", "fc3_2", "fc3_3", "fc4", "fc5")]).contiguous()
    del sd, w_fc2
    raw = torch.randn(P, 512, device=dev)
    out = {}

    def head():
        out["h"] = ops.hier_head(raw, b_fc2, emb, pairs["sub"], pairs["obj"], b.cats, b.supers, w_heads, b_heads, (15, 11, 24))

    rec("hier_head_kernel", P * (2048 + 8 + 220), head, "fc2 bias + label-embedding add + ReLU + 54-row heads GEMV + hierarchical log-softmax")
    # arithmetic intensity 2*512*54 FLOP / 2.3 KB = 24 FLOP/B is above the fp32 SIMT ridge (72 TFLOP/s / 6.5 TB/s = 11 FLOP/B):
    # this kernel is bound by the fp32 FMA pipe, not by HBM - report it against that peak too
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
    r = results[-1]
    r["fp32_tflops"] = P * 2 * 512 * 54 / (r["ms_median"] * 1e-3) / 1e12
    r["frac_of_fp32_simt_peak"] = r["fp32_tflops"] / fp32_peak
    r["bound"] = "fp32 SIMT (148 SMs x 128 FMA/clk x 1.965 GHz = %.1f TFLOP/s nominal)" % fp32_peak
    relation, sup, conn, logsig, _ = out["h"]
    del raw

    # R8/R9/R10 candidates + commonsense bitmap filter: 254 B/pair
    def cand():
        out["c"] = ops.candidates(relation, (15, 11, 24), True, pairs["ov"], logsig, pairs["sub"], pairs["obj"], b.cats, pipe.pass_bitmap,
                                  sup, want_top3=True)

    rec("candidates_kernel (+bitmap filter)", P * (200 + 12 + 4 + 1 + 8 + 3 * 8 + 5), cand,
        "per-super max/argmax, overlap mask, 1 125 000-bit commonsense bitmap lookup, + log sigma(conn)")
    cand_conf, cand_label, t3_conf, t3_super = out["c"]

    # R10-R12 top-100 + match + counters: read 8 B/candidate + 16 B/pair of ids + GT slots 4 B/pair
    cand_offsets = (pairs["offsets"] * 3).contiguous()
    counters = torch.zeros(tables.EV_SIZE, dtype=torch.int64, device=dev)

    def topk():
        ops.topk_match(cand_offsets, cand_conf, cand_label, 3, pairs["sub"], pairs["obj"], b.cats, b.boxes, pairs["offsets"], pairs["gt"],
                       pairs["sub"], pairs["obj"], b.cats, b.boxes, counters, zs_bitmap=pipe.zs_bitmap, mode=0)

    rec("topk_match_kernel (Evaluator)", P * (3 * 8 + 8 + 4), topk, "one CTA per image: radix select + bitonic top-100 + GT scan + counters")

    def topk3():
        ops.topk_match(pairs["offsets"], t3_conf, None, 1, pairs["sub"], pairs["obj"], b.cats, b.boxes, pairs["offsets"], pairs["gt"],
                       pairs["sub"], pairs["obj"], b.cats, b.boxes, torch.zeros(tables.T3_SIZE, dtype=torch.int64, device=dev), mode=1,
                       t3_labels=cand_label, t3_super=t3_super)

    rec("topk_match_kernel (Evaluator_Top3)", P * (4 + 12 + 1 + 8 + 4), topk3)

    # connectivity statistics: 12 B/pair
    stats = torch.zeros(5, dtype=torch.int64, device=dev)
    rec("conn_stats_kernel", P * 12, lambda: ops.connectivity_stats(conn, pairs["gt"], pairs["rel"], stats))
    del relation, sup, conn, logsig, cand_conf, cand_label
    return P


def run_gather(dev, iters, peak, results):
    """R3 stages that touch activations (box_select, tiled pair pooling) at cfg2 size with real shapes."""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    n_img, n_boxes = 64, 40
    b = window(n_img, n_boxes, dev)
    pipe = pipeline.RelationPipeline(None, dev, commonsense=False)
    pairs = pipe.enumerate_pairs(b)
    n_box = b.boxes.shape[0]
    t = torch.randn(n_img * 1024, 256, device=dev).to(torch.bfloat16)
    fill = torch.randn(256, device=dev).to(torch.bfloat16)
    med, best = timeit(lambda: ops.box_select(t, b.boxes, b.box_img, fill, 32), iters, flush)
    by = n_box * 1024 * 256 * 2 + n_img * 1024 * 256 * 2
    results.append({"kernel": "box_select_kernel", "size": "cfg2", "algorithmic_bytes": by, "ms_median": med, "ms_min": best,
                    "gbs": by / (med * 1e-3) / 1e9, "frac_of_hbm_peak": by / (med * 1e-3) / 1e9 / peak,
                    "note": "write 512 KiB per box, read each image map once"})
    u = torch.randn(n_box, 32, 32, 512, device=dev).to(torch.bfloat16)
    v = torch.randn(n_box, 32, 32, 512, device=dev).to(torch.bfloat16)
    b2 = torch.randn(512, device=dev)
    lut = ops.pair_lut_build(pairs["sub"], pairs["obj"], pairs["img"], b.box_offsets, n_box, n_boxes)
    off = pairs["offsets_host"]
    n_chunk_img = 10
    cnt = int(off[n_chunk_img] - off[0])
    outb = torch.empty(cnt, 16, 16, 512, dtype=torch.bfloat16, device=dev)
    med, best = timeit(lambda: ops.pair_relu_pool_tiled(u, v, None, b.box_offsets, lut, 0, n_chunk_img, 0, cnt, 32, out=outb), iters, flush)
    by = cnt * 16 * 16 * 512 * 2 + 2 * n_chunk_img * n_boxes * 1024 * 512 * 2
    results.append({"kernel": "pair_relu_pool_tiled_bf16_kernel", "size": "cfg2 chunk: %d images, %d pairs" % (n_chunk_img, cnt),
                    "algorithmic_bytes": by, "ms_median": med, "ms_min": best, "gbs": by / (med * 1e-3) / 1e9,
                    "frac_of_hbm_peak": by / (med * 1e-3) / 1e9 / peak,
                    "note": "write 256 KiB per pair + read U,V of the chunk's boxes once (outer-sum tiling)"})


def run_frontend(dev, iters, peak, results, n_images):
    from scene_graph_commonsense_b200 import frontend
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    base = synthetic.make_batch(list(range(16)), 20, with_maps=False)
    lg, bx = synthetic.make_detr_outputs(base, num_queries=100)
    rep = (n_images + 15) // 16
    out_dict = {"pred_logits": lg.repeat(rep, 1, 1)[:n_images].contiguous().to(dev), "pred_boxes": bx.repeat(rep, 1, 1)[:n_images].contiguous().to(dev)}
    p = frontend.detr_proposals(out_dict)
    by = n_images * 100 * (151 * 4 + 16) + p.n * (4 + 4 + 16 + 16 + 4 + 4)
    med, best = timeit(lambda: frontend.detr_proposals(out_dict), iters, flush)
    results.append({"kernel": "detr_proposals (expand + per-class NMS + scan + pack, 1 D2H)", "size": "%d images x 100 queries" % n_images,
                    "algorithmic_bytes": by, "ms_median": med, "ms_min": best, "gbs": by / (med * 1e-3) / 1e9,
                    "frac_of_hbm_peak": by / (med * 1e-3) / 1e9 / peak, "note": "%d proposals kept" % p.n})


def run_train(n_images, dev, iters, peak, results, tag):
    """N4 training-side kernels: per-call losses + d_logits (450 B/row), head backward (d_logits + pred in, d_pred out: 4 312 B/row)."""
    from scene_graph_commonsense_b200 import losses
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    base = synthetic.make_batch(list(range(64)), 40, with_maps=False, p_rel=0.3)
    samples = [base[i % 64] for i in range(n_images)]
    rows = losses.training_rows(samples, dev, group_size=64)
    args = synthetic.reference_args(run_mode="train_cs", hierar=True)
    crit = losses.RelationLoss(args, dev)
    sd = synthetic.head_state_dict(seed=0, logit_gain=3.0)
    f32 = lambda t: t.detach().to(dev, torch.float32).contiguous()
    names = ("fc3_1", "fc3_2", "fc3_3", "fc4", "fc5")
    w_heads = torch.cat([f32(sd[k + ".weight"]) for k in names]).contiguous()
    b_heads = torch.cat([f32(sd[k + ".bias"]) for k in names]).contiguous()
    del sd
    P = rows.n_rows
    pred = torch.relu(torch.randn(P, 512, device=dev))
    relation, sup, conn, _, _ = ops.hier_head(pred, None, None, None, None, None, None, w_heads, b_heads, (15, 11, 24))
    out = {}

    def loss():
        out["l"] = ops.hier_loss(relation, sup, conn, rows.row_target, rows.group_offsets, rows.group_rows, rows.group_weight,
                                 crit.class_weight, (15, 11, 24), True, (1.0, 1.0, 1.0), crit.aligned, crit.violated, rows.row_sub,
                                 rows.row_obj, rows.box_cat, crit.lambdas)

    def rec(name, bytes_alg, fn, note=""):
        med, best = timeit(fn, iters, flush)
        results.append({"kernel": name, "size": tag, "algorithmic_bytes": int(bytes_alg), "ms_median": med, "ms_min": best,
                        "gbs": bytes_alg / (med * 1e-3) / 1e9, "frac_of_hbm_peak": bytes_alg / (med * 1e-3) / 1e9 / peak, "note": note})

    rec("hier_loss (memset + hier_loss_kernel + loss_total_kernel)", P * 450, loss,
        "%d calls; per-call commonsense penalty + connectivity BCE + hierarchical NLL and d(step loss)/d(logits)" % rows.n_groups)
    d_logits = out["l"][2]
    rec("hier_head_bwd (d_pred + two-stage d_W/d_b)", P * 4312, lambda: ops.hier_head_bwd(d_logits, pred, w_heads),
        "fp32 SIMT: 2 x 2*54*512 FLOP per row")
    r = results[-1]
    r["fp32_tflops"] = P * 4 * 512 * 54 / (r["ms_median"] * 1e-3) / 1e12
    rec("hier_head_bwd: d_pred only", P * (216 + 2048), lambda: ops.hier_head_bwd(d_logits, pred, w_heads, want_weights=False))
    results[-1]["fp32_tflops"] = P * 2 * 512 * 54 / (results[-1]["ms_median"] * 1e-3) / 1e12
    rec("hier_head_bwd: d_W / d_b only", P * (216 + 2048), lambda: ops.hier_head_bwd(d_logits, pred, w_heads, want_pred=False))
    results[-1]["fp32_tflops"] = P * 2 * 512 * 54 / (results[-1]["ms_median"] * 1e-3) / 1e12
    return P


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--scale-images", type=int, default=2048)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    peak, src = peak_hbm()
    results = []
    p_small = run_size(64, 40, dev, args.iters, peak, results, "cfg2 (64 images x 40 boxes)")
    p_big = run_size(args.scale_images, 40, dev, max(args.iters // 2, 5), peak, results, "scaled (%d images x 40 boxes)" % args.scale_images)
    run_gather(dev, args.iters, peak, results)
    run_frontend(dev, args.iters, peak, results, 64)
    run_frontend(dev, args.iters, peak, results, 4096)
    run_train(64, dev, args.iters, peak, results, "training step, 64 images x 40 boxes (99 840 rows, 1 560 calls)")
    run_train(args.scale_images, dev, max(args.iters // 2, 5), peak, results, "training, %d images x 40 boxes" % args.scale_images)
    print(json.dumps({"hbm_peak_gbs": peak, "peak_source": src, "pairs_cfg2": p_small, "pairs_scaled": p_big,
                      "timing": "CUDA events on the launching stream, median of N launches after 3 warm-ups, 256 MB L2 flush between launches",
                      "kernels": results}, indent=1))


if __name__ == "__main__":
    main()
