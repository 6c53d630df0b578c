#!/usr/bin/env python
"""Benchmark of the HIERCOM relation-prediction hot path (BASELINE.json metric: directed relation pairs / second).

  python bench.py --gpus N --steps K --warmup W            our arm   (sm_100a kernels through the public pipeline API)
  python bench.py --impl reference --gpus N --steps K ...  reference arm: the reference's CPU formulation (oracle port,
                                                           torch fp32 on all host cores) on a bounded sample per step

A step = one pass of the path R1-R13 over one batch: cfg2 of BASELINE.json, 64 synthetic VG-shaped images x 40 boxes
per GPU (99 840 directed pairs, two-pass), PredCLS, eval_cs (commonsense filter on), reference batch skip rule.
N > 1: one process per GPU (torchrun), every rank owns its own 64 images (weak scaling), one int64 all-reduce of the
765-slot counter vector per step.  Timing: CUDA events on the launching stream, barrier + synchronize on both sides,
max over ranks.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMAGES_PER_GPU = 64
BOXES = 40
FLOP_IMG = 134_742_016            # conv1_1 + conv1_2 once per image
FLOP_BOX = 2_415_919_104          # subject-half + object-half of conv2_1 once per box
FLOP_PAIR_CONV3 = 2_415_919_104
FLOP_PAIR_FC1 = 536_870_912          # 2 * 65536 * 4096
FLOP_PAIR = 2_957_039_616         # conv3_1 + fc1 + fc2(dense) + heads per directed pair


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_sustained=d["bf16_tflops_sustained"], bf16_burst=d["bf16_tflops"], hbm=d["hbm_gbs"], source="measured")
    return dict(bf16_sustained=1400.0, bf16_burst=1590.0, hbm=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# --conv3: dense kernel; block-sparse over the cells EITHER box of a pair reaches (8x8 / 8x4-pixel blocks); or "shared": per pair
# only the cells BOTH boxes reach, the rest taken from per-box maps computed once per box (block_rows, shared)
CONV3_MODES = {"dense": (0, False, 8), "blocks8": (8, False, 8), "blocks4": (4, False, 8), "shared8": (8, True, 8), "shared4": (4, True, 8),
               "shared44": (4, True, 4)}       # name: (block rows, shared footprint, block columns) in conv3 pixels

WORKLOADS = {
    # name: images per GPU, boxes (proposals) per image, SGDET-style?, pair chunk
    "cfg2": dict(images=IMAGES_PER_GPU, boxes=BOXES, sgdet=False, chunk_pairs=16384,
                 text="cfg2: PredCLS %d images x %d boxes per GPU (%d directed pairs/GPU/step), two-pass, eval_cs, reference batch skip rule"),
    "cfg3": dict(images=8, boxes=100, sgdet=True, chunk_pairs=40960,
                 text="cfg3: SGDET-style %d images x %d proposals per GPU (%d directed pairs/GPU/step, 20 GT boxes/image), two-pass, "
                      "object-confidence add, synonym matching, top-100 triplets, eval_cs, reference batch skip rule"),
}


def make_samples(rank, n_images=IMAGES_PER_GPU, boxes=BOXES, with_maps=True, sgdet=False, boxes_mode="small"):
    from scene_graph_commonsense_b200 import synthetic
    ids = [rank * n_images + i for i in range(n_images)]
    if sgdet:
        return [synthetic.make_sgdet_image(i, 20, boxes, base_seed=0, p_rel=0.3, with_maps=with_maps, box_mode=boxes_mode) for i in ids]
    return synthetic.make_batch(ids, boxes, base_seed=0, p_rel=0.3, with_maps=with_maps, box_mode=boxes_mode)


# ======================================================================================================= reference arm
def cpu_reference_run(steps, warmup, budget_s=150.0, per_pair_s=0.024):
    """The reference's CPU formulation (oracle port of evaluate.py:111-217 + model.py + evaluator.py, torch fp32, all host
    threads) on a bounded sample of cfg2: the first `imgs` images restricted to their first `nb` boxes."""
    from oracle import hiercom_oracle as O
    from scene_graph_commonsense_b200 import synthetic, tables
    torch.set_num_threads(os.cpu_count() or 1)
    total_steps = max(steps + warmup, 1)
    pairs_budget = max(budget_s / total_steps / per_pair_s, 24)
    imgs, nb = 2, 4
    while imgs * (nb + 1) * nb <= pairs_budget and nb < BOXES:
        nb += 1
    full = make_samples(0, imgs, BOXES)
    batch = []
    for s in full:
        batch.append(synthetic.ImageSample(s.image_id, s.feat, s.depth, s.bbox[:nb], s.categories[:nb], s.super_categories[:nb],
                                           s.relationships[:nb - 1], s.subj_or_obj[:nb - 1]))
    sd = synthetic.head_state_dict(seed=0, logit_gain=40.0)
    head_fn = O.make_head_fn(sd)
    zs = set(tables.zero_shot_keys().tolist())
    al, vi = set(tables.commonsense_aligned_keys().tolist()), set(tables.commonsense_violated_keys().tolist())
    times, pairs = [], 0
    for it in range(total_steps):
        ev = O.OracleEvaluator((15, 11, 24), True, aligned=al, violated=vi, zero_shot=zs)
        t3 = O.OracleEvaluatorTop3((15, 11, 24))
        t0 = time.perf_counter()
        pairs = O.replay_predcls(batch, head_fn, ev, t3)
        ev.compute(per_class=True)
        t3.compute(per_class=True)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    mean_t = float(np.mean(times))
    sample = "%d images x first %d of %d boxes (%d directed pairs/step), replay of evaluate.py:111-217 incl. Evaluator+Top3" % (
        imgs, nb, BOXES, pairs)
    return pairs / mean_t, mean_t, sample, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v, mean_t, sample, cores = cpu_reference_run(args.steps, args.warmup)
    line = {"metric": "relation_pairs_per_sec", "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": mean_t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": "cfg2: PredCLS 64 images x 40 boxes per GPU, two-pass, eval_cs (bounded sample per step)",
                       "sample": sample},
            "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


# ======================================================================================================= our arm
def run_ours(args):
    from scene_graph_commonsense_b200 import dist as hdist
    from scene_graph_commonsense_b200 import model, ops, pipeline, synthetic, tables
    rank, local, world = hdist.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pk = peaks()

    sd = synthetic.head_state_dict(seed=0, logit_gain=40.0)
    packed = model.PackedHead(sd, dev)
    del sd
    wl = WORKLOADS[args.workload]
    chunk_pairs = args.chunk_pairs or wl["chunk_pairs"]
    pipe = pipeline.RelationPipeline(packed, dev, commonsense=True, chunk_pairs=chunk_pairs, conv3_m_sub=args.conv3_m_sub,
                                     overlap=not args.no_overlap, predcls=not wl["sgdet"], chunk_policy=args.chunk_policy,
                                     conv3_block_rows=CONV3_MODES[args.conv3][0], conv3_shared=CONV3_MODES[args.conv3][1],
                                     fc1_shared=args.fc1 == "shared", conv3_block_cols=CONV3_MODES[args.conv3][2])
    samples = make_samples(rank, wl["images"], wl["boxes"], sgdet=wl["sgdet"])
    host = pipeline.host_batch_from_samples(samples, skip_mode="batch", sgdet=wl["sgdet"])
    del samples
    batch = host.to_device(dev)
    torch.cuda.synchronize()

    global_counters = torch.zeros_like(pipe.counters)

    def step_resident():
        n = pipe.step(batch)
        hdist.allreduce_counters(pipe.counters, out=global_counters)      # one int64[765] all-reduce per step (C ABI / NCCL)
        return n

    def steps_e2e(k):
        """k windows through the public streaming API: pinned host buffers in, H2D copy of every window inside the timed
        region (window i+1's copy is issued while window i computes), counters read back to the host after every window."""
        out = None
        for out in pipe.run((host for _ in range(k)), before_step=lambda p: p.reset(),
                            after_step=lambda p: hdist.allreduce_counters(p.counters, out=global_counters)):
            pass
        return out

    for _ in range(args.warmup):
        pipe.reset()
        step_resident()
    torch.cuda.synchronize()

    # ---- timed region 1: inputs resident in HBM
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    hdist.barrier()
    torch.cuda.synchronize()
    ops.PROFILE["events"].clear()
    ops.PROFILE["on"] = True
    launches0 = ops.LAUNCHES["n"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pairs_step = 0
    for _ in range(args.steps):
        pipe.reset()
        pairs_step = step_resident()
    e1.record()
    torch.cuda.synchronize()
    hdist.barrier()
    ops.PROFILE["on"] = False
    launches = ops.LAUNCHES["n"] - launches0
    clocks = sampler.stop() if rank == 0 else None
    t_dev = e0.elapsed_time(e1) / 1e3
    t_max = hdist.max_over_ranks(t_dev, dev)
    pairs_total = hdist.sum_over_ranks(pairs_step, dev)
    value = pairs_total * args.steps / t_max
    counters_final = pipe.counters.cpu().numpy().copy()                 # rank 0's own images (the recall line of the JSON)
    # block-sparse conv3_1: work-list lengths of the last step (read back AFTER the timed region)
    blocks_step = int(pipe.last_n_blocks.sum().item()) if pipe.conv3_block_rows and pipe.last_n_blocks is not None else None

    # shared-footprint fc1: K cells visited per 256-row tile of the last step (read back AFTER the timed region)
    fc1_exec_frac, fc1_cells = 1.0, None
    if pipe.fc1_shared and pipe.last_k_masks is not None:
        fc1_cells = int(np.unpackbits(pipe.last_k_masks.cpu().numpy().view(np.uint8)).sum())
        fc1_exec_frac = fc1_cells * 256.0 / (pairs_step * 64.0)
    n_box_step = wl["images"] * wl["boxes"]
    # conv2_1 halves on the box footprint: listed 8 x block_rows-pixel blocks of the last step vs the 1024 pixels of every box map
    conv2_exec_frac = 1.0
    if getattr(pipe, "conv2_sparse", False) and packed.last_conv2_blocks is not None:
        nb_dev, rows_c2, boxes_c2 = packed.last_conv2_blocks
        conv2_exec_frac = int(nb_dev.item()) * 8.0 * rows_c2 / (boxes_c2 * 1024.0)

    per_tag = {}
    for tag, a, b in ops.PROFILE["events"]:
        per_tag.setdefault(tag, []).append(a.elapsed_time(b))
    ops.PROFILE["events"].clear()

    # ---- timed region 2: end to end through the public API with HOST buffers (H2D + D2H inside)
    steps_e2e(max(1, min(args.warmup, 2)))
    hdist.barrier()
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    steps_e2e(args.steps)
    f1.record()
    torch.cuda.synchronize()
    hdist.barrier()
    t_e2e = hdist.max_over_ranks(f0.elapsed_time(f1) / 1e3, dev)
    e2e_value = pairs_total * args.steps / t_e2e

    if rank != 0:
        return
    # ---- roofline of the dominant kernel (the tensor-core GEMM with the most time in the step: conv3_1 or fc1), live
    # CUDA-event timing inside the timed region
    conv3 = per_tag.get("conv3", [])
    fc1 = per_tag.get("fc1", [])
    conv3_exec_frac = 1.0           # executed / dense-equivalent FLOPs of conv3_1 (block-sparse mode visits only listed blocks)
    if blocks_step is not None:
        conv3_exec_frac = blocks_step * pipe.conv3_block_cols * pipe.conv3_block_rows / (pairs_step * 256.0)

    def roof_of(name, times, flop_step_kernel, extra):
        if not times:
            return None
        total_ms = float(np.sum(times))
        achieved = flop_step_kernel * args.steps / (total_ms * 1e-3) / 1e12
        r = {"kernel": name, "bound": "tensor", "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
             "frac": achieved / pk["bf16_sustained"], "traffic": None, "peak_source": pk["source"] + " sustained bf16",
             "avg_launch_ms": total_ms / len(times), "launches_timed": len(times), "ms_per_step": total_ms / args.steps,
             "algorithmic_flop_per_launch": flop_step_kernel * args.steps / len(times)}
        r.update(extra)
        return r

    conv3_times = conv3 + per_tag.get("conv3_box", [])
    roof_conv3 = roof_of("tc_gemm_kernel<256,%d> conv3_1 implicit GEMM + bias/ReLU/maxpool epilogue (%s)" % (args.conv3_m_sub, args.conv3),
                         conv3_times, conv3_exec_frac * pairs_step * FLOP_PAIR_CONV3,
                         {"note": "achieved counts EXECUTED FLOPs (listed blocks only, per-pair and per-box launches); "
                                  "dense-equivalent = achieved / executed_fraction",
                          "executed_fraction": conv3_exec_frac} if blocks_step is not None else {})
    if roof_conv3 is not None:
        if blocks_step is not None:
            roof_conv3["dense_equivalent_tflops"] = roof_conv3["achieved"] / conv3_exec_frac
            tp = os.path.join(ROOT, "profiles", "ncu_summary_r01y.json")      # committed ncu --set full capture of the default path
            if os.path.exists(tp) and args.conv3 == "shared44" and pipe.fc1_shared and args.workload == "cfg2":
                for r in json.load(open(tp)).get("launches", []):
                    if r.get("what", "").startswith("conv3_1 difference epilogue, full chunk"):
                        roof_conv3["traffic"] = r.get("dram_bytes_per_launch")
                        roof_conv3["traffic_note"] = ("dram__bytes_read+write of one full-chunk launch (3.58 ms under ncu, 12 480 pairs); "
                                                      "the kernel is bound by L2->SM delivery (35.2 GB per launch through xbar->L1), not DRAM")
        else:
            tp = os.path.join(ROOT, "profiles", "conv3_dram_bytes.json")    # the committed ncu DRAM figure is the dense kernel's
            if os.path.exists(tp):
                roof_conv3["traffic"] = json.load(open(tp)).get("dram_bytes_per_launch")
    fc1_times = fc1 + per_tag.get("fc1_box", [])
    fc1_box_flop = (2 * n_box_step + 1) * FLOP_PAIR_FC1 if pipe.fc1_shared else 0        # per-box fc1 rows (dense)
    roof_fc1 = roof_of("tc_gemm_kernel<256,2> fc1 [pairs,65536] x [65536,4096] + bias/ReLU epilogue (%s)" % (args.fc1 if pipe.fc1_shared else "dense"),
                       fc1_times, fc1_exec_frac * pairs_step * FLOP_PAIR_FC1 + fc1_box_flop,
                       {"note": "achieved counts EXECUTED FLOPs: the K cells each 256-row tile visits (pair launch) + the dense per-box "
                                "rows; dense-equivalent = achieved / executed_fraction",
                        "executed_fraction": (fc1_exec_frac * pairs_step * FLOP_PAIR_FC1 + fc1_box_flop) / (pairs_step * FLOP_PAIR_FC1)}
                       if pipe.fc1_shared else {})
    if roof_fc1 is not None and pipe.fc1_shared:
        roof_fc1["dense_equivalent_tflops"] = roof_fc1["achieved"] / roof_fc1["executed_fraction"]
    roofs = [r for r in (roof_conv3, roof_fc1) if r is not None]
    roofs.sort(key=lambda r: -r["ms_per_step"])
    roof = roofs[0] if roofs else None
    roof_second = roofs[1] if len(roofs) > 1 else None
    breakdown = {t: {"launches": len(v), "ms_per_step": float(np.sum(v)) / args.steps} for t, v in sorted(per_tag.items())}
    flop_dense = wl["images"] * FLOP_IMG + wl["images"] * wl["boxes"] * FLOP_BOX + pairs_step * FLOP_PAIR
    flop_step = flop_dense - (1.0 - conv3_exec_frac) * pairs_step * FLOP_PAIR_CONV3      # FLOPs actually executed
    flop_step += fc1_box_flop - (1.0 - fc1_exec_frac) * pairs_step * FLOP_PAIR_FC1
    flop_step -= (1.0 - conv2_exec_frac) * n_box_step * FLOP_BOX
    m = pipeline.metrics_from_counters(counters_final)

    cpu = None
    if world == 1 and not args.no_cpu_baseline and args.workload == "cfg2":
        v, mean_t, sample, cores = cpu_reference_run(1, 0, budget_s=20.0)
        cpu = {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample}

    line = {
        "metric": "relation_pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_max / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": wl["text"] % (wl["images"], wl["boxes"], pairs_step),
                   "parallelism": "images sharded over %d GPU(s), one int64[765] all-reduce per step" % world,
                   "l2": "no explicit flush: each step streams >10 GB of activations/weights (>> 126 MB L2)",
                   "weights": "random init, trained-scale logits (seed 0)", "chunk_pairs": chunk_pairs,
                   "pool_gemm_overlap": not args.no_overlap, "conv3_m_sub": args.conv3_m_sub, "chunk_policy": args.chunk_policy,
                   "conv3": args.conv3, "fc1": args.fc1 if pipe.fc1_shared else "dense",
                   "conv2": "box footprint" if getattr(pipe, "conv2_sparse", False) else "dense"},
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": host.h2d_bytes * world,
                "d2h_bytes_per_step": (tables.COUNTER_SIZE * 8 + 4) * world, "ms_per_step": t_e2e / args.steps * 1e3,
                "api": "RelationPipeline.run over pinned HostBatch windows (H2D of window k+1 issued under window k's kernels; "
                       "counters read back after every window)"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "roofline_second": roof_second,
        "step_tensor_frac": flop_step / (t_max / args.steps) / 1e12 / pk["bf16_sustained"],
        "algorithmic_tflop_per_step": flop_dense / 1e12, "executed_tflop_per_step": flop_step / 1e12,
        "conv3_blocks_per_step": blocks_step, "fc1_cells_per_step": fc1_cells, "conv2_executed_fraction": conv2_exec_frac,
        "kernel_breakdown": breakdown,
        "recall": {"R@20/50/100": m["evaluator"][0], "mR@20/50/100": [float(x) for x in m["evaluator"][2]]},
        "cpu_baseline": cpu,
    }
    emit(line)


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else any library prints (NCCL banners ...) was sent to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                      # C-level stdout writers (NCCL "version" banner) must not pollute the JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chunk-pairs", type=int, default=0, help="pairs per conv3/fc1 chunk (0 = the workload's default)")
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS),
                    help="cfg2 = the configuration BASELINE.json's metric is quoted on (default); cfg3 = SGDET-shaped scaling case")
    ap.add_argument("--conv3-m-sub", type=int, default=2)
    ap.add_argument("--conv3", default="shared44", choices=sorted(CONV3_MODES),
                    help="conv3_1 kernel: dense, or block-sparse over the dilated footprint of each pair's boxes (bit-identical output)")
    ap.add_argument("--fc1", default="shared", choices=["shared", "dense"],
                    help="shared = fc1 as per-box rows + a K-cell-sparse GEMM over the cells both boxes reach (needs --conv3 shared*); "
                         "dense = fc1 over the assembled conv3_1 output")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--chunk-policy", default="waves", choices=["waves", "greedy"],
                    help="waves = image-aligned chunks sized for the fc1 GEMM's wave quantisation (default); greedy = fill to the cap")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
        import torch.distributed as tdist
        if tdist.is_available() and tdist.is_initialized():
            tdist.barrier()
            tdist.destroy_process_group()


if __name__ == "__main__":
    main()
