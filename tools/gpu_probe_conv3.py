#!/usr/bin/env python
"""Probe: how much of a conv3_1 launch is NOT hidden behind the MMA main loop (epilogue, pipeline fill)?
T(K) for K = 9*c_in with c_in in {128, 256, 512} at fixed M, N: the intercept of the linear fit is the per-tile
non-overlapped cost."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scene_graph_commonsense_b200 import ops  # noqa: E402
from scene_graph_commonsense_b200._lib import ACT_NONE, EPI_BF16, EPI_POOL_BF16, GEMM_CONV3  # noqa: E402

dev = torch.device("cuda", 0)
n = int(os.environ.get("PROBE_PAIRS", "8192"))
p2 = torch.randn(n, 16, 16, 512, device=dev).to(torch.bfloat16)
bias = torch.randn(1024, device=dev)
res = []
for m_sub in (2, 1):
    for epi, ename in ((EPI_POOL_BF16, "pool"), (EPI_BF16, "bf16")):
        for c_in in (128, 256, 512):
            w = torch.randn(1024, 9 * c_in, device=dev).to(torch.bfloat16)
            out = torch.empty(n, 8, 8, 1024, dtype=torch.bfloat16, device=dev) if epi == EPI_POOL_BF16 else torch.empty(n * 256, 1024, dtype=torch.bfloat16, device=dev)

            def run():
                ops.tc_gemm(p2, w, out, n * 256, 1024, 9 * c_in, bias=bias, ldc=1024, mode=GEMM_CONV3, epilogue=epi, act=ACT_NONE, n_img=n,
                            h=16, w=16, c_total=512, c_base=0, c_in=c_in, group_m=1, m_sub=m_sub)
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            ts = []
            for _ in range(8):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); run(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            tf = 2.0 * n * 256 * 1024 * 9 * c_in / (ms * 1e-3) / 1e12
            res.append(dict(m_sub=m_sub, epilogue=ename, c_in=c_in, ms=ms, tflops=tf))
            print("m_sub=%d epi=%-4s c_in=%3d  %8.3f ms  %7.1f TFLOP/s" % (m_sub, ename, c_in, ms, tf), flush=True)
            del w, out
for m_sub in (2, 1):
    for ename in ("pool", "bf16"):
        pts = [(r["c_in"], r["ms"]) for r in res if r["m_sub"] == m_sub and r["epilogue"] == ename]
        a, b = np.polyfit([p[0] for p in pts], [p[1] for p in pts], 1)
        print("m_sub=%d epi=%-4s  T = %.4f ms * c_in/64 + %.3f ms  -> non-overlapped share at c_in=512: %.1f%%" % (
            m_sub, ename, a * 64, b, 100 * b / (a * 512 + b)))
print(json.dumps(res))
