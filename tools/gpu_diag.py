"""Blind-debug aid for the tcgen05 kernel: structured probes, each in its own subprocess with a timeout, so one gpurun
round trip tells WHICH part of the data path is wrong (descriptor layout, K advance, accumulation, conv shift, pooling)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PROBES = ["id_k64", "id_k128", "kslice", "rand_small", "rand_bn256", "rand_ms2", "rand_big", "bf16_epi", "conv_onehot", "conv_rand",
          "conv_pool", "fc1_shape"]


def report(name, got, ref, extra=None):
    import torch
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    bad = err > (1e-2 * ref.abs().max().item() + 1e-3)
    out = {"probe": name, "max_err": err.max().item(), "ref_scale": ref.abs().max().item(), "n_bad": int(bad.sum()), "numel": got.numel(),
           "nan": int(torch.isnan(got).sum())}
    if bad.any():
        idx = bad.nonzero()[:8].tolist()
        out["first_bad"] = [(i, float(got[tuple(i)]), float(ref[tuple(i)])) for i in idx]
        rows_bad = bad.reshape(bad.shape[0], -1).any(1).nonzero().flatten().tolist()
        out["bad_rows_head"] = rows_bad[:16]
        out["n_bad_rows"] = len(rows_bad)
        if bad.dim() == 2:
            cols_bad = bad.any(0).nonzero().flatten().tolist()
            out["bad_cols_head"] = cols_bad[:16]
            out["n_bad_cols"] = len(cols_bad)
    if extra:
        out.update(extra)
    print(json.dumps(out), flush=True)


def run(name):
    import torch
    import torch.nn.functional as F
    from scene_graph_commonsense_b200 import ops
    dev = "cuda"
    g = torch.Generator().manual_seed(0)
    ri = lambda *s: torch.randint(-3, 4, s, generator=g).float()
    if name in ("id_k64", "id_k128"):
        k = 64 if name == "id_k64" else 128
        a = ri(128, k).to(dev).to(torch.bfloat16)
        b = torch.zeros(128, k)
        for i in range(k):
            b[i, i] = 1
        b = b.to(dev).to(torch.bfloat16)
        out = torch.full((128, 128), -77.0, device=dev)
        ops.tc_gemm(a, b, out, 128, 128, k, lda=k, epilogue=ops.EPI_F32)
        torch.cuda.synchronize()
        report(name, out, a.float() @ b.float().t())
    elif name == "kslice":
        a = torch.zeros(128, 64); a[:, 16:32] = ri(128, 16)
        b = ri(128, 64)
        a, b = a.to(dev).to(torch.bfloat16), b.to(dev).to(torch.bfloat16)
        out = torch.full((128, 128), -77.0, device=dev)
        ops.tc_gemm(a, b, out, 128, 128, 64, lda=64, epilogue=ops.EPI_F32)
        torch.cuda.synchronize()
        report(name, out, a.float() @ b.float().t())
    elif name in ("rand_small", "rand_bn256", "rand_ms2", "rand_big"):
        m, n, k, ms = {"rand_small": (200, 128, 192, 1), "rand_bn256": (300, 256, 320, 1), "rand_ms2": (700, 512, 256, 2),
                       "rand_big": (4096, 1024, 4608, 2)}[name]
        a = ri(m, k).to(dev).to(torch.bfloat16)
        b = ri(n, k).to(dev).to(torch.bfloat16)
        bias = ri(n).to(dev)
        out = torch.full((m, n), -77.0, device=dev)
        ops.tc_gemm(a, b, out, m, n, k, bias=bias, lda=k, epilogue=ops.EPI_F32, m_sub=ms, group_m=2)
        torch.cuda.synchronize()
        report(name, out, a.float() @ b.float().t() + bias)
    elif name == "bf16_epi":
        m, n, k = 256, 256, 128
        a = ri(m, k).to(dev).to(torch.bfloat16)
        b = ri(n, k).to(dev).to(torch.bfloat16)
        bias = ri(n).to(dev)
        out = torch.zeros((m, n), dtype=torch.bfloat16, device=dev)
        ops.tc_gemm(a, b, out, m, n, k, bias=bias, lda=k, epilogue=ops.EPI_BF16, act=ops.ACT_RELU)
        torch.cuda.synchronize()
        report(name, out, torch.relu(a.float() @ b.float().t() + bias))
    elif name == "conv_onehot":
        x = torch.zeros(1, 16, 16, 64); x[0, 5, 9, 3] = 1.0
        w = torch.zeros(128, 9 * 64)
        for tap in range(9):
            w[tap, tap * 64 + 3] = float(tap + 1)           # output channel `tap` sees only tap `tap`
        x, w = x.to(dev).to(torch.bfloat16), w.to(dev).to(torch.bfloat16)
        out = torch.zeros(1, 16, 16, 128, dtype=torch.bfloat16, device=dev)
        ops.tc_gemm(x, w, out, 256, 128, 576, ldc=128, mode=ops.GEMM_CONV3, epilogue=ops.EPI_BF16, n_img=1, h=16, w=16, c_total=64, c_base=0, c_in=64)
        torch.cuda.synchronize()
        wt = w.float().view(128, 3, 3, 64).permute(0, 3, 1, 2)
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt, padding=1).permute(0, 2, 3, 1)
        nz = out.float().nonzero().tolist()
        report(name, out, ref, {"got_nonzero": [(i, float(out[tuple(i)])) for i in nz[:12]],
                                "ref_nonzero": [(i, float(ref[tuple(i)])) for i in ref.nonzero().tolist()[:12]]})
    elif name in ("conv_rand", "conv_pool"):
        n_img, hw, c, n = 3, 16, 128, 256
        x = ri(n_img, hw, hw, c).to(dev).to(torch.bfloat16)
        w = (ri(n, 9 * c) * 0.25).to(dev).to(torch.bfloat16)
        bias = ri(n).to(dev)
        wt = w.float().view(n, 3, 3, c).permute(0, 3, 1, 2)
        y = F.conv2d(x.float().permute(0, 3, 1, 2), wt, bias, padding=1)
        if name == "conv_rand":
            out = torch.zeros(n_img, hw, hw, n, dtype=torch.bfloat16, device=dev)
            ops.tc_gemm(x, w, out, n_img * hw * hw, n, 9 * c, bias=bias, ldc=n, mode=ops.GEMM_CONV3, epilogue=ops.EPI_BF16, n_img=n_img, h=hw, w=hw,
                        c_total=c, c_base=0, c_in=c, m_sub=2)
            torch.cuda.synchronize()
            report(name, out, y.permute(0, 2, 3, 1))
        else:
            out = torch.zeros(n_img, hw // 2, hw // 2, n, dtype=torch.bfloat16, device=dev)
            ops.tc_gemm(x, w, out, n_img * hw * hw, n, 9 * c, bias=bias, ldc=n, mode=ops.GEMM_CONV3, epilogue=ops.EPI_POOL_BF16, n_img=n_img, h=hw,
                        w=hw, c_total=c, c_base=0, c_in=c, m_sub=2)
            torch.cuda.synchronize()
            report(name, out, F.max_pool2d(torch.relu(y), 2, 2).permute(0, 2, 3, 1))
    elif name == "fc1_shape":
        m, n, k = 512, 4096, 65536
        a = (torch.randn(m, k, generator=g) * 0.1).to(dev).to(torch.bfloat16)
        b = (torch.randn(n, k // 8, generator=g) * 0.1).repeat(1, 8).to(dev).to(torch.bfloat16)
        out = torch.zeros(m, n, dtype=torch.bfloat16, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ops.tc_gemm(a, b, out, m, n, k, lda=k, epilogue=ops.EPI_BF16, m_sub=2, group_m=37)
        e0.record()
        ops.tc_gemm(a, b, out, m, n, k, lda=k, epilogue=ops.EPI_BF16, m_sub=2, group_m=37)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        report(name, out, a.float() @ b.float().t(), {"ms": ms, "tflops": 2.0 * m * n * k / ms / 1e9})


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1])
        sys.exit(0)
    for p in PROBES:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), p], capture_output=True, text=True, timeout=180)
            print(r.stdout.strip() or json.dumps({"probe": p, "no_output": True}))
            if r.returncode != 0:
                print(json.dumps({"probe": p, "exit": r.returncode, "stderr": r.stderr[-800:]}))
        except subprocess.TimeoutExpired:
            print(json.dumps({"probe": p, "timeout": True}), flush=True)
