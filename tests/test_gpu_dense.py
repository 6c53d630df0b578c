"""GPU parity of the tcgen05 dense kernel and the streaming prep kernels against plain torch fp32 references of the
same ops on the same (bf16-rounded) inputs.  Tolerances: fp32 outputs 2e-3 relative to the output scale, bf16 outputs
one bf16 ulp (2^-8 relative) plus the same accumulation slack."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    from scene_graph_commonsense_b200 import ops
    return ops


def _rand(shape, seed, scale=1.0, dev="cuda"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev)


def _close(got, ref, rel, what):
    got, ref = got.float(), ref.float()
    scale = ref.abs().max().item() + 1e-6
    err = (got - ref).abs().max().item()
    assert err <= rel * scale, "%s: max abs err %.4g vs scale %.4g (limit %.3g)" % (what, err, scale, rel * scale)


@pytest.mark.parametrize("m,n,k,m_sub", [(128, 256, 64, 1), (128, 128, 128, 1), (300, 256, 320, 1), (1000, 512, 1024, 2),
                                         (77, 128, 4096, 1), (513, 4096, 256, 2), (2048, 256, 65536 // 8, 2)])
def test_plain_gemm_f32(m, n, k, m_sub):
    ops = _ops()
    a = _rand((m, k), 1).to(torch.bfloat16)
    b = _rand((n, k), 2).to(torch.bfloat16)
    bias = _rand((n,), 3)
    out = torch.full((m, n), float("nan"), device="cuda")
    ops.tc_gemm(a, b, out, m, n, k, bias=bias, lda=k, ldc=n, epilogue=ops.EPI_F32, m_sub=m_sub, group_m=3)
    ref = a.float() @ b.float().t() + bias
    _close(out, ref, 2e-3, "gemm f32")


@pytest.mark.parametrize("act", ["none", "relu", "tanh"])
def test_plain_gemm_bf16_epilogue_with_offset(act):
    ops = _ops()
    m, n, k, ldc, off = 640, 128, 320, 256, 128
    a = _rand((m, k), 4, 0.2).to(torch.bfloat16)
    b = _rand((n, k), 5, 0.2).to(torch.bfloat16)
    bias = _rand((n,), 6, 0.1)
    out = torch.zeros((m, ldc), dtype=torch.bfloat16, device="cuda")
    code = {"none": ops.ACT_NONE, "relu": ops.ACT_RELU, "tanh": ops.ACT_TANH}[act]
    ops.tc_gemm(a, b, out, m, n, k, bias=bias, lda=k, ldc=ldc, c_off=off, epilogue=ops.EPI_BF16, act=code)
    ref = a.float() @ b.float().t() + bias
    ref = {"none": ref, "relu": torch.relu(ref), "tanh": torch.tanh(ref)}[act]
    _close(out[:, off:], ref, 6e-3, "gemm bf16 " + act)
    assert (out[:, :off] == 0).all(), "columns outside [c_off, c_off+N) must be untouched"


def _conv_ref(x_nhwc, w_packed, c_in, bias=None):
    n_out = w_packed.shape[0]
    w = w_packed.float().view(n_out, 3, 3, c_in).permute(0, 3, 1, 2)
    y = F.conv2d(x_nhwc.float().permute(0, 3, 1, 2), w, bias, padding=1)
    return y


@pytest.mark.parametrize("n_img,hw,c_total,c_base,c_in,n_out,m_sub", [(3, 16, 64, 0, 64, 128, 1), (2, 32, 256, 128, 128, 512, 2),
                                                                      (5, 16, 512, 0, 512, 256, 2), (2, 32, 256, 0, 256, 256, 1)])
def test_implicit_conv_bf16(n_img, hw, c_total, c_base, c_in, n_out, m_sub):
    ops = _ops()
    x = _rand((n_img, hw, hw, c_total), 7, 0.5).to(torch.bfloat16)
    w = _rand((n_out, 9 * c_in), 8, 0.05).to(torch.bfloat16)
    out = torch.zeros((n_img, hw, hw, n_out), dtype=torch.bfloat16, device="cuda")
    ops.tc_gemm(x, w, out, n_img * hw * hw, n_out, 9 * c_in, ldc=n_out, mode=ops.GEMM_CONV3, epilogue=ops.EPI_BF16,
                n_img=n_img, h=hw, w=hw, c_total=c_total, c_base=c_base, c_in=c_in, m_sub=m_sub)
    ref = _conv_ref(x[..., c_base:c_base + c_in], w, c_in).permute(0, 2, 3, 1)
    _close(out, ref, 6e-3, "conv bf16")


@pytest.mark.parametrize("n_img,hw,c_in,n_out,m_sub", [(4, 16, 512, 1024, 2), (3, 32, 256, 512, 2), (3, 16, 128, 256, 1)])
def test_implicit_conv_relu_pool(n_img, hw, c_in, n_out, m_sub):
    ops = _ops()
    x = _rand((n_img, hw, hw, c_in), 9, 0.5).to(torch.bfloat16)
    w = _rand((n_out, 9 * c_in), 10, 0.05).to(torch.bfloat16)
    bias = _rand((n_out,), 11, 0.2)
    out = torch.zeros((n_img, hw // 2, hw // 2, n_out), dtype=torch.bfloat16, device="cuda")
    ops.tc_gemm(x, w, out, n_img * hw * hw, n_out, 9 * c_in, bias=bias, ldc=n_out, mode=ops.GEMM_CONV3, epilogue=ops.EPI_POOL_BF16,
                n_img=n_img, h=hw, w=hw, c_total=c_in, c_base=0, c_in=c_in, m_sub=m_sub)
    ref = F.max_pool2d(torch.relu(_conv_ref(x, w, c_in, bias)), 2, 2).permute(0, 2, 3, 1)
    _close(out, ref, 6e-3, "conv relu pool")


def test_pack_pixels_and_box_select_and_pair_pool():
    ops = _ops()
    feat, depth = _rand((3, 256, 32, 32), 12), _rand((3, 1, 32, 32), 13)
    x = ops.pack_pixels(feat, depth, 320)
    ref = torch.cat((feat, depth), 1).permute(0, 2, 3, 1).reshape(-1, 257)
    assert torch.equal(x[:, :257], ref.to(torch.bfloat16)) and (x[:, 257:] == 0).all()
    t = _rand((3, 1024, 256), 14).to(torch.bfloat16)
    boxes = torch.tensor([[2, 9, 3, 30], [0, 32, 0, 32], [5, 5, 1, 8], [31, 40, -4, 7]], dtype=torch.int32, device="cuda")
    box_img = torch.tensor([0, 2, 1, 1], dtype=torch.int32, device="cuda")
    fill = _rand((256,), 15).to(torch.bfloat16)
    got = ops.box_select(t, boxes, box_img, fill)
    for j in range(4):
        m = torch.zeros(32, 32, dtype=torch.bool)
        b = boxes[j].tolist()
        m[b[2]:b[3], b[0]:b[1]] = True                     # python slice semantics == reference mask
        m = m.cuda()
        ref = torch.where(m[..., None], t[box_img[j]].view(32, 32, 256), fill.view(1, 1, 256))
        assert torch.equal(got[j], ref)
    u, v = _rand((4, 32, 32, 512), 16).to(torch.bfloat16), _rand((4, 32, 32, 512), 17).to(torch.bfloat16)
    bias = _rand((512,), 18)
    ps = torch.tensor([0, 3, 2, 1, 1], dtype=torch.int32, device="cuda")
    po = torch.tensor([1, 0, 2, 3, 0], dtype=torch.int32, device="cuda")
    got = ops.pair_relu_pool(u, v, bias, ps, po)
    s = torch.relu(u[ps.long()].float() + (v[po.long()].float() + bias))
    ref = F.max_pool2d(s.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(got, ref)
    # bias == None: V already carries the conv2 bias; packed bf16x2 add/max == rounding the fp32 sum, bit for bit
    got = ops.pair_relu_pool(u, v, None, ps, po)
    s = torch.relu((u[ps.long()].float() + v[po.long()].float()).to(torch.bfloat16).float())
    ref = F.max_pool2d(s.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(got, ref)
    # wide dynamic range (exponent gaps > 16 bits, exact cancellation, signed zeros) keeps the identity
    u2 = (u.float() * torch.exp2(torch.randint(-30, 30, u.shape, device="cuda").float())).to(torch.bfloat16)
    v2 = torch.where(torch.rand(v.shape, device="cuda") < 0.1, -u2[[1, 0, 3, 2]], v)
    got = ops.pair_relu_pool(u2, v2, None, ps, po)
    s = torch.relu((u2[ps.long()].float() + v2[po.long()].float()).to(torch.bfloat16).float())
    ref = F.max_pool2d(s.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(got.float(), ref.float())            # compares values: -0.0 == +0.0


def test_shape_errors_are_reported_not_swallowed():
    ops = _ops()
    a = torch.zeros(128, 100, dtype=torch.bfloat16, device="cuda")
    b = torch.zeros(128, 100, dtype=torch.bfloat16, device="cuda")
    out = torch.zeros(128, 128, device="cuda")
    with pytest.raises(RuntimeError, match="HC_E_SHAPE"):
        ops.tc_gemm(a, b, out, 128, 128, 100, lda=100, epilogue=ops.EPI_F32)
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        ops.tc_gemm(a.cpu(), b, out, 128, 128, 64, lda=64, epilogue=ops.EPI_F32)


@pytest.mark.parametrize("with_bias", [True, False])
@pytest.mark.parametrize("mode", ["batch", "per_image"])
def test_tiled_pair_pool_equals_gather_kernel_on_enumerated_pairs(mode, with_bias):
    """The outer-sum tiled pooling kernel (LUT-addressed, image-aligned chunks, second stream) writes exactly what the
    generic per-pair gather kernel writes, including when the skip rule removes pairs."""
    from scene_graph_commonsense_b200 import pipeline, synthetic
    ops = _ops()
    samples = synthetic.make_batch([500, 501, 502, 503], [7, 12, 2, 9], with_maps=False)
    b = pipeline.batch_from_samples(samples, "cuda", skip_mode=mode, with_maps=False)
    pipe = pipeline.RelationPipeline(None, "cuda", commonsense=False, chunk_pairs=60)
    pairs = pipe.enumerate_pairs(b)
    assert pairs["n"] > 0
    n_box = b.boxes.shape[0]
    u, v = _rand((n_box, 32, 32, 512), 31).to(torch.bfloat16), _rand((n_box, 32, 32, 512), 32).to(torch.bfloat16)
    bias = _rand((512,), 33) if with_bias else None        # None: the packed-bf16 kernels (bias folded into V upstream)
    ref = ops.pair_relu_pool(u, v, bias, pairs["sub"], pairs["obj"])
    lut = ops.pair_lut_build(pairs["sub"], pairs["obj"], pairs["img"], b.box_offsets, n_box, 12)
    chunks = pipe._image_chunks(pairs["offsets_host"])
    assert len(chunks) >= 2 and sum(c[3] for c in chunks) == pairs["n"]
    for img0, n_img, base, cnt in chunks:
        out = torch.full((cnt + 3, 16, 16, 512), 7.0, dtype=torch.bfloat16, device="cuda")
        ops.pair_relu_pool_tiled(u, v, bias, b.box_offsets, lut, img0, n_img, base, cnt, out=out)
        assert torch.equal(out[:cnt], ref[base:base + cnt])
        assert (out[cnt:] == 7.0).all()
