"""Drop-in twins of the reference's relation-head modules (reference model.py), backed by the sm_100a kernels.

Same class names, constructor arguments, parameter names/shapes (so `load_state_dict(torch.load('HierRelationModel_*.pth'))`
works, with or without DDP's `module.` prefix) and the same forward signature / return tuple:

  BayesianRelationClassifier.forward(h_sub, h_obj, c1, c2, s1, s2, rank, h_sub_aug=None, h_obj_aug=None)
      -> (relation_1 [B,G], relation_2 [B,P], relation_3 [B,S], super_relation [B,3], connectivity [B,1], pred [B,512], pred_aug)
  FlatRelationClassifier.forward(...) -> (relation [B,50], connectivity [B,1], pred, pred_aug)        (model.py:94-102)
  BayesianHead.forward(h) -> (relation_1, relation_2, relation_3, super_relation)                     (model.py:24-34)

The nn.Conv2d / nn.Linear children are parameter containers only; forward never calls them.  Weights are re-packed
once (bf16, GEMM-friendly layouts) into a `PackedHead`, lazily and again whenever a parameter changes.
Inference only (eval-mode semantics: dropout is the identity); there is no CPU path.
"""
import os

import torch
import torch.nn as nn

from . import ops
from ._lib import (ACT_NONE, ACT_RELU, ACT_TANH, EPI_BF16, EPI_F32, EPI_POOL_BF16, EPI_POOL_DIFF_BF16, GEMM_CONV3, GEMM_CONV3_BLOCKS,
                   GEMM_PLAIN)

K1_PAD = 320      # 257 input channels of the 1x1 convolutions, zero padded to a multiple of 64
HIDDEN = 512


def strip_module_prefix(state_dict):
    """utils.py:207-214 - checkpoints are saved from the DDP wrapper, keys carry a `module.` prefix."""
    return {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}


def supers_to_table(s_list, device):
    """list of super-class id tensors (utils.py:136-149 input format) -> int8 [n,4], -1 padded: the RAW lists.  The kernels
    apply the reference's summing rule (first entry + last entry of a 2..4-entry list, `head.cu` `super_used`); a list longer
    than 4 contributes its first entry only in the reference (its `range(1, 4)` loop never matches), so only that is kept."""
    out = torch.full((len(s_list), 4), -1, dtype=torch.int8)
    for r, s in enumerate(s_list):
        vals = [int(v) for v in (s.tolist() if hasattr(s, "tolist") else list(s))]
        vals = vals if len(vals) <= 4 else vals[:1]
        for j, v in enumerate(vals):
            out[r, j] = v
    return out.to(device)


def _env_operand_dtype():
    v = os.environ.get("HC_OPERANDS", "fp16").lower()
    if v not in ("fp16", "bf16"):
        raise RuntimeError("hiercom_b200: HC_OPERANDS must be fp16 or bf16")
    return torch.float16 if v == "fp16" else torch.bfloat16


DEFAULT_OPERAND_DTYPE = _env_operand_dtype()     # 16-bit operand format of PackedHead / the drop-in modules when none is given


class PackedHead:
    """Device-resident, kernel-ready copies of the relation-head weights.

    conv1_x [128,257,1,1]  -> w1 bf16 [256, 320] (rows 0-127 conv1_1, 128-255 conv1_2), b1 f32 [256], fill = tanh(b1) bf16
    conv2_1 [512,256,3,3]  -> w2 bf16 [512, 9*256] (k = tap*256 + c); w2s / w2o bf16 [512, 9*128] subject / object halves
    conv3_1 [1024,512,3,3] -> w3 bf16 [1024, 9*512]
    fc1 [4096, 65536]      -> columns permuted from NCHW flatten (c*64 + y*8 + x) to pixel-major ((y*8+x)*1024 + c)
    fc2 [512, 4096+L]      -> w_fc2 bf16 [512,4096]; emb f32 [L,512] = the label columns, transposed
    heads                  -> w_heads f32 [G+P+S+1+3, 512] = [fc3_1; fc3_2; fc3_3; fc4; fc5] (flat: [fc3; fc4])
    """

    def __init__(self, sd, device, flat=False, operand_dtype=None):
        """operand_dtype: the 16-bit format of every tensor-core operand and stored activation - torch.float16 (package default,
        `DEFAULT_OPERAND_DTYPE` / env HC_OPERANDS: same tcgen05 kind::f16 rate as bf16, 3 more mantissa bits = 8x smaller operand
        rounding error, which is what holds north_star's 2e-3 probability bar on trained-scale weights; stores saturate at +-65504
        and weights are range-checked here) or torch.bfloat16 (north_star's nominal "bf16 in, fp32 accumulate": full fp32 range,
        2e-3 only on low-gain weights - tests/test_gpu_parity_at_scale.py measures both)."""
        operand_dtype = DEFAULT_OPERAND_DTYPE if operand_dtype is None else operand_dtype
        sd = strip_module_prefix(sd)
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("hiercom_b200: the relation head runs on CUDA only")
        if operand_dtype not in (torch.bfloat16, torch.float16):
            raise ValueError("operand_dtype must be torch.bfloat16 or torch.float16")
        self.act_dtype = operand_dtype
        f32 = lambda t: t.detach().to(dev, torch.float32)

        def bf(t):
            if operand_dtype == torch.float16 and float(t.abs().max()) > 6.0e4:
                raise RuntimeError("hiercom_b200: a weight exceeds the fp16 range - use operand_dtype=torch.bfloat16 for this checkpoint")
            return t.to(operand_dtype).contiguous()
        c = sd["conv1_1.weight"].shape[0]
        if c != 128 or sd["fc1.weight"].shape[1] != 65536:
            raise RuntimeError("hiercom_b200: kernels are built for hidden_dim=128, feature_size=32 (config.yaml:30,35)")
        w1 = torch.zeros(2 * c, K1_PAD, device=dev)
        w1[:c, :2 * c + 1] = f32(sd["conv1_1.weight"]).view(c, -1)
        w1[c:, :2 * c + 1] = f32(sd["conv1_2.weight"]).view(c, -1)
        self.w1 = bf(w1)
        self.b1 = torch.cat((f32(sd["conv1_1.bias"]), f32(sd["conv1_2.bias"]))).contiguous()
        self.fill = bf(torch.tanh(self.b1))
        w2 = f32(sd["conv2_1.weight"])                                   # [512, 256, 3, 3]
        self.w2 = bf(w2.permute(0, 2, 3, 1).reshape(w2.shape[0], -1))
        self.w2s = bf(w2[:, :c].permute(0, 2, 3, 1).reshape(w2.shape[0], -1))
        self.w2o = bf(w2[:, c:].permute(0, 2, 3, 1).reshape(w2.shape[0], -1))
        self.b2 = f32(sd["conv2_1.bias"]).contiguous()
        w3 = f32(sd["conv3_1.weight"])
        self.w3 = bf(w3.permute(0, 2, 3, 1).reshape(w3.shape[0], -1))
        self.b3 = f32(sd["conv3_1.bias"]).contiguous()
        wf = sd["fc1.weight"].detach().to(dev)                           # [4096, 1024*64] NCHW-flatten columns
        self.w_fc1 = bf(wf.view(4096, 1024, 64).permute(0, 2, 1).reshape(4096, 65536))
        del wf
        self.b_fc1 = f32(sd["fc1.bias"]).contiguous()
        w_fc2 = f32(sd["fc2.weight"])
        self.w_fc2 = bf(w_fc2[:, :4096])
        self.emb = w_fc2[:, 4096:].t().contiguous()                      # [L, 512]
        self.has_super = self.emb.shape[0] > 300
        self.b_fc2 = f32(sd["fc2.bias"]).contiguous()
        self.flat = flat
        if flat:
            self.splits = (sd["fc3.weight"].shape[0], 0, 0)
            heads = [sd["fc3.weight"], sd["fc4.weight"]]
            biases = [sd["fc3.bias"], sd["fc4.bias"]]
        else:
            self.splits = tuple(sd["fc3_%d.weight" % k].shape[0] for k in (1, 2, 3))
            heads = [sd["fc3_1.weight"], sd["fc3_2.weight"], sd["fc3_3.weight"], sd["fc4.weight"], sd["fc5.weight"]]
            biases = [sd["fc3_1.bias"], sd["fc3_2.bias"], sd["fc3_3.bias"], sd["fc4.bias"], sd["fc5.bias"]]
        self.w_heads = torch.cat([f32(w) for w in heads]).contiguous()
        self.b_heads = torch.cat([f32(b) for b in biases]).contiguous()
        self.device = dev
        self._p3_bg = None
        self._uv_bg = None
        self.last_conv2_blocks = None

    # ---------------------------------------------------------------------------- dense stages (model.py:138-150,175)
    def conv2_halves(self, abox, m_sub=1):
        """conv2_1 split into its subject / object input halves (linear before the ReLU, model.py:143): abox [n,32,32,256] bf16
        (subject channels 0-127, object 128-255) -> U, V [n,32,32,512] bf16.  The conv2 bias rides on the object half (added in
        fp32 before the one rounding to bf16), so the pair stage is relu(maxpool(U[s] + V[o])) on packed bf16."""
        n_box, fs = abox.shape[0], abox.shape[1]
        u = torch.empty(n_box, fs, fs, 512, dtype=self.act_dtype, device=abox.device)
        v = torch.empty(n_box, fs, fs, 512, dtype=self.act_dtype, device=abox.device)
        for out, w, base, bias in ((u, self.w2s, 0, None), (v, self.w2o, 128, self.b2)):
            ops.tc_gemm(abox, w, out, n_box * fs * fs, 512, 9 * 128, bias=bias, ldc=512, mode=GEMM_CONV3, epilogue=EPI_BF16, act=ACT_NONE,
                        n_img=n_box, h=fs, w=fs, c_total=256, c_base=base, c_in=128, group_m=1, m_sub=m_sub, tag="conv2_half")
        return u, v

    def uv_background(self):
        """conv2_1 halves [1,32,32,512] bf16 of a box whose mask is empty (tanh(conv1 bias) everywhere): what U / V of any box equal
        farther than one pixel from the box.  Weights-only: computed once, through the dense kernel."""
        if self._uv_bg is None:
            fs = 32
            abox = self.fill.view(1, 1, 1, -1).expand(1, fs, fs, self.fill.numel()).contiguous()
            self._uv_bg = self.conv2_halves(abox)
        return self._uv_bg

    def conv2_halves_sparse(self, abox, boxes, m_sub=1, block_rows=4, prefill=True, poison=False):
        """`conv2_halves` on the box footprint: the implicit GEMM visits only the 8 x block_rows-pixel blocks within one pixel of each
        box (`ops.conv2_box_blocks`).  prefill=True: U / V are pre-filled with the background maps first, the result is bit-identical
        to the dense halves everywhere.  prefill=False: nothing else is written (saves 2 MiB of background per box) and the maps are
        only defined inside each box's footprint rectangle - consumers take `self.uv_footprint(boxes, block_rows)` as `fp` and read
        the background maps elsewhere."""
        n_box, fs = abox.shape[0], abox.shape[1]
        u_bg, v_bg = self.uv_background()
        u = torch.empty(n_box, fs, fs, 512, dtype=self.act_dtype, device=abox.device)
        v = torch.empty(n_box, fs, fs, 512, dtype=self.act_dtype, device=abox.device)
        if prefill:
            ops.broadcast_rows(u_bg, n_box, u)
            ops.broadcast_rows(v_bg, n_box, v)
        elif poison:                 # tests: anything read outside a footprint rectangle shows up as NaN
            u.fill_(float("nan"))
            v.fill_(float("nan"))
        blocks, n_blocks = ops.conv2_box_blocks(boxes, block_rows, fs)
        self.last_conv2_blocks = (n_blocks, block_rows, n_box)      # device count: bench.py reads it after the timed region
        for out, w, base, bias in ((u, self.w2s, 0, None), (v, self.w2o, 128, self.b2)):
            ops.tc_gemm(abox, w, out, n_box * fs * fs, 512, 9 * 128, bias=bias, ldc=512, mode=GEMM_CONV3_BLOCKS, epilogue=EPI_BF16, act=ACT_NONE,
                        n_img=n_box, h=fs, w=fs, c_total=256, c_base=base, c_in=128, group_m=1, m_sub=m_sub, tag="conv2_half", blocks=blocks,
                        n_blocks=n_blocks, block_rows=block_rows, block_cols=8)
        return u, v

    def uv_footprint(self, boxes, block_rows=4):
        """The `fp` argument of the pooling ops for U / V made by `conv2_halves_sparse(prefill=False)` on `boxes`."""
        u_bg, v_bg = self.uv_background()
        return (boxes, u_bg, v_bg, block_rows)

    def p3_background(self):
        """Pooled conv3_1 output [1,8,8,1024] bf16 of a pair whose two box masks are empty: tanh(conv1 bias) everywhere
        (train_test.py:391,398 zero the map outside the box) pushed through the same kernels as real pairs, so a real pair's
        output equals it bit for bit wherever the receptive field misses both boxes.  Weights-only: computed once."""
        if self._p3_bg is None:
            fs = 32
            u, v = self.uv_background()
            zero = torch.zeros(1, dtype=torch.int32, device=self.device)
            p2 = ops.pair_relu_pool(u, v, None, zero, zero, fs)
            p3 = torch.empty(1, 8, 8, 1024, dtype=self.act_dtype, device=self.device)
            ops.tc_gemm(p2, self.w3, p3, 256, 1024, 9 * 512, bias=self.b3, ldc=1024, mode=GEMM_CONV3, epilogue=EPI_POOL_BF16,
                        n_img=1, h=16, w=16, c_total=512, c_base=0, c_in=512, group_m=1, m_sub=2, tag="conv3_bg")
            self._p3_bg = p3
        return self._p3_bg

    def conv3_blocks(self, p2, p3, n, blocks, n_blocks, block_rows, m_sub=2, tag="conv3", block_cols=8, cta_pairs=0):
        """conv3_1+ReLU+pool on the listed blocks of p2 [>=n,16,16,512] into the PRE-FILLED p3 [>=n,8,8,1024].
        cta_pairs=1: the tcgen05 cta_group::2 pair kernel (4x4-pixel blocks; same bits)."""
        ops.tc_gemm(p2, self.w3, p3, n * 256, 1024, 9 * 512, bias=self.b3, ldc=1024, mode=GEMM_CONV3_BLOCKS, epilogue=EPI_POOL_BF16,
                    n_img=n, h=16, w=16, c_total=512, c_base=0, c_in=512, group_m=1, m_sub=m_sub, tag=tag, blocks=blocks,
                    n_blocks=n_blocks, block_rows=block_rows, block_cols=block_cols, cta_pairs=cta_pairs)
        return p3

    def pair_scratch(self, n):
        """Work map of the pair kernel's split difference epilogue: pooled conv3_1 values by local pair, [n,8,8,1024] (grown on demand,
        reused by every launch: launches on one stream are ordered)."""
        cur = getattr(self, "_pair_scratch", None)
        if cur is None or cur.shape[0] < n:
            self._pair_scratch = cur = torch.empty(n, 8, 8, 1024, dtype=self.act_dtype, device=self.w3.device)
        return cur

    def conv3_diff(self, p2, d, n, blocks, n_blocks, block_rows, sub_maps, obj_maps, pair_sub, pair_obj, pair_row, m_sub=2, tag="conv3",
                   block_cols=8, cta_pairs=0, scratch=None):
        """conv3_1+ReLU+pool on the listed blocks of p2 [>=n,16,16,512]; local pair i's cells go to row pair_row[i] of d [rows,8,8,1024]
        as the DIFFERENCE to its per-box maps, (x - sub_maps[pair_sub[i]]) - (obj_maps[pair_obj[i]] - background): the operand of
        the shared-footprint fc1 (`fc1_shared_fc2`), exactly zero wherever only one box of the pair reaches."""
        ops.tc_gemm(p2, self.w3, d, n * 256, 1024, 9 * 512, bias=self.b3, ldc=1024, mode=GEMM_CONV3_BLOCKS, epilogue=EPI_POOL_DIFF_BF16,
                    n_img=n, h=16, w=16, c_total=512, c_base=0, c_in=512, group_m=1, m_sub=m_sub, tag=tag, blocks=blocks,
                    n_blocks=n_blocks, block_rows=block_rows, block_cols=block_cols, diff_sub=sub_maps, diff_obj=obj_maps, diff_bg=self.p3_background(),
                    pair_sub=pair_sub, pair_obj=pair_obj, pair_row=pair_row, cta_pairs=cta_pairs,
                    scratch=(scratch if scratch is not None else self.pair_scratch(n)) if cta_pairs else None)
        return d

    def fc1_background(self):
        """fc1 (no bias) of the background map, f32 [4096]: weights-only, computed once."""
        if getattr(self, "_fc1_bg", None) is None:
            self._fc1_bg = self.fc1_rows(self.p3_background(), 1)[0].contiguous()
        return self._fc1_bg

    @staticmethod
    def longest_first(k_masks):
        """Visiting order of the 256-row tiles of a K-cell-sparse GEMM: most cells first (stable), int32 [tiles] on the device.
        The kernel's persistent CTAs take tiles round-robin; with tiles of 0-64 cells the slowest CTA ends well above the mean
        unless the long tiles go first (tools/fc1_order_sim.py)."""
        cells = ((k_masks.unsqueeze(1) >> torch.arange(64, device=k_masks.device, dtype=torch.int64)) & 1).sum(1)
        return torch.sort(cells, descending=True, stable=True)[1].to(torch.int32).contiguous()

    def fc1_background_cells(self):
        """fc1 (no bias) of the background map cell by cell, f32 [64, 4096]: row c = W1[:, cell c] . background[cell c] (fc1's K axis
        is packed cell-major).  Weights-only, computed once (fp32 products of the 16-bit operands, as the tensor core forms them)."""
        if getattr(self, "_fc1_bg_cells", None) is None:
            bg = self.p3_background().reshape(64, 1024).float()
            w = self.w_fc1.view(4096, 64, 1024)
            self._fc1_bg_cells = torch.stack([w[:, c, :].float() @ bg[c] for c in range(64)]).contiguous()
        return self._fc1_bg_cells

    def fc1_rows_sparse(self, maps_sorted, n, k_masks, out_rows):
        """fc1 (no bias) of n per-box maps [n,8,8,1024] whose rows are SORTED by the box's own cell rectangle -> f32 [n,4096] in
        `out_rows` order: a K-cell-sparse GEMM over the cells of each 256-row tile's mask plus, per tile, the background's contribution
        of the cells the tile skips (there every map of the tile equals the background bit for bit).  Same sums as `fc1_rows`."""
        d = maps_sorted
        out = torch.empty(n, 4096, dtype=torch.float32, device=d.device)
        pairs = int(os.environ.get("HC_FC1_BOX_PAIRS", "1"))
        order = self.longest_first(k_masks) if os.environ.get("HC_FC1_LPT", "1") != "0" else None
        ops.tc_gemm(d, self.w_fc1, out, n, 4096, 65536, lda=65536, ldc=4096, epilogue=EPI_F32, group_m=1 if pairs else 9,
                    m_sub=1 if pairs else 2, tag="fc1_box", k_masks=k_masks, k_cell=1024, out_rows=out_rows, cta_pairs=pairs, m_order=order)
        # + sum over the cells a tile does NOT visit of the background's per-cell fc1 rows (tiny: [tiles, 64] x [64, 4096], fp32)
        bits = ((k_masks.unsqueeze(1) >> torch.arange(64, device=d.device, dtype=torch.int64)) & 1).to(torch.float32)
        skipped = (1.0 - bits) @ self.fc1_background_cells()                       # [tiles, 4096]
        tile_of = torch.empty(n, dtype=torch.int64, device=d.device)
        tile_of[out_rows.long()] = torch.arange(n, device=d.device, dtype=torch.int64) // 256
        out += skipped[tile_of]
        return out

    def fc1_rows(self, maps, n):
        """fc1 WITHOUT bias / activation of n pooled maps [n,8,8,1024] bf16 -> f32 [n,4096] (the per-box terms of the shared fc1)."""
        out = torch.empty(n, 4096, dtype=torch.float32, device=maps.device)
        pairs = int(os.environ.get("HC_FC1_BOX_PAIRS", "1")) if n > 128 else 0      # cta_group::2 pairs on 256-row tiles (see fc1_shared_fc2)
        ops.tc_gemm(maps, self.w_fc1, out, n, 4096, 65536, lda=65536, ldc=4096, epilogue=EPI_F32, group_m=int(os.environ.get("HC_FC1_BOX_GROUP_M", "37")),
                    m_sub=int(os.environ.get("HC_FC1_BOX_MSUB", "1")) if pairs else (2 if n > 128 else 1), tag="fc1_box", cta_pairs=pairs)
        return out

    def fc1_shared_fc2(self, d, n, k_masks, f_sub, f_obj, row_sub, row_obj, bias_eff, out_rows, raw, group_m=None):
        """model.py:149,175 on the difference operand d [n,8,8,1024] (rows in sorted order, zero outside the cells both boxes reach):
        h1 = relu(d @ W1^T [only the cells in the tile's mask] + f_sub[row_sub] + f_obj[row_obj] + bias_eff), raw[out_rows] = h1 @ W2^T."""
        h1 = torch.empty(n, 4096, dtype=self.act_dtype, device=d.device)
        # rasterisation: a band of 9 M tiles x all 16 N tiles = 144 CTAs run together.  The rows are sorted by cell rectangle, so the
        # 9 tiles walk (nearly) the same K cells in step: each weight panel and each operand tile comes out of HBM once per band
        # and is shared through L2 (bands of 37 x 4 re-read the operand 4x and thrashed L2: 38.9 GB of DRAM reads per launch, ncu r01y)
        # HC_FC1_PAIRS=1: tcgen05 cta_group::2 pairs - a pair of CTAs owns one 256 x 256 tile (128 rows each, half of the weight tile's
        # columns each), so 74 tiles = 4.6 M tiles are in flight instead of 9.25: half as many distinct weight slabs stream through L2
        # at a time (the single-CTA launch re-reads the weights from HBM for nearly every (M tile, cell): 41 GB per launch, ncu r02i)
        pairs = int(os.environ.get("HC_FC1_PAIRS", "1"))
        if group_m is None:
            # band height of the rasterisation: with pairs and longest-first tiles one M tile at a time (its 16 N tiles run side by
            # side on 16 CTA pairs and share its operand rows): fc1 7.4 -> 6.7 ms against bands of 4 (profiles/bench_fc1band_*_r02L.json)
            group_m = int(os.environ.get("HC_FC1_GROUP_M", "1" if pairs else "9"))
        m_sub = int(os.environ.get("HC_FC1_MSUB", "1")) if pairs else 2
        order = None
        if os.environ.get("HC_FC1_LPT", "1") != "0":
            km = k_masks
            if pairs and m_sub == 2:             # two-unit pair tiles walk the union of their two masks
                if km.numel() % 2:
                    km = torch.cat((km, km.new_zeros(1)))
                km = km[0::2] | km[1::2]
            order = self.longest_first(km)
        ops.tc_gemm(d, self.w_fc1, h1, n, 4096, 65536, bias=bias_eff, lda=65536, ldc=4096, epilogue=EPI_BF16, act=ACT_RELU, group_m=group_m,
                    m_sub=m_sub, tag="fc1", k_masks=k_masks, k_cell=1024, add_a=f_sub, m_order=order,
                    add_a_rows=row_sub, add_b=f_obj,
                    add_b_rows=row_obj, cta_pairs=pairs)
        ops.tc_gemm(h1, self.w_fc2, raw, n, HIDDEN, 4096, lda=4096, ldc=HIDDEN, epilogue=EPI_F32, group_m=8, tag="fc2", out_rows=out_rows)
        return raw

    def conv3_fc(self, p2, m_sub=2, raw=None, n=None, blocks=None, n_blocks=None, block_rows=0, p3=None, block_cols=8):
        """conv3_1+ReLU+pool -> fc1+ReLU -> fc2 (raw, fp32) on pooled conv2 activations p2 [>=n,16,16,512] bf16.
        With a work list (`ops.conv3_active_blocks`) conv3_1 visits only the listed blocks of each pair; the rest of its output is
        the background (`p3`, if given, is a buffer the caller has ALREADY pre-filled with it)."""
        n = p2.shape[0] if n is None else n
        dev = p2.device
        if blocks is None:
            p3 = torch.empty(n, 8, 8, 1024, dtype=self.act_dtype, device=dev)
            ops.tc_gemm(p2, self.w3, p3, n * 256, 1024, 9 * 512, bias=self.b3, ldc=1024, mode=GEMM_CONV3, epilogue=EPI_POOL_BF16,
                        n_img=n, h=16, w=16, c_total=512, c_base=0, c_in=512, group_m=1, m_sub=m_sub, tag="conv3")
        else:
            if p3 is None:
                p3 = ops.broadcast_rows(self.p3_background(), n, torch.empty(n, 8, 8, 1024, dtype=self.act_dtype, device=dev))
            self.conv3_blocks(p2, p3, n, blocks, n_blocks, block_rows, m_sub=m_sub, block_cols=block_cols)
        h1 = torch.empty(n, 4096, dtype=self.act_dtype, device=dev)
        ops.tc_gemm(p3, self.w_fc1, h1, n, 4096, 65536, bias=self.b_fc1, lda=65536, ldc=4096, epilogue=EPI_BF16, act=ACT_RELU,
                    group_m=37, m_sub=2 if n > 128 else 1, tag="fc1")
        if raw is None:
            raw = torch.empty(n, HIDDEN, dtype=torch.float32, device=dev)
        ops.tc_gemm(h1, self.w_fc2, raw, n, HIDDEN, 4096, lda=4096, ldc=HIDDEN, epilogue=EPI_F32, group_m=8, tag="fc2")
        return raw

    def legacy_hidden(self, h_sub, h_obj):
        """model.py:138-150 on pre-masked [bs,257,32,32] inputs -> fc2 pre-activation of the feature part [bs,512]."""
        bs = h_sub.shape[0]
        dev = h_sub.device
        a = torch.empty(bs, 32, 32, 256, dtype=self.act_dtype, device=dev)
        for role, h in enumerate((h_sub, h_obj)):
            x = ops.pack_pixels(h.to(torch.float32), None, K1_PAD, dtype=self.act_dtype)
            ops.tc_gemm(x, self.w1[128 * role:128 * (role + 1)], a, bs * 1024, 128, K1_PAD, bias=self.b1[128 * role:128 * (role + 1)],
                        lda=K1_PAD, ldc=256, c_off=128 * role, epilogue=EPI_BF16, act=ACT_TANH)
        p2 = torch.empty(bs, 16, 16, 512, dtype=self.act_dtype, device=dev)
        ops.tc_gemm(a, self.w2, p2, bs * 1024, 512, 9 * 256, bias=self.b2, ldc=512, mode=GEMM_CONV3, epilogue=EPI_POOL_BF16,
                    n_img=bs, h=32, w=32, c_total=256, c_base=0, c_in=256, group_m=1, m_sub=2)
        return self.conv3_fc(p2)

    def heads(self, raw, row_sub, row_obj, box_cat, box_super, temps=(1.0, 1.0, 1.0), want_pred=False):
        return ops.hier_head(raw, self.b_fc2, self.emb, row_sub, row_obj, box_cat, box_super if self.has_super else None,
                             self.w_heads, self.b_heads, self.splits, flat=self.flat, temps=temps, want_pred=want_pred)


class _HeadBase(nn.Module):
    def _build_trunk(self, args, input_dim, feature_size, num_classes, num_super_classes):
        self.input_dim = input_dim
        self.num_classes = num_classes
        self.num_super_classes = num_super_classes
        self.conv1_1 = nn.Conv2d(2 * input_dim + 1, input_dim, kernel_size=1, stride=1, padding=0)
        self.conv1_2 = nn.Conv2d(2 * input_dim + 1, input_dim, kernel_size=1, stride=1, padding=0)
        self.conv2_1 = nn.Conv2d(2 * input_dim, 4 * input_dim, kernel_size=3, stride=1, padding=1)
        self.conv3_1 = nn.Conv2d(4 * input_dim, 8 * input_dim, kernel_size=3, stride=1, padding=1)
        self.dropout1 = nn.Dropout(p=0.5)
        self.dropout2 = nn.Dropout(p=0.5)
        self.maxpool = nn.MaxPool2d(kernel_size=2, stride=2)
        self.fc1 = nn.Linear(8 * input_dim * (feature_size // 4) ** 2, 4096)
        if args['dataset']['dataset'] == 'vg':
            self.fc2 = nn.Linear(4096 + 2 * (num_classes + num_super_classes), 512)
        else:
            self.fc2 = nn.Linear(4096 + 2 * num_classes, 512)
        self._packed = None
        self._packed_versions = None

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._packed = None
        return super().load_state_dict(strip_module_prefix(state_dict), strict=strict, **kw)

    def packed(self):
        """Kernel-ready weights; rebuilt when any parameter tensor has been modified or moved."""
        versions = tuple((p.data_ptr(), p._version) for p in self.parameters()) + (getattr(self, "operand_dtype", None),)
        if self._packed is None or versions != self._packed_versions:
            dev = next(self.parameters()).device
            self._packed = PackedHead(self.state_dict(), dev, flat=self._flat, operand_dtype=getattr(self, "operand_dtype", None))
            self._packed_versions = versions
        return self._packed

    def _legacy_forward(self, h_sub, h_obj, c1, c2, s1, s2):
        pk = self.packed()
        dev = h_sub.device
        bs = h_sub.shape[0]
        raw = pk.legacy_hidden(h_sub, h_obj)
        box_cat = torch.stack((c1.to(dev, torch.int32), c2.to(dev, torch.int32)), dim=1).reshape(-1).contiguous()
        box_super = None
        if s1 is not None and pk.has_super:
            box_super = torch.stack((supers_to_table(s1, dev), supers_to_table(s2, dev)), dim=1).reshape(-1, 4).contiguous()
        elif pk.has_super:
            box_super = torch.full((2 * bs, 4), -1, dtype=torch.int8, device=dev)
        row_sub = torch.arange(0, 2 * bs, 2, dtype=torch.int32, device=dev)
        row_obj = row_sub + 1
        temps = (float(getattr(self, "T1", 1)), float(getattr(self, "T2", 1)), float(getattr(self, "T3", 1)))
        return pk.heads(raw, row_sub, row_obj, box_cat, box_super, temps=temps, want_pred=True)


class BayesianRelationClassifier(_HeadBase):
    """reference model.py:105-186."""
    _flat = False

    def __init__(self, args, input_dim=128, feature_size=32, num_classes=150, num_super_classes=17, num_geometric=15,
                 num_possessive=11, num_semantic=24, T1=1, T2=1, T3=1):
        super().__init__()
        self._build_trunk(args, input_dim, feature_size, num_classes, num_super_classes)
        self.fc3_1 = nn.Linear(512, num_geometric)
        self.fc3_2 = nn.Linear(512, num_possessive)
        self.fc3_3 = nn.Linear(512, num_semantic)
        self.fc4 = nn.Linear(512, 1)
        self.fc5 = nn.Linear(512, 3)
        self.T1, self.T2, self.T3 = T1, T2, T3

    @torch.no_grad()
    def forward(self, h_sub, h_obj, c1, c2, s1, s2, rank, h_sub_aug=None, h_obj_aug=None):
        relation, sup, conn, _, pred = self._legacy_forward(h_sub, h_obj, c1, c2, s1, s2)
        pred_aug = None
        if h_sub_aug is not None:
            pred_aug = self._legacy_forward(h_sub_aug, h_obj_aug, c1, c2, s1, s2)[4]
        g, p = self.fc3_1.out_features, self.fc3_2.out_features
        return relation[:, :g], relation[:, g:g + p], relation[:, g + p:], sup, conn.view(-1, 1), pred, pred_aug


class FlatRelationClassifier(_HeadBase):
    """reference model.py:37-102 (flat ablation)."""
    _flat = True

    def __init__(self, args, input_dim=128, output_dim=50, feature_size=32, num_classes=150, num_super_classes=17):
        super().__init__()
        self._build_trunk(args, input_dim, feature_size, num_classes, num_super_classes)
        self.fc3 = nn.Linear(512, output_dim)
        self.fc4 = nn.Linear(512, 1)

    @torch.no_grad()
    def forward(self, h_sub, h_obj, c1, c2, s1, s2, rank, h_sub_aug=None, h_obj_aug=None, one_hot=True):
        relation, _, conn, _, pred = self._legacy_forward(h_sub, h_obj, c1, c2, s1, s2)
        pred_aug = None
        if h_sub_aug is not None:
            pred_aug = self._legacy_forward(h_sub_aug, h_obj_aug, c1, c2, s1, s2)[4]
        return relation, conn.view(-1, 1), pred, pred_aug


class BayesianHead(nn.Module):
    """reference model.py:9-34 - the hierarchical head alone (512 -> G/P/S/3)."""

    def __init__(self, input_dim=512, num_geometric=15, num_possessive=11, num_semantic=24, T1=1, T2=1, T3=1):
        super().__init__()
        self.fc3_1 = nn.Linear(input_dim, num_geometric)
        self.fc3_2 = nn.Linear(input_dim, num_possessive)
        self.fc3_3 = nn.Linear(input_dim, num_semantic)
        self.fc5 = nn.Linear(input_dim, 3)
        self.T1, self.T2, self.T3 = T1, T2, T3

    @torch.no_grad()
    def forward(self, h):
        if h.shape[1] != HIDDEN:
            raise RuntimeError("hiercom_b200: BayesianHead kernels are built for input_dim=512")
        dev = h.device
        z = lambda *s: torch.zeros(*s, device=dev)
        w = torch.cat((self.fc3_1.weight, self.fc3_2.weight, self.fc3_3.weight, z(1, HIDDEN), self.fc5.weight)).float().contiguous()
        b = torch.cat((self.fc3_1.bias, self.fc3_2.bias, self.fc3_3.bias, z(1), self.fc5.bias)).float().contiguous()
        splits = (self.fc3_1.out_features, self.fc3_2.out_features, self.fc3_3.out_features)
        rel, sup, _, _, _ = ops.hier_head(h.float().contiguous(), None, None, None, None, None, None, w, b, splits,
                                          temps=(float(self.T1), float(self.T2), float(self.T3)))
        g, p = splits[0], splits[1]
        return rel[:, :g], rel[:, g:g + p], rel[:, g + p:], sup
