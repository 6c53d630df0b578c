"""ctypes binding of libhiercom_b200.so (the C ABI in include/hiercom_b200.h).

The library is built in-tree by `scene_graph_commonsense_b200.build`; it is the ONLY implementation of the hot path:
if it is missing, or the device is not sm_100, calls raise - there is no CPU or PyTorch fallback.
"""
import ctypes as C
import os

import torch

from . import build as _build

HC_OK = 0
ABI_VERSION = 5        # include/hiercom_b200.h HC_ABI_VERSION
ERRORS = {-1: "HC_E_SHAPE", -2: "HC_E_ALIGN", -3: "HC_E_ARCH", -4: "HC_E_CUDA", -5: "HC_E_NULL"}

GEMM_PLAIN, GEMM_CONV3, GEMM_CONV3_BLOCKS = 0, 1, 2
EPI_BF16, EPI_F32, EPI_POOL_BF16, EPI_SPLIT3_BF16, EPI_POOL_DIFF_BF16 = 0, 1, 2, 3, 4
ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 2


class GemmDesc(C.Structure):
    _fields_ = [("a", C.c_void_p), ("b", C.c_void_p), ("bias", C.c_void_p), ("out", C.c_void_p),
                ("m", C.c_int64), ("n", C.c_int64), ("k", C.c_int64),
                ("lda", C.c_int64), ("ldc", C.c_int64), ("c_off", C.c_int64),
                ("mode", C.c_int32), ("epilogue", C.c_int32), ("act", C.c_int32),
                ("n_img", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("c_total", C.c_int32), ("c_base", C.c_int32), ("c_in", C.c_int32),
                ("group_m", C.c_int32), ("m_sub", C.c_int32), ("mul", C.c_void_p), ("ld_mul", C.c_int64),
                ("blocks", C.c_void_p), ("n_blocks", C.c_void_p), ("block_rows", C.c_int32), ("block_cols", C.c_int32), ("cta_pairs", C.c_int32),
                ("k_masks", C.c_void_p), ("k_cell", C.c_int64),
                ("add_a", C.c_void_p), ("add_a_rows", C.c_void_p), ("add_b", C.c_void_p), ("add_b_rows", C.c_void_p), ("ld_add", C.c_int64),
                ("out_rows", C.c_void_p),
                ("diff_sub", C.c_void_p), ("diff_obj", C.c_void_p), ("diff_bg", C.c_void_p),
                ("pair_sub", C.c_void_p), ("pair_obj", C.c_void_p), ("pair_row", C.c_void_p), ("scratch", C.c_void_p),
                ("m_order", C.c_void_p), ("operand_f16", C.c_int32)]


class UVFootprint(C.Structure):
    """hc_uv_footprint: U / V hold a box's conv2_1 values only inside its footprint rectangle; the background maps apply elsewhere."""
    _fields_ = [("boxes", C.c_void_p), ("u_bg", C.c_void_p), ("v_bg", C.c_void_p), ("block_rows", C.c_int32)]


class RelationWorkspace(C.Structure):
    """hc_relation_workspace: device bytes of every buffer of the batched relation path for one window."""
    _fields_ = [(n, C.c_int64) for n in ("pixels_packed", "conv1_out", "box_select", "conv2_halves", "pooled_conv2", "work_lists",
                                         "box_maps", "box_fc1_rows", "fc1_operand", "row_maps", "pooled_conv3", "fc1_out", "fc2_raw",
                                         "head_out", "candidates", "total")]


_P, _I32, _I64, _F, _D = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double

SIGNATURES = {
    "hc_last_error": (C.c_char_p, []),
    "hc_abi_version": (C.c_int, []),
    "hc_device_check": (C.c_int, []),
    "hc_cs_bitmap_build": (C.c_int, [_P, _I64, _P, _I64, _P]),
    "hc_pairs_enumerate": (C.c_int, [_P, _P, _I32, _P, _I32, _I32, _P, _P, _P, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "hc_tc_gemm": (C.c_int, [C.POINTER(GemmDesc), _P]),
    "hc_conv3_active_blocks": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "hc_conv3_shared_blocks": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "hc_p3_assemble": (C.c_int, [_P, _P, _P, _P, _P, _P, _I32, _I32, _P, _P]),
    "hc_broadcast_rows": (C.c_int, [_P, _I64, _I64, _P, _P]),
    "hc_conv2_box_blocks": (C.c_int, [_P, _I32, _I32, _I32, _P, _P, _P]),
    "hc_pair_cell_keys": (C.c_int, [_P, _P, _P, _I32, _I32, _P, _P]),
    "hc_tile_cell_masks": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _P, _P]),
    "hc_cells_zero": (C.c_int, [_P, _I32, _I64, _I32, _I64, _P, _P]),
    "hc_pack_pixels": (C.c_int, [_P, _I32, _P, _I32, _I32, _I32, _I32, _P, _I32, _P]),
    "hc_box_select": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _P, _P, _P]),
    "hc_pair_relu_pool": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _I32, _P, C.POINTER(UVFootprint), _P, _I32, _P]),
    "hc_pair_lut_build": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _P, _P]),
    "hc_pair_relu_pool_tiled": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P, C.POINTER(UVFootprint), _P, _I32, _P]),
    "hc_pair_cover_masks": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _I32, _P, _P]),
    "hc_hier_head": (C.c_int, [_P, _I64, _I32, _I32, _P, _P, _I32, _I32, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32,
                               _F, _F, _F, _P, _P, _P, _P, _P, _P, _P]),
    "hc_box_label_embed": (C.c_int, [_P, _I32, _I32, _P, _P, _I32, _I32, _P, _P]),
    "hc_candidates": (C.c_int, [_P, _I64, _I32, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                _I32, _P]),
    "hc_topk_match": (C.c_int, [_P, _I32, _P, _P, _P, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _P, _I32,
                                _D, _I32, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P]),
    "hc_topk_select": (C.c_int, [_P, _I32, _P, _I32, _P, _P]),
    "hc_connectivity_stats": (C.c_int, [_P, _P, _P, _I32, _P, _P]),
    "hc_sgb_pair_gather": (C.c_int, [_P, _P, _I32, _I32, _I32, _P, _P]),
    "hc_split_bf16x3": (C.c_int, [_P, _I64, _I64, _I32, _P, _P]),
    "hc_sgb_hier_softmax": (C.c_int, [_P, _I64, _I32, _I32, _I32, _I32, _P, _I32, _P, _P, _P, _P, _P]),
    "hc_sgb_candidates": (C.c_int, [_P, _I32, _I32, _I32, _P, _P, _P, _P, _P, _I32, _P, _P, _P, _P]),
    "hc_sgb_rank_match": (C.c_int, [_P, _P, _P, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _F, _I32, _I32, _I32, _I32,
                                    _P, _P, _P, _P, _P, _P]),
    "hc_detr_proposals": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P, _I32, _D, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "hc_proposals_pack": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _P, _I32, _P, _P, _P, _P, _P, _P, _P]),
    "hc_match_object_categories": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _P, _P, _P, _P, _P, _P]),
    "hc_match_object_categories_fill": (C.c_int, [_P, _P, _P, _P, _P, _I32, _P, _P, _P, _P, _I32, _P, _P, _P, _P, _P, _P, _P]),
    "hc_targets_flat": (C.c_int, [_P, _P, _P, _P, _I32, _P, _P, _P, _P, _P, _P]),
    "hc_hier_loss": (C.c_int, [_P, _I64, _P, _P, _I32, _I32, _I32, _I32, _I32, _F, _F, _F, _P, _P, _P, _I32, _P, _P, _P, _P, _P, _P,
                               _P, _F, _F, _F, _F, _F, _P, _P, _P, _I32, _P]),
    "hc_hier_head_bwd": (C.c_int, [_P, _I32, _P, _I64, _I32, _I32, _P, _P, _P, _P, _P, _P, _I32, _P]),
    "hc_counts_allreduce": (C.c_int, [_P, _P, _I64, _P]),
    "hc_nccl_unique_id": (C.c_int, [_P]),
    "hc_nccl_comm_create": (C.c_int, [_P, _I32, _I32, C.POINTER(C.c_void_p)]),
    "hc_nccl_comm_destroy": (C.c_int, [_P]),
    "hc_pairs_enumerate_workspace_bytes": (C.c_int64, [_I64, _I32, _I32, _I32, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "hc_conv3_blocks_capacity": (C.c_int64, [_I64, _I32, _I32]),
    "hc_conv2_box_blocks_capacity": (C.c_int64, [_I64, _I32]),
    "hc_relation_workspace_bytes": (C.c_int, [_I32, _I64, _I64, _I64, _I32, C.POINTER(RelationWorkspace)]),
}

_lib = None


def library_path():
    """The in-tree shared object; HC_LIB names another build of the same ABI (A/B runs of two kernel versions on one box)."""
    return os.environ.get("HC_LIB") or _build.LIB


def load(build_if_missing=False):
    """Load (once) and return the ctypes library.  Raises if the shared object is absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        if build_if_missing:
            _build.build()
        else:
            raise RuntimeError("hiercom_b200: %s is missing - run `python -m scene_graph_commonsense_b200.build` "
                               "(there is no CPU fallback)" % path)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here means the .so does not export the ABI
        fn.restype = res
        fn.argtypes = args
    if lib.hc_abi_version() != ABI_VERSION:
        raise RuntimeError("hiercom_b200: ABI version mismatch")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != HC_OK:
        msg = load().hc_last_error().decode("utf-8", "replace")
        raise RuntimeError("hiercom_b200 %s failed: %s (%s)" % (what, ERRORS.get(rc, rc), msg))


def ptr(t):
    """Device (or host) pointer of a tensor / numpy array, None -> NULL."""
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return t.data_ptr() if t.numel() else None
    return t.ctypes.data


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("hiercom_b200: all operands must be CUDA tensors - this path has no CPU implementation")
