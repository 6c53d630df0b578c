"""Constant integer tables of the HIERCOM relation path, in the packed forms the kernels read.

Every table is either converted from a reference data file by `tools/convert_reference_data.py`
(committed under `data/`) or restates a literal of the reference, cited below.  Host-only, NumPy.

Key packing for (subject, predicate, object) triplets:
    key = (s * NUM_PRED + p) * NUM_OBJ + o        <  150 * 50 * 150 = 1 125 000
Bitmaps are uint32 words, bit `key & 31` of word `key >> 5`.
"""
import os

import numpy as np

NUM_OBJ = 150          # config.yaml:31  num_classes
NUM_PRED = 50          # config.yaml:32  num_relations
NUM_SUPER_OBJ = 17     # config.yaml:33  num_super_classes
TRIPLET_SPACE = NUM_OBJ * NUM_PRED * NUM_OBJ
BITMAP_WORDS = (TRIPLET_SPACE + 31) // 32
TOP_K = (20, 50, 100)  # evaluate.py:79

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def _load(name):
    return np.load(os.path.join(_DATA, name))


def pack_key(s, p, o):
    return (np.asarray(s, dtype=np.int64) * NUM_PRED + np.asarray(p, dtype=np.int64)) * NUM_OBJ + np.asarray(o, dtype=np.int64)


def keys_to_bitmap(keys):
    """Packed keys -> uint32[BITMAP_WORDS] membership bitmap (replaces the python dict of evaluator.py:191-192)."""
    keys = np.asarray(keys, dtype=np.int64)
    if keys.size and (keys.min() < 0 or keys.max() >= TRIPLET_SPACE):
        raise ValueError("triplet key out of range")
    bm = np.zeros(BITMAP_WORDS, dtype=np.uint32)
    np.bitwise_or.at(bm, keys >> 5, (np.uint32(1) << (keys & 31).astype(np.uint32)))
    return bm


def dict_to_keys(d):
    """Reference-format commonsense dict {(s,p,o): count} or iterable of 's_p_o' strings -> packed keys."""
    out = []
    for k in (d.keys() if hasattr(d, "keys") else d):
        if isinstance(k, str):
            s, p, o = (int(t) for t in k.split("_"))
        else:
            s, p, o = (int(t) for t in k)
        out.append((s * NUM_PRED + p) * NUM_OBJ + o)
    return np.asarray(out, dtype=np.int64)


def commonsense_aligned_keys():
    """triplets/commonsense_aligned_triplets.pt (evaluator.py:80), 20 884 keys."""
    return _load("cs_aligned_keys.npy").astype(np.int64)


def commonsense_violated_keys():
    """triplets/commonsense_violated_triplets.pt (evaluator.py:81), 1 524 keys."""
    return _load("cs_violated_keys.npy").astype(np.int64)


def zero_shot_keys():
    """datasets/vg_scene_graph_annot/zero_shot_triplets.pt (evaluator.py:39), 4 314 keys."""
    return _load("zero_shot_keys.npy").astype(np.int64)


def train_triplet_keys():
    """datasets/vg_scene_graph_annot/train_triplets.pt (evaluator.py:37); only used by an assert (:342)."""
    return _load("train_triplet_keys.npy").astype(np.int64)


def commonsense_pass_bitmap(aligned_keys=None, violated_keys=None):
    """Bitmap of triplets that SURVIVE the filter: in aligned AND NOT in violated (evaluator.py:261-266)."""
    al = keys_to_bitmap(commonsense_aligned_keys() if aligned_keys is None else aligned_keys)
    vi = keys_to_bitmap(commonsense_violated_keys() if violated_keys is None else violated_keys)
    return al & ~vi


def sub2super_table():
    """sub2super_cat_dict.pt as int8[150,4], -1 padded (evaluate.py:288,368; dataset_utils.py:576-578)."""
    return _load("sub2super.npy")


def alp2fre():
    """dataset_utils.object_class_alp2fre (dataset_utils.py:606-614) as int32[151]: DETR's alphabetical object label ->
    the frequency-ordered label of the relation path; 150 ("no object") maps to itself (evaluate.py:319-323)."""
    return _load("alp2fre.npy")


def vg_predicate_counts():
    """utils.py:258-265 get_num_each_class_reordered (VG branch); only used to draw synthetic GT predicates."""
    return _load("vg_predicate_counts.npy")


def object_synonym_matrix():
    """uint8[150,150] truth table of utils.compare_object_cat (utils.py:355-373): symmetric synonym groups plus
    three hypernym rows (vehicle / animal / food match their members in either argument order)."""
    equiv = [[1, 5, 11, 23, 38, 44, 121, 124, 148, 149], [0, 50], [92, 137]]
    hyper = {123: [14, 63, 95, 87, 123], 108: [89, 102, 67, 72, 71, 81, 96, 105, 90, 111, 108],
             60: [145, 106, 142, 144, 77, 60]}
    m = np.eye(NUM_OBJ, dtype=np.uint8)
    for g in equiv:
        for a in g:
            for b in g:
                m[a, b] = 1
    for k, lst in hyper.items():
        for b in lst:
            m[k, b] = 1
            m[b, k] = 1
    return m


# counter-vector layout shared by the kernels, the all-reduce and the metric code
NK = len(TOP_K)
EV_HITS = 0                                # [NK]
EV_HITS_PC = EV_HITS + NK                  # [NK, NUM_PRED]
EV_NGT = EV_HITS_PC + NK * NUM_PRED        # [1]
EV_NGT_PC = EV_NGT + 1                     # [NUM_PRED]
EV_BLOCK = EV_NGT_PC + NUM_PRED            # 204
EV_ZS = EV_BLOCK                           # zero-shot twin of the same block
EV_SIZE = 2 * EV_BLOCK                     # 408
T3_HITS = 0
T3_HITS_PC = T3_HITS + NK
T3_TOP1 = T3_HITS_PC + NK * NUM_PRED
T3_TOP1_PC = T3_TOP1 + NK
T3_NGT = T3_TOP1_PC + NK * NUM_PRED
T3_NGT_PC = T3_NGT + 1
T3_SIZE = T3_NGT_PC + NUM_PRED             # 357
COUNTER_SIZE = EV_SIZE + T3_SIZE           # 765
