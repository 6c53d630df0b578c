"""Convert the reference's shipped `.pt` data fixtures into packed integer tables.

Run ONCE in the build container (needs /root/reference); the outputs are committed under
`scene_graph_commonsense_b200/data/` so nothing at run time ever reads /root/reference.

Sources (reference file:line that consumes each table):
  triplets/commonsense_aligned_triplets.pt   dict {(s,p,o): count}     evaluator.py:80
  triplets/commonsense_violated_triplets.pt  dict {(s,p,o): count}     evaluator.py:81
  datasets/vg_scene_graph_annot/zero_shot_triplets.pt  list['s_p_o']   evaluator.py:39,341
  datasets/vg_scene_graph_annot/train_triplets.pt      dict{'s_p_o'}   evaluator.py:37,342
  datasets/vg_scene_graph_annot/sub2super_cat_dict.pt  dict{c:[sc..]}  evaluate.py:288,368
  utils.get_num_each_class_reordered (VG predicate counts)             utils.py:258-265
  dataset_utils.object_class_alp2fre (DETR alphabetical -> frequency object labels, 150 -> 150 = "no object")
                                                                       evaluate.py:289,319-322; dataset_utils.py:606-614

Key packing: key = (s*50 + p)*150 + o  with s,o in [0,150), p in [0,50).
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("HIERCOM_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "scene_graph_commonsense_b200", "data")


def pack(s, p, o):
    return (int(s) * 50 + int(p)) * 150 + int(o)


def main():
    os.makedirs(OUT, exist_ok=True)
    al = torch.load(os.path.join(REF, "triplets/commonsense_aligned_triplets.pt"))
    vi = torch.load(os.path.join(REF, "triplets/commonsense_violated_triplets.pt"))
    zs = torch.load(os.path.join(REF, "datasets/vg_scene_graph_annot/zero_shot_triplets.pt"))
    tr = torch.load(os.path.join(REF, "datasets/vg_scene_graph_annot/train_triplets.pt"))
    s2s = torch.load(os.path.join(REF, "datasets/vg_scene_graph_annot/sub2super_cat_dict.pt"))

    def keys_of_tuples(d):
        ks = np.array(sorted(pack(*k) for k in d.keys()), dtype=np.int32)
        assert len(np.unique(ks)) == len(d)
        return ks

    def keys_of_strings(it):
        ks = np.array([pack(*map(int, k.split("_"))) for k in it], dtype=np.int32)
        return ks

    np.save(os.path.join(OUT, "cs_aligned_keys.npy"), keys_of_tuples(al))
    np.save(os.path.join(OUT, "cs_violated_keys.npy"), keys_of_tuples(vi))
    # the zero-shot list is kept in file order (duplicates, if any, are harmless for membership)
    np.save(os.path.join(OUT, "zero_shot_keys.npy"), keys_of_strings(zs))
    np.save(os.path.join(OUT, "train_triplet_keys.npy"), np.sort(keys_of_strings(tr.keys())))

    tab = -np.ones((150, 4), dtype=np.int8)
    for c, lst in s2s.items():
        assert 1 <= len(lst) <= 4
        tab[int(c), :len(lst)] = lst
    assert (tab[:, 0] >= 0).all()
    np.save(os.path.join(OUT, "sub2super.npy"), tab)

    sys.path.insert(0, REF)
    import types
    sys.modules.setdefault("torchmetrics", types.ModuleType("torchmetrics"))
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        import utils as ref_utils
        freq = ref_utils.get_num_each_class_reordered({"dataset": {"dataset": "vg"}}).numpy().astype(np.int64)
    finally:
        os.chdir(cwd)
    np.save(os.path.join(OUT, "vg_predicate_counts.npy"), freq)
    os.chdir(REF)
    try:
        import dataset_utils as ref_du
        a2f = ref_du.object_class_alp2fre()
    finally:
        os.chdir(cwd)
    tab = np.array([a2f[i] for i in range(151)], dtype=np.int32)
    assert sorted(tab.tolist()) == list(range(151)) and tab[150] == 150
    np.save(os.path.join(OUT, "alp2fre.npy"), tab)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
