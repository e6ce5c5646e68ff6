"""Generate golden vectors by executing the REFERENCE'S OWN function bodies in this container.

Run once here (the reference checkout is not present on the GPU box):
    python tests/golden/make_golden.py [/root/reference]

Nothing from the reference is copied into the repository: the functions are pulled out of the
reference files with `ast` at generation time, compiled, and run on seeded synthetic inputs; only
their inputs' seeds and their OUTPUTS are stored (tests/golden/*.npz, *.json).

  code/search_image.py   get_similarity (:105-117), eval_threshold (:39-56), find_thresholds (:58-103)
  code/utils.py          cls_acc (:15-39)  -> the only topk in the repo (:17)
  code/main_custom.py    eval_threshold (:27-45), find_thresholds (:46-91: the overlap-range grid),
                         get_similarity (:93-105: column slice of a similarity matrix)
  CLIP/lab3.py           evaluate_thresholds (:39-65)
  CLIP/union_dataset.py  process_images (:247-260, fed a stand-in "model" whose image tower returns
                         the seeded feature batches), calc_combined_metrics (:133-231)
  tool/find_repeated.py  calculate_image_hash (:6-19), get_all_images (:21-33),
                         find_and_remove_duplicate_images (:35-71)
  code/search_image.py   get_image_text_features (:119-140), get_cluster_features (:185-232), outlier_filter
                         (:295-318) on a tiny on-disk dataset with a stand-in encoder (a fixed projection)
  code/merge_dataset.py  clip_en_predict (:259-284: logit_scale * F.cosine_similarity, thresholded)
  tool/find_repeated_in_same_folder.py
                         get_all_images (:24-36), find_and_remove_duplicate_images (:56-106) with stub
                         hash functions (imagehash is absent): pins the greedy first-keeper walk

The one patch applied: torch.Tensor.cuda is made a no-op (this container has no GPU and
search_image.py:107 hard-codes `.cuda()`), so the reference's arithmetic runs as its torch CPU
fp32 path -- exactly the oracle BASELINE.json names.
"""
from __future__ import annotations

import ast
import json
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from golden_inputs import (cosine_inputs, dedup_image_set, greedy_inputs, lab3_inputs, overlap_grid_inputs,  # noqa: E402
                           query_builder_setup, similarity_inputs, topk_inputs, union_inputs)


def extract_functions(path: Path, names: list[str], namespace: dict) -> dict:
    tree = ast.parse(path.read_text(encoding="utf-8"))
    picked = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    missing = set(names) - {n.name for n in picked}
    if missing:
        raise RuntimeError(f"{path}: functions not found: {missing}")
    mod = ast.Module(body=picked, type_ignores=[])
    exec(compile(mod, str(path), "exec"), namespace)
    return namespace


def main(ref_root: str) -> None:
    ref = Path(ref_root)
    torch.Tensor.cuda = lambda self, *a, **k: self  # the single patch (see module docstring)

    # ---- search_image.py ------------------------------------------------------------------
    ns = {"np": np, "torch": torch}
    extract_functions(ref / "code" / "search_image.py",
                      ["get_similarity", "eval_threshold", "find_thresholds"], ns)
    out = {}
    for name, (features, targets, label, ref_feature) in similarity_inputs().items():
        pos, neg = ns["get_similarity"](features, targets, label, ref_feature)
        out[f"{name}_pos"] = pos
        out[f"{name}_neg"] = neg
        probe = np.linspace(min(pos.min(), neg.min()), max(pos.max(), neg.max()), 7)
        with np.errstate(all="ignore"):
            ev = np.array([ns["eval_threshold"](pos, neg, t) for t in probe], dtype=np.float64)
            best_f1 = ns["find_thresholds"](pos, neg, "golden", verbose=False)
        out[f"{name}_probe"] = probe
        out[f"{name}_eval"] = ev
        out[f"{name}_best_f1"] = np.float64(best_f1)
    np.savez(HERE / "search_image_golden.npz", **out)

    # ---- main_custom.py: the other threshold grid (:46-50) and its get_similarity (:93-105) -------------
    ns6 = {"np": np, "torch": torch}
    extract_functions(ref / "code" / "main_custom.py", ["eval_threshold", "find_thresholds", "get_similarity"], ns6)
    out6 = {}
    for name, (pos, neg) in overlap_grid_inputs().items():
        try:
            with np.errstate(all="ignore"):
                out6[f"{name}_best_f1"] = np.float64(ns6["find_thresholds"](pos, neg, "golden", verbose=False))
            out6[f"{name}_raises"] = np.array("")
        except Exception as e:                                   # separable sets: negative sample count (:49-50)
            out6[f"{name}_best_f1"] = np.float64("nan")
            out6[f"{name}_raises"] = np.array(type(e).__name__)
    feats, targets, label, _ = similarity_inputs()["small"]
    sim = 100.0 * feats @ feats[:7].t()                          # a [N, classes] similarity matrix as main_custom builds it
    p6, n6 = ns6["get_similarity"](sim, torch.from_numpy(targets), label)
    out6["sliced_pos"], out6["sliced_neg"] = p6, n6
    np.savez(HERE / "main_custom_golden.npz", **out6)

    # ---- utils.py cls_acc / topk ----------------------------------------------------------
    ns2 = {"torch": torch}
    extract_functions(ref / "code" / "utils.py", ["cls_acc"], ns2)
    out2 = {}
    for name, (logits, target) in topk_inputs().items():
        for k in (1, 3):
            out2[f"{name}_acc_k{k}"] = np.float64(ns2["cls_acc"](logits, target, topk=k))
            v, i = logits.topk(k, 1, True, True)   # the expression at utils.py:17
            out2[f"{name}_topk{k}_values"] = v.numpy()
            out2[f"{name}_topk{k}_indices"] = i.numpy()
    np.savez(HERE / "utils_topk_golden.npz", **out2)

    # ---- CLIP/lab3.py evaluate_thresholds (:39-65) -----------------------------------------------
    ns4 = {}
    extract_functions(ref / "CLIP" / "lab3.py", ["evaluate_thresholds"], ns4)
    sims, thresholds, pos_cls, neg_cls = lab3_inputs()
    res = ns4["evaluate_thresholds"](sims, thresholds, pos_cls, neg_cls)
    keys = ["threshold", "precision", "recall", "f1", "TP", "FP", "TN", "FN"]
    np.savez(HERE / "lab3_golden.npz", **{k: np.array([r[k] for r in res], dtype=np.float64) for k in keys})

    # ---- CLIP/union_dataset.py process_images (:247-260) + calc_combined_metrics (:133-231) ------------
    import contextlib, io
    ns5 = {"os": os, "torch": torch, "tqdm": lambda it, **kw: it}
    extract_functions(ref / "CLIP" / "union_dataset.py", ["process_images", "calc_combined_metrics"], ns5)
    u = union_inputs()

    class FeatureTower:                      # stands in for the CLIP models: the "images" ARE the features
        def encode_image(self, imgs):        # model_type == "en" (:251)
            return imgs
        def get_image_features(self, pixel_values):   # model_type == "cn" (:252)
            return pixel_values

    def loader(feats, labels, paths, batch=64):       # what the DataLoader yields: (imgs, labels, paths)
        for lo in range(0, feats.shape[0], batch):
            yield feats[lo:lo + batch], labels[lo:lo + batch], paths[lo:lo + batch]

    sims = {}
    for key, pos in (("en", u["en_pos"]), ("cn", u["cn_pos"])):
        feats, text, labels, paths = u[key]
        sims[key] = ns5["process_images"](loader(feats, labels, paths), FeatureTower(), text, pos, "cpu", model_type=key)
    with contextlib.redirect_stdout(io.StringIO()):   # the reference prints debug lines
        combined = ns5["calc_combined_metrics"](sims["en"], sims["cn"], u["en_threshs"], u["cn_threshs"],
                                                u["en_pos"], u["en_neg"], u["cn_pos"], u["cn_neg"])
    (HERE / "union_golden.json").write_text(json.dumps(
        {"sims": {k: {cls: [[it["similarity"], it["true_label"], it["file_path"]] for it in v] for cls, v in d.items()}
                  for k, d in sims.items()},
         "combined": combined}, indent=0, sort_keys=True))

    # ---- tool/find_repeated.py --------------------------------------------------------------
    import hashlib
    from collections import defaultdict
    from PIL import Image
    ns3 = {"os": os, "hashlib": hashlib, "Image": Image, "defaultdict": defaultdict}
    extract_functions(ref / "tool" / "find_repeated.py",
                      ["calculate_image_hash", "get_all_images", "find_and_remove_duplicate_images"], ns3)
    with tempfile.TemporaryDirectory() as tmp:
        ref_dir, del_dir = dedup_image_set(tmp)
        hashes = {os.path.relpath(p, tmp): ns3["calculate_image_hash"](p)
                  for p in sorted(ns3["get_all_images"](tmp))}
        deleted, kept, n_ref, n_del = ns3["find_and_remove_duplicate_images"](ref_dir, del_dir)
        rel = lambda p: os.path.relpath(p, tmp)
        golden = {
            "hashes": hashes,
            "deleted": sorted([rel(a), rel(b)] for a, b in deleted),
            "kept": sorted(rel(p) for p in kept),
            "n_ref": n_ref, "n_del": n_del,
            "remaining_in_delete_folder": sorted(rel(p) for p in ns3["get_all_images"](del_dir)),
        }
    (HERE / "find_repeated_golden.json").write_text(json.dumps(golden, indent=1, sort_keys=True))

    # ---- code/search_image.py query builders ------------------------------------------------------------
    from collections import Counter
    from sklearn.cluster import KMeans
    with tempfile.TemporaryDirectory() as tmp:
        dataset_path, tower, preprocess, class_embeddings, class_to_idx, samples = query_builder_setup(tmp)
        ns7 = {"os": os, "np": np, "torch": torch, "Image": Image, "KMeans": KMeans, "Counter": Counter,
               "dataset_path": dataset_path, "class_to_idx": class_to_idx}
        extract_functions(ref / "code" / "search_image.py",
                          ["get_image_text_features", "get_cluster_features", "outlier_filter"], ns7)
        out7 = {}
        for cls in samples:
            with torch.no_grad():
                img_f, img_txt_f = ns7["get_image_text_features"](tower, preprocess, class_embeddings.clone(),
                                                                  samples[cls], cls)
            out7[f"{cls}_image_features"] = img_f.numpy()
            out7[f"{cls}_image_text_features"] = img_txt_f.numpy()
            out7[f"{cls}_outlier_filter"] = ns7["outlier_filter"](tower, preprocess, samples[cls], cls).numpy()
        for cls, shots in (("alpha", 5), ("beta", 8)):
            np.random.seed(0)                                   # the reference's KMeans is unseeded (:199)
            with contextlib.redirect_stdout(io.StringIO()):     # it prints the cluster sizes (:208)
                out7[f"{cls}_cluster_features"] = ns7["get_cluster_features"](tower, preprocess, samples[cls], shots, cls).numpy()
        np.savez(HERE / "query_builders_golden.npz", **out7)

    # ---- code/merge_dataset.py: logit_scale * F.cosine_similarity (:275-278) ------------------------------
    ns8 = {"torch": torch, "device": "cpu"}
    extract_functions(ref / "code" / "merge_dataset.py", ["clip_en_predict"], ns8)
    x, t, logit_scale = cosine_inputs()

    class CosModel:
        def __init__(self):
            self.logit_scale = torch.log(logit_scale)
        def eval(self):
            return self
        def encode_image(self, images):
            return images.clone()               # the function normalises in place (:270)
        def encode_text(self, text_inputs):
            return text_inputs.clone()

    labels = torch.zeros(x.shape[0], dtype=torch.int64)
    loader = [(x[lo:lo + 64], labels[lo:lo + 64]) for lo in range(0, x.shape[0], 64)]
    out8 = {}
    with np.errstate(all="ignore"):
        for thr in (-3.0, 0.0, 2.5, 70.0):
            preds, _ = ns8["clip_en_predict"](CosModel(), loader, t, thr)
            out8[f"preds_{thr:g}"] = np.array([int(p) for p in preds], dtype=np.int64)
    # the similarity the function thresholds (its lines :269-278 on the same inputs)
    xi = x / x.norm(dim=-1, keepdim=True)
    ti = t / t.norm(dim=-1, keepdim=True)
    out8["similarity"] = (CosModel().logit_scale.exp() * torch.nn.functional.cosine_similarity(xi, ti)).numpy()
    np.savez(HERE / "merge_dataset_golden.npz", **out8)

    # ---- tool/find_repeated_in_same_folder.py: the greedy first-keeper loop (:56-106) ----------------------
    with tempfile.TemporaryDirectory() as tmp:
        folder, ids, similar, unreadable = greedy_inputs(tmp)
        ns9 = {"os": os}
        extract_functions(ref / "tool" / "find_repeated_in_same_folder.py",
                          ["get_all_images", "find_and_remove_duplicate_images"], ns9)
        ns9["calculate_perceptual_hash"] = lambda p, hash_size=8: None if ids[p] in unreadable else (ids[p],)
        ns9["compare_hashes"] = lambda h1, h2, threshold=5: frozenset((h1[0], h2[0])) in similar
        with contextlib.redirect_stdout(io.StringIO()):
            deleted, refs, total = ns9["find_and_remove_duplicate_images"](folder, 5)
        golden9 = {"deleted": [[ids[a], ids[b]] for a, b in deleted], "representatives": [ids[p] for p in refs],
                   "total": total, "remaining": sorted(ids[p] for p in ns9["get_all_images"](folder))}
    (HERE / "same_folder_golden.json").write_text(json.dumps(golden9, indent=1, sort_keys=True))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
