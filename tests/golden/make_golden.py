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
from golden_inputs import (dedup_image_set, lab3_inputs, overlap_grid_inputs, similarity_inputs, topk_inputs,  # noqa: E402
                           union_inputs)


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
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
