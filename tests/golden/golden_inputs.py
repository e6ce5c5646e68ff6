"""Seeded inputs shared by make_golden.py (which feeds them to the reference's own functions)
and by the tests (which feed the same inputs to the oracle and to the CUDA path)."""
from __future__ import annotations

import os

import numpy as np
import torch


def similarity_inputs() -> dict:
    """name -> (features [N, D] fp32 unit rows, targets [N] int, label, ref_feature [D]).

    Mirrors what code/search_image.py:main feeds get_similarity: a unit-normalised gallery
    (:157), integer class targets (:178), and a query that is a mean of unit vectors and hence
    NOT unit norm (:310-318, :387)."""
    out = {}
    for name, (n, d, n_cls, seed) in {"small": (257, 64, 5, 11), "clip512": (2000, 512, 10, 1),
                                      "taiyi768": (1001, 768, 6, 7)}.items():
        g = torch.Generator().manual_seed(seed)
        centers = torch.randn(n_cls, d, generator=g)
        targets = torch.randint(0, n_cls, (n,), generator=g)
        feats = centers[targets] * 0.6 + torch.randn(n, d, generator=g)
        feats = feats / feats.norm(dim=-1, keepdim=True)
        label = 2
        members = feats[targets == label][:10]
        text = centers[label] / centers[label].norm()
        ref_feature = (members.mean(dim=0) + text) / 2      # search_image.py:387, un-normalised
        out[name] = (feats.contiguous(), targets.numpy(), label, ref_feature.contiguous())
    return out


def topk_inputs() -> dict:
    """name -> (logits [N, C] fp32, target [N] int64); includes rows with exact ties."""
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(64, 6, generator=g)
    target = torch.randint(0, 6, (64,), generator=g)
    tied = torch.randn(32, 6, generator=g).round()   # many exact ties per row
    tied_t = torch.randint(0, 6, (32,), generator=g)
    return {"random": (logits, target), "tied": (tied, tied_t)}


def dedup_image_set(root: str):
    """Two folders of small generated images; some files of the delete folder are pixel-identical
    to reference images (one saved in a different lossless format), one reference image is
    duplicated inside the reference folder, one file is not an image."""
    from PIL import Image
    rng = np.random.default_rng(3)
    ref_dir = os.path.join(root, "reference")
    del_dir = os.path.join(root, "delete")
    os.makedirs(os.path.join(ref_dir, "sub"))
    os.makedirs(os.path.join(del_dir, "nested", "deep"))
    imgs = [rng.integers(0, 256, size=(16 + i, 20, 3), dtype=np.uint8) for i in range(8)]

    def save(arr, path):
        Image.fromarray(arr, "RGB").save(path)

    save(imgs[0], os.path.join(ref_dir, "a.png"))
    save(imgs[1], os.path.join(ref_dir, "b.bmp"))
    save(imgs[2], os.path.join(ref_dir, "sub", "c.png"))
    save(imgs[2], os.path.join(ref_dir, "sub", "c_again.png"))   # same hash twice: last one wins (:52)
    save(imgs[3], os.path.join(ref_dir, "d.tiff"))
    save(imgs[0], os.path.join(del_dir, "a_copy.bmp"))            # dup of a.png, other container
    save(imgs[2], os.path.join(del_dir, "nested", "c_copy.png"))
    save(imgs[4], os.path.join(del_dir, "unique1.png"))
    save(imgs[5], os.path.join(del_dir, "nested", "deep", "unique2.bmp"))
    save(imgs[3], os.path.join(del_dir, "nested", "deep", "d_copy.png"))
    save(imgs[6], os.path.join(del_dir, "skipped.webp"))          # extension not in the list (:26)
    with open(os.path.join(del_dir, "broken.png"), "wb") as f:    # unreadable: hash None -> kept (:67-69)
        f.write(b"not an image")
    return ref_dir, del_dir


def lab3_inputs():
    """(similarities list-of-dicts, thresholds, positive_class, negative_class) as CLIP/lab3.py:108-122 builds
    them: float(fp32 cosine), string labels, a 1001-point linspace; a third class is present and must be ignored."""
    g = torch.Generator().manual_seed(21)
    n = 3000
    labels = ["dog", "others", "cat"]
    lab = torch.randint(0, 3, (n,), generator=g)
    sim = (0.22 + 0.05 * torch.randn(n, generator=g) + 0.06 * (lab == 0)).to(torch.float32)
    similarities = [{"similarity": float(s), "true_label": labels[int(l)], "file_path": f"{i}.jpg"}
                    for i, (s, l) in enumerate(zip(sim, lab))]
    thresholds = np.linspace(0.0, 0.5, 1001)
    return similarities, thresholds, "dog", "others"


def union_inputs():
    """Inputs for CLIP/union_dataset.py process_images (:247-260) and calc_combined_metrics (:133-231):
    two "models" (EN D=512, CN D=768) over overlapping sets of files.  Returns a dict with, per model,
    un-normalised image features [N, D], unit text features {class: [1, D]}, labels, paths (some label
    "error", some basenames shared between the models, one basename repeated inside a model), plus
    class lists and per-pair thresholds."""
    g = torch.Generator().manual_seed(33)
    en_pos, en_neg = ["dog", "cat"], ["wolf", "lynx"]
    cn_pos, cn_neg = ["gou", "mao"], ["lang", "shelizi"]

    def model(d, pos, neg, n, prefix, shift):
        classes = pos + neg + ["other"]
        lab_idx = torch.randint(0, len(classes), (n,), generator=g)
        centers = torch.randn(len(classes), d, generator=g)
        feats = centers[lab_idx] * 0.12 + torch.randn(n, d, generator=g)          # weak signal: both outcomes on both sides of the thresholds
        feats = feats * (0.5 + torch.rand(n, 1, generator=g))             # NOT unit norm
        labels = [classes[int(i)] for i in lab_idx]
        for i in range(0, n, 17):
            labels[i] = "error"                                              # unreadable image (lab3.py:116)
        paths = [f"/data/{prefix}/{labels[i]}/img_{(i + shift) % (n + 5)}.jpg" for i in range(n)]
        paths[5] = paths[3]                                                  # a basename twice inside one model
        text = {}
        for c, cls in enumerate(pos):
            t = centers[c] + 0.3 * torch.randn(d, generator=g)
            text[cls] = (t / t.norm()).reshape(1, d)
        return feats.contiguous(), text, labels, paths

    en = model(512, en_pos, en_neg, 150, "en", 0)
    cn = model(768, cn_pos, cn_neg, 140, "cn", 3)
    # make the CN label names line up with the EN file basenames class-wise: same index -> same kind
    return dict(en=en, cn=cn, en_pos=en_pos, en_neg=en_neg, cn_pos=cn_pos, cn_neg=cn_neg,
                en_threshs=[0.10, 0.09], cn_threshs=[0.095, 0.11])


def overlap_grid_inputs() -> dict:
    """name -> (pos_res, neg_res) fp32 score sets for code/main_custom.py find_thresholds (:46-91), whose grid
    spans the OVERLAP of the two sets: two overlapping pairs, one overlapping over less than 0.1 (zero
    grid points) and one separable pair (negative width: numpy raises inside the reference)."""
    rng = np.random.default_rng(17)
    return {
        "wide": ((25 + 5 * rng.standard_normal(400)).astype(np.float32), (15 + 5 * rng.standard_normal(3000)).astype(np.float32)),
        "skewed": ((22 + 2 * rng.standard_normal(50)).astype(np.float32), (20 + 3 * rng.standard_normal(20000)).astype(np.float32)),
        "narrow": (np.array([0.50, 0.53], dtype=np.float32), np.array([0.48, 0.52], dtype=np.float32)),
        "separable": (np.array([30.0, 31.0, 35.0], dtype=np.float32), np.array([10.0, 12.0, 29.0], dtype=np.float32)),
    }


# ---- query builders (code/search_image.py:119-140, :185-232, :295-318) ---------------------------------------
QB_CLASSES = ["alpha", "beta", "gamma"]


def query_builder_setup(root: str):
    """A tiny on-disk dataset + stand-in encoder for the reference's query builders, which load sample
    images from `dataset_path/class_name/`, `preprocess` them, stack and call `clip_model.encode_image`.
    Returns (dataset_path, clip_model, preprocess, class_embeddings [3, 48], class_to_idx, samples) where
    samples = {class: [file names]}; class "alpha" has two visual modes of very different sizes (14 + 6
    images: k-means majority branch), class "beta" two balanced modes (10 + 10: global-mean branch)."""
    from PIL import Image
    rng = np.random.default_rng(9)
    g = torch.Generator().manual_seed(9)
    dataset_path = os.path.join(root, "data", "search")
    samples = {}
    modes = {"alpha": (14, 6), "beta": (10, 10), "gamma": (7, 0)}
    for cls, (n_a, n_b) in modes.items():
        os.makedirs(os.path.join(dataset_path, cls))
        base_a = rng.integers(0, 256, size=(8, 8, 3))
        base_b = rng.integers(0, 256, size=(8, 8, 3))
        names = []
        for i in range(n_a + n_b):
            base = base_a if i < n_a else base_b
            arr = np.clip(base + rng.integers(-12, 13, size=(8, 8, 3)), 0, 255).astype(np.uint8)
            name = f"img_{i:02d}.png"
            Image.fromarray(arr, "RGB").save(os.path.join(dataset_path, cls, name))
            names.append(name)
        samples[cls] = names
    proj = torch.randn(192, 48, generator=g) / 14.0

    def preprocess(img):
        return torch.from_numpy(np.asarray(img, dtype=np.float32).copy()).permute(2, 0, 1) / 255.0 - 0.5

    class Tower:
        def encode_image(self, images):
            return images.flatten(1) @ proj

    class_embeddings = torch.randn(3, 48, generator=g) * 2.0          # NOT unit norm (:131 normalises)
    class_to_idx = {c: i for i, c in enumerate(QB_CLASSES)}
    return dataset_path, Tower(), preprocess, class_embeddings, class_to_idx, samples


def cosine_inputs():
    """(image_features [B, 512] un-normalised, text_features [1, 512], logit_scale) for
    `logit_scale * F.cosine_similarity(image_features, text_features)` (code/merge_dataset.py:275-278)."""
    g = torch.Generator().manual_seed(44)
    x = torch.randn(300, 512, generator=g) * (0.2 + 3 * torch.rand(300, 1, generator=g))
    t = torch.randn(1, 512, generator=g)
    x[:40] += 1.5 * t
    return x.contiguous(), t.contiguous(), torch.tensor(100.0)


def greedy_inputs(root: str):
    """Files of distinct sizes plus a symmetric 'similar' relation for the same-folder greedy loop
    (tool/find_repeated_in_same_folder.py:56-106) run with stub hash functions: the hash of a file is its
    id, two hashes compare similar iff the pair is in the relation.  Returns (folder, ids {path: id},
    similar {frozenset({a, b})}, unreadable {id})."""
    from PIL import Image
    rng = np.random.default_rng(21)
    folder = os.path.join(root, "same_folder")
    os.makedirs(os.path.join(folder, "sub"))
    n = 40
    order = rng.permutation(n)
    ids = {}
    for rank, i in enumerate(order):                                 # distinct file sizes, unrelated to the id order
        side = 4 + rank
        arr = rng.integers(0, 256, size=(side, side, 3), dtype=np.uint8)
        path = os.path.join(folder, "sub" if i % 5 == 0 else "", f"f{i:02d}.bmp")
        Image.fromarray(arr, "RGB").save(path)
        ids[path] = int(i)
    similar = set()
    for _ in range(45):
        a, b = (int(v) for v in rng.choice(n, size=2, replace=False))
        similar.add(frozenset((a, b)))
    return folder, ids, similar, {3, 29}
