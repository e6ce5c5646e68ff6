"""Parity of the CUDA self-join and the file-level dedup wrappers with the oracle."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from golden_inputs import dedup_image_set

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,d,frac", [(2, 8, 0.5), (300, 64, 0.1), (129, 16, 0.2), (5000, 512, 0.02), (20_000, 128, 0.01)])
def test_pairs_bit_exact(mm, oracle, n, d, frac):
    x, planted = oracle.synthetic_dedup(n, d, dup_frac=frac, seed=n)
    want = oracle.dedup_pairs(x, 0.95)
    got = mm.find_duplicate_pairs(x, 0.95)
    assert got.dtype == torch.int64 and torch.equal(got, want)
    assert [tuple(p) for p in got.tolist()] == planted


def test_no_pairs_and_all_pairs(mm, oracle):
    x = oracle.synthetic_gallery(500, 64, seed=1, dtype=torch.float32)
    assert mm.find_duplicate_pairs(x, 0.95).shape == (0, 2)
    got = mm.find_duplicate_pairs(x, -2.0)             # every pair qualifies: exercises buffer regrowth
    assert got.shape[0] == 500 * 499 // 2 and torch.equal(got, oracle.dedup_pairs(x, -2.0))


def test_row_ranges_partition_the_join(mm, oracle):
    from mmrs_b200.dedup import selfjoin_raw, sort_pairs, _device_f32
    from mmrs_b200.sharded import triangle_bounds
    x, planted = oracle.synthetic_dedup(10_000, 64, dup_frac=0.03, seed=3)
    xd = _device_f32(x)
    parts = [selfjoin_raw(xd, 0.95, lo, hi) for lo, hi in triangle_bounds(10_000, 4)]
    got = sort_pairs(torch.cat(parts), 10_000).cpu()
    assert [tuple(p) for p in got.tolist()] == planted


def test_same_folder_wrapper(mm, oracle, tmp_path):
    dedup_image_set(str(tmp_path))
    folder = str(tmp_path)
    paths = mm.get_all_images(folder)
    before = set(paths)
    deleted, refs, total = mm.find_and_remove_duplicate_images(folder, cosine_threshold=0.9999)
    assert total == len(paths)
    # oracle: same embedding, oracle pairs, oracle greedy walk in file-size order
    from mmrs_b200.dedup import pixel_embedding
    # (files were deleted; recompute the expectation from a fresh copy)
    import shutil
    fresh = tmp_path.parent / (tmp_path.name + "_fresh")
    fresh.mkdir()
    dedup_image_set(str(fresh))
    fpaths = mm.get_all_images(str(fresh))
    fpaths.sort(key=lambda p: os.path.getsize(p), reverse=True)
    emb, ok = pixel_embedding(fpaths)
    usable = [p for p, g in zip(fpaths, ok) if g]
    pairs = oracle.dedup_pairs(emb, 0.9999)
    reps, dups = oracle.greedy_keep_first(len(usable), pairs.tolist(), list(range(len(usable))))
    rel = lambda p, root: os.path.relpath(p, root)
    assert sorted(rel(p, folder) for p in refs) == sorted(rel(usable[r], fresh) for r in reps)
    assert sorted((rel(a, folder), rel(b, folder)) for a, b in deleted) == \
        sorted((rel(usable[d], fresh), rel(usable[o], fresh)) for d, o in dups)
    assert len(deleted) == 4                                   # a_copy, c_again/c_copy (2 of 3), d_copy
    assert set(mm.get_all_images(folder)) == before - {d for d, _ in deleted}
    shutil.rmtree(fresh)


def test_cross_folder_wrapper_matches_reference_golden(mm, tmp_path):
    """The GPU top-1 search finds the candidates, the reference's own predicate (equal MD5 of the RGB
    bytes) confirms them: the result tuple equals the one recorded from find_repeated.py, including
    WHICH of two identical reference images is reported (the last one, find_repeated.py:52)."""
    gold = json.loads((GOLDEN / "find_repeated_golden.json").read_text())
    ref_dir, del_dir = dedup_image_set(str(tmp_path))
    rel = lambda p: os.path.relpath(p, tmp_path)
    deleted, kept, n_ref, n_del = mm.find_and_remove_duplicate_images(ref_dir, del_dir)
    assert (n_ref, n_del) == (gold["n_ref"], gold["n_del"])
    assert sorted([rel(a), rel(b)] for a, b in deleted) == gold["deleted"]
    assert ["delete/nested/c_copy.png", "reference/sub/c_again.png"] in gold["deleted"]      # the last duplicate won
    assert sorted(rel(p) for p in kept) == gold["kept"]
    assert sorted(rel(p) for p in mm.get_all_images(del_dir)) == gold["remaining_in_delete_folder"]


def _save(arr, path):
    from PIL import Image
    Image.fromarray(arr, "RGB").save(path)


def test_cross_folder_default_deletes_only_pixel_identical_files(mm, tmp_path):
    """The default embedder is not injective (16 x 16 box resize, mean removed, unit norm): a brighter
    copy, a flat-colour image and an upscaled copy all score cos = 1 against a reference image.  The
    exact confirmation must keep them; only the pixel-identical file goes."""
    rng = np.random.default_rng(5)
    ref, dele = tmp_path / "ref", tmp_path / "del"
    ref.mkdir(); dele.mkdir()
    base = rng.integers(40, 160, size=(32, 32, 3), dtype=np.uint8)
    flat = np.full((32, 32, 3), 90, dtype=np.uint8)
    _save(base, ref / "base.png")
    _save(flat, ref / "flat.png")
    _save(base, dele / "identical.bmp")
    _save((base.astype(np.int32) + 40).astype(np.uint8), dele / "brighter.png")
    _save(np.full((32, 32, 3), 200, dtype=np.uint8), dele / "other_flat.png")
    _save(np.kron(base, np.ones((2, 2, 1), dtype=np.uint8)), dele / "upscaled.png")
    from mmrs_b200.dedup import pixel_embedding
    emb_r, _ = pixel_embedding([str(ref / "base.png")])
    emb_d, _ = pixel_embedding([str(dele / "brighter.png"), str(dele / "upscaled.png")])
    assert float((emb_d @ emb_r.T).min()) > 0.9999             # the embedding alone would delete them
    deleted, kept, n_ref, n_del = mm.find_and_remove_duplicate_images(str(ref), str(dele))
    assert [os.path.basename(a) for a, _ in deleted] == ["identical.bmp"]
    assert sorted(os.path.basename(p) for p in kept) == ["brighter.png", "other_flat.png", "upscaled.png"]
    assert sorted(os.listdir(dele)) == ["brighter.png", "other_flat.png", "upscaled.png"]
    # an explicit embedding predicate (confirm="none") is the caller's decision
    deleted, kept, _, _ = mm.find_and_remove_duplicate_images(str(ref), str(dele), confirm="none", dry_run=True)
    assert len(deleted) == 3 and kept == []


@pytest.mark.parametrize("radius", [0, 5, 10, 1.5, 0.0])
def test_hamming_radius_is_refused_not_reinterpreted(mm, tmp_path, radius):
    """The reference's same-folder form takes a Hamming radius (0 = strictest, its main passes 5).  Neither
    value may be read as a cosine: 5 would never match, 0 would delete nearly the whole folder."""
    dedup_image_set(str(tmp_path))
    before = sorted(mm.get_all_images(str(tmp_path)))
    with pytest.raises(ValueError, match="Hamming radius"):
        mm.find_and_remove_duplicate_images(str(tmp_path), radius)
    with pytest.raises(ValueError, match="Hamming radius"):
        mm.find_and_remove_duplicate_images(str(tmp_path), cosine_threshold=radius)
    with pytest.raises(ValueError, match="Hamming radius"):
        mm.find_and_remove_near_duplicate_images(str(tmp_path), radius)
    assert sorted(mm.get_all_images(str(tmp_path))) == before      # nothing was deleted


def test_non_unit_rows_keep_the_exact_pair_set(mm, oracle):
    """The bf16 prefilter's margin bounds the rounding of UNIT rows; it is scaled by the largest squared
    row norm, so raw (non-normalised) embeddings give the exact mode's pair set on the tensor-core path."""
    gen = torch.Generator().manual_seed(11)
    n, d = 4096, 128
    x = oracle.l2_normalize(torch.randn(n, d, generator=gen)) * 3.0           # every row has norm 3
    tau = 0.95 * 9.0
    for t, c in enumerate(torch.linspace(0.940, 0.960, 200).tolist()):
        a = x[t] / 3.0
        r = torch.randn(d, generator=gen)
        r = oracle.l2_normalize(r - (r @ a) * a)
        x[2000 + t] = 3.0 * (c * a + (1 - c * c) ** 0.5 * r)
    want = mm.find_duplicate_pairs(x, tau, method="fp32")
    got = mm.find_duplicate_pairs(x, tau, method="tc")
    assert torch.equal(got, want) and 60 < got.shape[0] < 140
    assert torch.equal(mm.find_duplicate_pairs(x, tau), want)                 # "auto" takes the tensor-core path here
    from mmrs_b200.dedup import row_norm_range, _device_f32
    lo, hi = row_norm_range(_device_f32(x))
    assert abs(lo - 3.0) < 1e-4 and abs(hi - 3.0) < 1e-4


def test_library_pair_sort(mm):
    """mmrs_sort_pairs (bitonic, packed keys) against numpy's lexsort, sizes around the tile boundaries."""
    from mmrs_b200.dedup import sort_pairs
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 1000, 2047, 2048, 2049, 4096, 5000, 70_001, 300_000):
        a = rng.integers(0, 1 << 31, size=(n, 2), dtype=np.int64)
        a[: n // 3, 0] = 7                                         # many equal first components
        got = sort_pairs(torch.from_numpy(a).cuda(), 1 << 31).cpu().numpy()
        want = a[np.lexsort((a[:, 1], a[:, 0]))]
        np.testing.assert_array_equal(got, want)


# ---- tensor-core prefilter + exact recheck -----------------------------------------------------------
@pytest.mark.parametrize("n,d,frac", [(2048, 64, 0.05), (5000, 512, 0.02), (20_000, 128, 0.01), (33_333, 768, 0.01),
                                      (100_000, 64, 0.002)])
def test_tc_pairs_bit_exact(mm, oracle, n, d, frac):
    x, planted = oracle.synthetic_dedup(n, d, dup_frac=frac, seed=n)
    got = mm.find_duplicate_pairs(x, 0.95, method="tc")
    assert [tuple(p) for p in got.tolist()] == planted
    if n <= 20_000:
        assert torch.equal(got, oracle.dedup_pairs(x, 0.95))
        assert torch.equal(got, mm.find_duplicate_pairs(x, 0.95, method="fp32"))


def test_tc_near_threshold_pairs_follow_the_exact_mode(mm, oracle):
    """No guard band: pairs planted at cosines scattered around tau.  The bf16 pass may let extra
    candidates through, the fp32 recheck must then take exactly the exact mode's decisions."""
    gen = torch.Generator().manual_seed(7)
    n, d = 6000, 256
    x = oracle.l2_normalize(torch.randn(n, d, generator=gen))
    for t, c in enumerate(torch.linspace(0.940, 0.960, 400).tolist()):
        a = x[t]
        r = torch.randn(d, generator=gen)
        r = oracle.l2_normalize(r - (r @ a) * a)
        x[3000 + t] = c * a + (1 - c * c) ** 0.5 * r          # cos(x[t], x[3000 + t]) = c up to rounding
    want = mm.find_duplicate_pairs(x, 0.95, method="fp32")
    got = mm.find_duplicate_pairs(x, 0.95, method="tc")
    assert torch.equal(got, want) and 150 < got.shape[0] < 250


def test_tc_rank_panels_partition_the_join(mm, oracle):
    from mmrs_b200.dedup import selfjoin_tc_raw, sort_pairs, _device_f32
    x, planted = oracle.synthetic_dedup(40_000, 64, dup_frac=0.01, seed=3)
    xd = _device_f32(x)
    parts = [selfjoin_tc_raw(xd, 0.95, r, 3) for r in range(3)]
    assert sum(p.shape[0] for p in parts) == len(planted)
    got = sort_pairs(torch.cat(parts), 40_000).cpu()
    assert [tuple(p) for p in got.tolist()] == planted


@pytest.mark.parametrize("pair", ["0", "1"])
def test_tc_single_cta_and_cta_pair_kernels(mm, oracle, monkeypatch, pair):
    """Both schedules of the tensor-core join (chosen by estimated duration, forced here): single CTAs
    (M = 128) and CTA pairs (tcgen05.mma.cta_group::2, M = 256); odd i-block counts, rank panels."""
    from mmrs_b200.dedup import selfjoin_tc_raw, sort_pairs, _device_f32
    monkeypatch.setenv("MMRS_SJ_PAIR", pair)
    for n, d, seed in [(4_225, 64, 5), (33_333, 768, 33_333), (70_001, 128, 9)]:     # 4225 = 33 i-blocks + 1 row
        x, planted = oracle.synthetic_dedup(n, d, dup_frac=0.01, seed=seed)
        xd = _device_f32(x)
        got = sort_pairs(selfjoin_tc_raw(xd, 0.95), n).cpu()
        assert [tuple(p) for p in got.tolist()] == planted
        parts = [selfjoin_tc_raw(xd, 0.95, r, 3) for r in range(3)]
        assert [tuple(p) for p in sort_pairs(torch.cat(parts), n).cpu().tolist()] == planted


def test_tc_buffers_regrow(mm, oracle):
    x = oracle.synthetic_gallery(3000, 64, seed=1, dtype=torch.float32)
    from mmrs_b200.dedup import selfjoin_tc_raw, sort_pairs, _device_f32
    got = sort_pairs(selfjoin_tc_raw(_device_f32(x), 0.05, capacity=16), 3000).cpu()   # thousands of pairs
    assert torch.equal(got, oracle.dedup_pairs(x, 0.05))


def test_cross_set_leakage_wrapper(mm, tmp_path, capsys):
    """tool/delete repeated.py signature: train images that match a test image are deleted; returns None."""
    ref_dir, del_dir = dedup_image_set(str(tmp_path))       # reference = "test set", delete = "train set"
    before = set(mm.get_all_images(del_dir))
    out = mm.detect_and_remove_cross_set_duplicates(ref_dir, del_dir, 8, 0)
    assert out is None
    from mmrs_b200.dedup import last_cross_set_summary as s
    # this tool's extension list has .webp but no .tiff (delete repeated.py:35): d.tiff is not part of the
    # test set, so d_copy.png stays
    assert s["test_images"] == 4 and s["train_images"] == 7
    assert s["duplicates_found"] == 2 and s["deleted_files"] == 2
    after = set(mm.get_all_images(del_dir))
    assert {os.path.basename(p) for p in before - after} == {"a_copy.bmp", "c_copy.png"}
    assert "操作摘要" in capsys.readouterr().out
    assert mm.detect_and_remove_cross_set_duplicates(str(tmp_path / "missing"), del_dir) is None


def test_gallery_from_feature_cache_and_path_lookup(mm, oracle, tmp_path):
    import pickle
    g = oracle.synthetic_gallery(300, 64, seed=2, dtype=torch.float32)
    d = {f"cls{i % 3}/{i}.jpg": g[i].numpy() for i in range(300)}
    with open(tmp_path / "features.pkl", "wb") as f:
        pickle.dump(d, f)                                   # the format of search_image.py:159-160
    gal = mm.DeviceGallery.from_feature_cache(str(tmp_path / "features.pkl"))
    v, i = mm.search_topk(g[7:9], gal, 3)
    paths = gal.lookup_paths(i)
    assert paths[0][0] == "cls1/7.jpg" and paths[1][0] == "cls2/8.jpg" and abs(float(v[0, 0]) - 1.0) < 1e-5
