"""Parity of the CUDA search path with the oracle (all calls go through the C ABI)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN
from golden_inputs import similarity_inputs

pytestmark = pytest.mark.gpu


def explain_index_mismatches(got_i, ext_i, ext_v, tol):
    """north_star: identical top-k indices, ties by index.  fp32 accumulation order differs between
    MKL and any GPU kernel, so a differing position is accepted ONLY if the
    reference's own scores of the two rows involved lie within `tol`.  `ext_i` / `ext_v` are the
    reference's top-(k+1) (k when the gallery has only k rows): position k - 1 is checked like every
    other one -- the row returned there must be a reference near-tie of position k - 1 or k.
    Returns #mismatches."""
    k = got_i.shape[1]
    assert ext_i.shape[1] >= k
    bad = (got_i != ext_i[:, :k])
    n_bad = int(bad.sum())
    if n_bad == 0:
        return 0
    for q, j in zip(*np.nonzero(bad)):
        # the row we returned at position j must be one whose reference score is within tol of
        # the reference's j-th score
        lo, hi = max(0, j - 8), min(ext_v.shape[1], j + 9)
        near = np.abs(ext_v[q, lo:hi] - ext_v[q, j]) <= tol
        cand = set(ext_i[q, lo:hi][near].tolist())
        assert int(got_i[q, j]) in cand, (q, j, got_i[q, j], ext_i[q, j])
    return n_bad


def oracle_topk_ext(oracle, q, g, k, **kw):
    """(want_v [Q, k], want_i [Q, k], ext_v, ext_i [Q, min(k + 1, N)]) from the oracle."""
    ext_v, ext_i = oracle.search_topk(q, g, min(k + 1, g.shape[0]), **kw)
    return ext_v[:, :k].contiguous(), ext_i[:, :k].contiguous(), ext_v, ext_i


def check_fp32(mm, oracle, g, q, k, max_swapped=0.005, **kw):
    want_v, want_i, ext_v, ext_i = oracle_topk_ext(oracle, q, g, k, mode="fp32", **kw)
    v, i = mm.search_topk(q, mm.DeviceGallery(g, mode="fp32"), k, **kw)
    assert v.shape == (q.shape[0], k) and i.dtype == torch.int64 and v.dtype == torch.float32
    scale = abs(kw.get("scale", 1.0))
    np.testing.assert_allclose(v.cpu().numpy(), want_v.numpy(), atol=1e-5 * max(scale, 1.0), rtol=0)
    n_bad = explain_index_mismatches(i.cpu().numpy(), ext_i.numpy(), ext_v.numpy(), 1e-5 * max(scale, 1.0))
    assert n_bad <= max(1, int(i.numel() * max_swapped)), n_bad
    # descending, ties by ascending index
    vv, ii = v.cpu().numpy(), i.cpu().numpy()
    assert np.all(vv[:, 1:] <= vv[:, :-1])
    same = vv[:, 1:] == vv[:, :-1]
    assert np.all(ii[:, 1:][same] > ii[:, :-1][same])
    return n_bad


@pytest.mark.parametrize("n,d,nq,k", [
    (1, 8, 1, 1), (5, 8, 3, 5), (127, 64, 2, 10), (128, 64, 4, 128), (129, 64, 5, 100),
    (1000, 100, 7, 10),            # D not a multiple of 8: zero-padded columns
    (4096, 512, 9, 100), (16384, 32, 3, 10), (16385, 32, 3, 10), (20000, 768, 8, 100),
    (70001, 64, 6, 1), (70001, 64, 17, 1024), (300000, 128, 2, 100),
])
def test_topk_fp32_matches_oracle(mm, oracle, n, d, nq, k):
    g = oracle.synthetic_gallery(n, d, seed=n % 97, dtype=torch.float32)
    q = oracle.synthetic_queries(nq, d, seed=d)
    check_fp32(mm, oracle, g, q, k)


def test_config_c1_shape_fp32(mm, oracle):
    """C1: 10k x 512 gallery, 100 queries, top-10 cosine, fp32 mode (BASELINE.json configs[0]).
    Nearly collinear features like a random-init ViT-B/32 gives: a common direction
    plus small noise, so scores differ only in the 3rd decimal -- the stress test for H1."""
    gen = torch.Generator().manual_seed(1)
    base = torch.randn(512, generator=gen)
    g = oracle.l2_normalize(base + 0.1 * torch.randn(10_000, 512, generator=gen))
    q = base + 0.1 * torch.randn(100, 512, generator=gen)
    # every swapped position is checked to be a reference near-tie (<= 1e-5); on this collinear
    # data about 1 % of positions are such near-ties (8 of 1000 measured on B200)
    n_bad = check_fp32(mm, oracle, g, q, 10, max_swapped=0.03)
    print("C1-shape near-tie index mismatches:", n_bad)


def test_scale_and_unnormalised_queries(mm, oracle):
    g = oracle.synthetic_gallery(3000, 64, seed=1, dtype=torch.float32)
    q = oracle.synthetic_queries(4, 64) * 3.0
    check_fp32(mm, oracle, g, q, 20, normalize_queries=False, scale=100.0)
    check_fp32(mm, oracle, g, q, 20, normalize_queries=True, scale=-1.0)   # negative scale: smallest cosine first


def test_exact_ties_fall_to_lower_index(mm, oracle):
    g = oracle.synthetic_gallery(40_000, 64, seed=2, dtype=torch.float32)
    q = oracle.synthetic_queries(3, 64)
    best = oracle.search_topk(q, g, 1)[1][:, 0]
    for qi, b in enumerate(best.tolist()):  # replicate each query's best row at scattered places
        for pos in (7, 12_345, 39_990, 20_000):
            g[pos + qi] = g[b]
    want_v, want_i = oracle.search_topk(q, g, 10)
    v, i = mm.search_topk(q, mm.DeviceGallery(g), 10)
    assert torch.equal(i.cpu(), want_i)
    # and the duplicates really are exact ties in OUR arithmetic
    vv = v.cpu().numpy()
    assert (vv[:, 0] == vv[:, 1]).all()


def test_bf16_mode_recall_and_scores(mm, oracle):
    g = oracle.synthetic_gallery(200_000, 512, seed=0, dtype=torch.bfloat16)
    q = oracle.synthetic_queries(4, 512)
    want_v, want_i = oracle.search_topk(q, g, 100, mode="bf16")
    v, i = mm.search_topk(q, mm.DeviceGallery(g), 100)
    assert mm.DeviceGallery(g).mode == "bf16"
    np.testing.assert_allclose(v.cpu().numpy(), want_v.numpy(), atol=1e-2, rtol=0)   # north_star tolerance
    hits = sum(len(set(a) & set(b)) for a, b in zip(i.cpu().tolist(), want_i.tolist()))
    assert hits / want_i.numel() >= 0.999
    # in practice far tighter: bf16 products are exact in fp32, only the summation order differs
    assert np.abs(v.cpu().numpy() - want_v.numpy()).max() < 1e-5


def test_overflow_falls_back_to_exhaustive(mm, oracle):
    """Scores correlated with the tile-stride pattern: every tile OUTSIDE the seed sample beats the
    sample, so the candidate list overflows and the exhaustive path must still be exact."""
    n, d = 70_000, 32
    g = oracle.synthetic_gallery(n, d, seed=5, dtype=torch.float32)
    q = oracle.synthetic_queries(2, d)
    qn = oracle.l2_normalize(q)
    tiles = torch.arange(n) // 128
    outside = (tiles % 16 != 0)
    g[outside] = oracle.l2_normalize(g[outside] + 2.0 * qn[0])      # all of them score high for query 0
    check_fp32(mm, oracle, g, q, 10)


def test_host_and_device_queries_agree(mm, oracle):
    g = oracle.synthetic_gallery(50_000, 256, seed=8, dtype=torch.bfloat16)
    gal = mm.DeviceGallery(g)
    q = oracle.synthetic_queries(3, 256)
    vh, ih = mm.search_topk(q, gal, 50)                 # host in -> host out (C ABI *_host entry)
    vd, idd = mm.search_topk(q.cuda(), gal, 50)         # device in -> device out
    assert not vh.is_cuda and vd.is_cuda
    assert torch.equal(ih, idd.cpu()) and torch.equal(vh, vd.cpu())
    # numpy in
    vn, inn = mm.search_topk(q.numpy(), gal, 50)
    assert torch.equal(inn, ih)


def test_row_offset_and_merge(mm, oracle):
    g = oracle.synthetic_gallery(30_000, 64, seed=3, dtype=torch.float32)
    g[10] = g[29_000]
    q = oracle.synthetic_queries(5, 64)
    want_v, want_i = oracle.search_topk(q, g, 25)
    vs, is_ = [], []
    for lo, hi in mm.shard_bounds(30_000, 4):
        v, i = mm.search_topk(q.cuda(), mm.DeviceGallery(g[lo:hi], row_offset=lo), 25)
        vs.append(v); is_.append(i)
    from mmrs_b200.sharded import _merge_cuda
    v, i = _merge_cuda(torch.stack(vs), torch.stack(is_), 25)
    assert torch.equal(i.cpu(), want_i)
    np.testing.assert_allclose(v.cpu().numpy(), want_v.numpy(), atol=1e-5)
    ov, oi = oracle.merge_topk(torch.stack(vs).cpu(), torch.stack(is_).cpu(), 25)
    assert torch.equal(i.cpu(), oi) and torch.equal(v.cpu(), ov)


def test_edge_cases(mm, oracle):
    g = oracle.synthetic_gallery(100, 16, seed=1, dtype=torch.float32)
    gal = mm.DeviceGallery(g)
    v, i = mm.search_topk(torch.empty(0, 16), gal, 5)
    assert v.shape == (0, 5) and i.shape == (0, 5)
    with pytest.raises(RuntimeError, match="out of range"):
        mm.search_topk(torch.randn(1, 16), gal, 101)
    q = torch.randn(3, 16); q[1] = 0
    with pytest.raises(mm._cabi.MmrsError) as e:
        mm.search_topk(q, gal, 5)                       # reference would give NaN (no eps at :157)
    assert e.value.code == mm._cabi.ERR_ZERO_NORM
    v, i = mm.search_topk(q, gal, 5, normalize_queries=False)   # a zero query is fine un-normalised
    assert i[1].tolist() == [0, 1, 2, 3, 4] and (v[1] == 0).all()
    with pytest.raises(ValueError):
        mm.search_topk(torch.randn(1, 8), gal, 5)


def test_full_scores_and_get_similarity_golden(mm, oracle):
    gold = np.load(GOLDEN / "search_image_golden.npz")
    for name, (features, targets, label, ref) in similarity_inputs().items():
        pos, neg = mm.get_similarity(features, targets, label, ref)
        assert pos.dtype == np.float32 and pos.shape == gold[f"{name}_pos"].shape
        # scale 100: 1e-5 relative to the cosine == 1e-3 on these scores
        np.testing.assert_allclose(pos, gold[f"{name}_pos"], atol=1e-3, rtol=0)
        np.testing.assert_allclose(neg, gold[f"{name}_neg"], atol=1e-3, rtol=0)
        s = mm.full_scores(ref[None], features, normalize_queries=False, scale=1.0)
        want = oracle.full_scores(ref[None], features, normalize_queries=False)
        np.testing.assert_allclose(s.numpy(), want.numpy(), atol=1e-5, rtol=0)
    # multi-class scoring in one call (CLIP/lab3.py:113-114 loops over 5 classes)
    g = oracle.synthetic_gallery(5000, 768, seed=4, dtype=torch.float32)
    t = oracle.synthetic_queries(5, 768)
    s = mm.full_scores(t, g)
    np.testing.assert_allclose(s.numpy(), oracle.full_scores(t, g).numpy(), atol=1e-5, rtol=0)


def test_threshold_sweep_golden(mm, oracle):
    gold = np.load(GOLDEN / "search_image_golden.npz")
    for name in ("small", "clip512", "taiyi768"):
        pos, neg = gold[f"{name}_pos"], gold[f"{name}_neg"]
        assert mm.find_thresholds(pos, neg, name) == gold[f"{name}_best_f1"]
        for t, want in zip(gold[f"{name}_probe"], gold[f"{name}_eval"]):
            got = np.array(mm.eval_threshold(pos, neg, t), dtype=np.float64)
            np.testing.assert_array_equal(got, want)
    rng = np.random.default_rng(0)
    pos = rng.normal(1, 1, 100_000).astype(np.float32); neg = rng.normal(0, 1, 1_000_000).astype(np.float32)
    thr = np.linspace(min(pos.min(), neg.min()), max(pos.max(), neg.max()), 200)
    tp, fp = mm.threshold_sweep_counts(pos, neg, thr)
    np.testing.assert_array_equal(tp, [(pos >= t).sum() for t in thr])
    np.testing.assert_array_equal(fp, [(neg >= t).sum() for t in thr])


@pytest.mark.parametrize("nq", [1, 4])
def test_full_size_c2_small_batch(mm, oracle, nq):
    """C2 at full size: 1M x 512 bf16, top-100 (the oracle takes seconds for <= 4 queries)."""
    g = oracle.synthetic_gallery(1_000_000, 512, seed=0, dtype=torch.bfloat16)
    q = oracle.synthetic_queries(nq, 512, seed=1)
    want_v, want_i, ext_v, ext_i = oracle_topk_ext(oracle, q, g, 100, mode="bf16")
    v, i = mm.search_topk(q, mm.DeviceGallery(g), 100)
    np.testing.assert_allclose(v.numpy(), want_v.numpy(), atol=1e-2, rtol=0)
    hits = sum(len(set(a) & set(b)) for a, b in zip(i.tolist(), want_i.tolist()))
    assert hits / want_i.numel() >= 0.999
    assert explain_index_mismatches(i.numpy(), ext_i.numpy(), ext_v.numpy(), 1e-6) <= 2


# ---- K2 (tcgen05) explicitly -----------------------------------------------------------------------
def bf16_unit_queries(q):
    """Normalise and round to bf16 ONCE on the host.  Strict comparisons feed these to both sides with
    normalize_queries=False: rounding the normalised query to bf16 is discontinuous, so two
    implementations whose fp32 norms differ in the last ulp flip a bf16 query element about once
    per 1e5 elements, which moves that query's scores by ~1e-4 (seen at 2 500 x 128 queries) --
    inside north_star's 1e-2, far outside a 1e-5 check."""
    return (q / q.norm(dim=-1, keepdim=True)).to(torch.bfloat16).to(torch.float32)


def check_bf16(mm, oracle, g, q, k, path, **kw):
    if kw.get("normalize_queries", True):
        # fused normalisation: north_star's own tolerances (scores 1e-2, recall@k 0.999) ...
        lv, li = mm.search_topk(q, mm.DeviceGallery(g), k, path=path, **kw)
        ov, oi = oracle.search_topk(q, g, k, mode="bf16", **kw)
        np.testing.assert_allclose(lv.cpu().numpy(), ov.numpy(), atol=1e-2, rtol=0)
        hits = sum(len(set(a) & set(b)) for a, b in zip(li.cpu().tolist(), oi.tolist()))
        assert hits / oi.numel() >= 0.999
        # ... then the strict comparison on identical bf16-valued operands
        q = bf16_unit_queries(q)
        kw = dict(kw, normalize_queries=False)
    want_v, want_i, ext_v, ext_i = oracle_topk_ext(oracle, q, g, k, mode="bf16", **kw)
    v, i = mm.search_topk(q, mm.DeviceGallery(g), k, path=path, **kw)
    v, i = v.cpu().numpy(), i.cpu().numpy()
    np.testing.assert_allclose(v, want_v.numpy(), atol=1e-5, rtol=0)       # far inside the 1e-2 of north_star
    n_bad = explain_index_mismatches(i, ext_i.numpy(), ext_v.numpy(), 2e-6)
    hits = sum(len(set(a) & set(b)) for a, b in zip(i.tolist(), want_i.tolist()))
    assert hits / want_i.numel() >= 0.999
    assert np.all(v[:, 1:] <= v[:, :-1])
    return n_bad


@pytest.mark.parametrize("n,d,nq,k", [
    (128, 64, 16, 10), (100, 64, 1, 5), (1000, 512, 5, 10), (4096, 72, 7, 33),   # D=72: K tail zero-filled by TMA
    (20_000, 768, 33, 100), (70_001, 520, 64, 100), (300_000, 512, 256, 100),
    (50_000, 512, 300, 10),                                                     # > 256 queries: two passes
    (16_385, 104, 100, 1),
])
def test_mma_path_matches_oracle(mm, oracle, n, d, nq, k):
    g = oracle.synthetic_gallery(n, d, seed=(n + d) % 89, dtype=torch.bfloat16)
    q = oracle.synthetic_queries(nq, d, seed=nq)
    check_bf16(mm, oracle, g, q, k, "mma")


@pytest.mark.parametrize("n,d,nq,k", [
    (20_000, 768, 70, 20),       # UMMA N = 96: each CTA of the pair stages 48 query rows
    (50_000, 512, 96, 100),
    (70_001, 512, 130, 100),     # N = 160
    (40_000, 768, 200, 50),      # N = 224
    (8_321, 64, 700, 5),         # three query chunks as work units, odd tile count, few tiles
    (33_000, 520, 1024, 10),     # four chunks
    (129, 64, 129, 3),           # two tiles = one pair, single phase
    (128, 64, 256, 128),         # one tile: the second CTA of the pair has nothing to read
])
def test_mma_cta_pair_mode_matches_oracle(mm, oracle, n, d, nq, k):
    """More than 64 queries run as CTA pairs (tcgen05.mma.cta_group::2, M = 256)."""
    g = oracle.synthetic_gallery(n, d, seed=(n + d) % 89, dtype=torch.bfloat16)
    q = oracle.synthetic_queries(nq, d, seed=nq)
    check_bf16(mm, oracle, g, q, k, "mma")


def test_mma_cta_pair_and_single_cta_agree_bitwise(mm, oracle, monkeypatch):
    g = oracle.synthetic_gallery(90_000, 512, seed=8, dtype=torch.bfloat16)
    gal = mm.DeviceGallery(g)
    q = bf16_unit_queries(oracle.synthetic_queries(600, 512, seed=4))
    v2, i2 = mm.search_topk(q, gal, 100, path="mma", normalize_queries=False)
    s2 = mm.full_scores(q[:200], gal, path="mma", normalize_queries=False)
    monkeypatch.setenv("MMRS_K2_NO_PAIR", "1")
    v1, i1 = mm.search_topk(q, gal, 100, path="mma", normalize_queries=False)
    s1 = mm.full_scores(q[:200], gal, path="mma", normalize_queries=False)
    assert torch.equal(v1, v2) and torch.equal(i1, i2)     # same MMA shapes along K, same accumulation order
    assert torch.equal(s1, s2)
    want = oracle.full_scores(q[:200], g, mode="bf16", normalize_queries=False)
    np.testing.assert_allclose(s2.cpu().numpy(), want.numpy(), atol=1e-5, rtol=0)


def test_mma_and_gemv_paths_agree_bitwise_on_indices(mm, oracle):
    g = oracle.synthetic_gallery(150_000, 512, seed=6, dtype=torch.bfloat16)
    gal = mm.DeviceGallery(g)
    q = oracle.synthetic_queries(8, 512, seed=2)
    v1, i1 = mm.search_topk(q, gal, 100, path="gemv")
    v2, i2 = mm.search_topk(q, gal, 100, path="mma")
    np.testing.assert_allclose(v1.numpy(), v2.numpy(), atol=2e-6, rtol=0)
    assert (i1 == i2).float().mean().item() > 0.995     # only summation-order near-ties may swap


def test_mma_full_scores(mm, oracle):
    g = oracle.synthetic_gallery(10_000, 768, seed=4, dtype=torch.bfloat16)
    t = oracle.synthetic_queries(21, 768)
    s = mm.full_scores(t, g, path="mma", scale=100.0)
    want = oracle.full_scores(t, g, mode="bf16", scale=100.0)
    np.testing.assert_allclose(s.numpy(), want.numpy(), atol=1e-4, rtol=0)


@pytest.mark.parametrize("nq", [16, 64])
def test_full_size_c2_batched(mm, oracle, nq):
    """C2 at full size through the auto path (K2 for nq > 4)."""
    g = oracle.synthetic_gallery(1_000_000, 512, seed=0, dtype=torch.bfloat16)
    q = oracle.synthetic_queries(nq, 512, seed=1)
    n_bad = check_bf16(mm, oracle, g, q, 100, "auto")
    print(f"C2 nq={nq}: near-tie index swaps {n_bad} of {nq * 100}")


def test_clustered_gallery_grouped_by_class(mm, oracle):
    """A gallery stored class by class (the reference's cache order, search_image.py:146-149) with
    the query's own class LAST: a prefix sample would never see it; the strided sample does."""
    gen = torch.Generator().manual_seed(3)
    centers = torch.randn(10, 256, generator=gen)
    labels = torch.arange(10).repeat_interleave(30_000)
    g = oracle.l2_normalize(centers[labels] + 0.7 * torch.randn(300_000, 256, generator=gen)).to(torch.bfloat16)
    q = centers[9:10] + 0.1 * torch.randn(12, 256, generator=gen)
    check_bf16(mm, oracle, g, q, 100, "auto")
    check_bf16(mm, oracle, g, q[:3], 100, "gemv")


# ---- larger-than-oracle sizes: a torch-on-GPU reference (test infrastructure only) -----------------
def torch_gpu_topk(qn, gal_data, k, block=1 << 20):
    """fp32 matmul of the same bf16-valued operands (qn = bf16_unit_queries) on the GPU, blocked over
    rows, stable order."""
    qn = qn.cuda()
    best_v = best_i = None
    for lo in range(0, gal_data.shape[0], block):
        s = qn @ gal_data[lo:lo + block].to(torch.float32).t()
        v, i = torch.sort(s, dim=1, descending=True, stable=True)
        v, i = v[:, :k], i[:, :k] + lo
        if best_v is None:
            best_v, best_i = v, i
        else:
            cv, ci = torch.cat([best_v, v], 1), torch.cat([best_i, i], 1)
            o = torch.sort(cv, dim=1, descending=True, stable=True)[1][:, :k]
            best_v, best_i = torch.gather(cv, 1, o), torch.gather(ci, 1, o)
    return best_v.cpu(), best_i.cpu()


def test_c4_shard_size_12p5m_x_768(mm):
    """One GPU's share of C4 (100M x 768 over 8 GPUs = 12.5M rows, 19.2 GB): 4-phase plan."""
    from bench import device_gallery_shard
    dev = torch.device("cuda", 0)
    data = device_gallery_shard(torch, 12_500_000, 768, 0, 0, dev)
    gal = mm.DeviceGallery(data, row_offset=25_000_000)
    q = bf16_unit_queries(torch.randn(16, 768, generator=torch.Generator().manual_seed(3)))
    v, i = mm.search_topk(q, gal, 100, normalize_queries=False)
    wv, wi = torch_gpu_topk(q, data, 101)
    np.testing.assert_allclose(v.numpy(), wv[:, :100].numpy(), atol=2e-6, rtol=0)
    assert explain_index_mismatches(i.numpy() - 25_000_000, wi.numpy(), wv.numpy(), 2e-6) <= 8
    v1, i1 = mm.search_topk(q[:2], gal, 100, path="gemv", normalize_queries=False)
    assert (i1 == i[:2]).float().mean().item() > 0.99
    del gal, data
    torch.cuda.empty_cache()


def test_many_queries_super_chunks(mm, oracle):
    """2 500 queries: more than one 1024-query super-chunk and ten 256-query K2 passes."""
    g = oracle.synthetic_gallery(120_000, 128, seed=12, dtype=torch.bfloat16)
    q = bf16_unit_queries(oracle.synthetic_queries(2500, 128, seed=13))
    v, i = mm.search_topk(q.cuda(), mm.DeviceGallery(g), 10, normalize_queries=False)
    wv, wi = torch_gpu_topk(q, g.cuda(), 11)
    np.testing.assert_allclose(v.cpu().numpy(), wv[:, :10].numpy(), atol=2e-6, rtol=0)
    assert explain_index_mismatches(i.cpu().numpy(), wi.numpy(), wv.numpy(), 2e-6) <= 25
    ov, oi = oracle.search_topk(q[:64], g, 11, mode="bf16", normalize_queries=False)
    assert explain_index_mismatches(i[:64].cpu().numpy(), oi.numpy(), ov.numpy(), 2e-6) <= 2


def test_async_search_handles(mm, oracle):
    g = oracle.synthetic_gallery(100_000, 256, seed=14, dtype=torch.bfloat16)
    gal = mm.DeviceGallery(g)
    qs = [oracle.synthetic_queries(8, 256, seed=s) for s in range(4)]
    pend = [mm.search_topk(q, gal, 20, sync=False) for q in qs]          # four host batches in flight
    for q, pnd in zip(qs, pend):
        v, i = pnd.wait()
        wv, wi = mm.search_topk(q, gal, 20)
        assert torch.equal(i, wi) and torch.equal(v, wv)
    pd = mm.search_topk(qs[0].cuda(), gal, 20, sync=False)
    assert torch.equal(pd.wait()[1].cpu(), pend[0].wait()[1])


def test_device_resident_threshold_sweep(mm, oracle):
    """f2: scores stay on the GPU; grid, counts and best F1 equal the reference's find_thresholds."""
    gold = np.load(GOLDEN / "search_image_golden.npz")
    for name, (features, targets, label, ref) in similarity_inputs().items():
        gal = mm.DeviceGallery(features, mode="fp32")
        scores = mm.full_scores(ref[None].cuda(), gal, normalize_queries=False, scale=100.0)[0]   # CUDA [N]
        best_f1, best_thr, p, r, thresholds, f1s = mm.best_threshold_on_device(scores, targets, label)
        s_host = scores.cpu().numpy()
        want = oracle.find_thresholds(s_host[targets == label], s_host[targets != label])
        np.testing.assert_array_equal(thresholds, want[4])           # the linspace grid, bit for bit
        np.testing.assert_array_equal(f1s, want[5])
        assert (best_f1, best_thr) == (want[0], want[1])
        # and against the recorded output of the reference's own function (scores differ by <= 1e-3)
        assert abs(best_f1 - float(gold[f"{name}_best_f1"])) < 0.02


# ---- fp32 mode on the tensor cores (bf16 x 3 split, six MMAs per tile) ----------------------------------
@pytest.mark.parametrize("n,d,nq,k", [(4096, 512, 9, 100), (20_000, 768, 48, 10), (70_001, 72, 49, 100),
                                      (300_000, 128, 100, 10), (16_385, 520, 130, 1)])
def test_fp32_tensor_core_path_matches_oracle(mm, oracle, n, d, nq, k):
    g = oracle.synthetic_gallery(n, d, seed=(n * 7 + d) % 83, dtype=torch.float32)
    q = oracle.synthetic_queries(nq, d, seed=nq + 1)
    gal = mm.DeviceGallery(g, mode="fp32")
    want_v, want_i, ext_v, ext_i = oracle_topk_ext(oracle, q, g, k, mode="fp32")
    v, i = mm.search_topk(q, gal, k)                       # >= 9 queries: split planes + tcgen05
    assert gal._split is not None
    np.testing.assert_allclose(v.numpy(), want_v.numpy(), atol=1e-5, rtol=0)
    n_bad = explain_index_mismatches(i.numpy(), ext_i.numpy(), ext_v.numpy(), 1e-5)
    assert n_bad <= max(1, i.numel() // 200)
    # the CUDA-core exact path agrees to fp32 rounding
    v1, i1 = mm.search_topk(q[:8], gal, k, path="gemv")
    np.testing.assert_allclose(v1.numpy(), v[:8].numpy(), atol=2e-6, rtol=0)


def test_fp32_tensor_core_c1_shape_and_full_scores(mm, oracle):
    gen = torch.Generator().manual_seed(1)
    base = torch.randn(512, generator=gen)
    g = oracle.l2_normalize(base + 0.1 * torch.randn(10_000, 512, generator=gen))   # C1: collinear features
    q = base + 0.1 * torch.randn(100, 512, generator=gen)
    check_fp32(mm, oracle, g, q, 10, max_swapped=0.03)
    s = mm.full_scores(q, mm.DeviceGallery(g, mode="fp32"), scale=100.0)
    np.testing.assert_allclose(s.numpy(), oracle.full_scores(q, g, scale=100.0).numpy(), atol=1e-3, rtol=0)


def test_lab3_evaluate_thresholds_golden(mm):
    """f3: the lab scripts' 1001-point sweep through the GPU histogram equals the reference's own output."""
    from golden_inputs import lab3_inputs
    g = np.load(GOLDEN / "lab3_golden.npz")
    sims, thresholds, pos_cls, neg_cls = lab3_inputs()
    res = mm.evaluate_thresholds(sims, thresholds, pos_cls, neg_cls)
    assert len(res) == 1001 and set(res[0]) == {"threshold", "precision", "recall", "f1", "TP", "FP", "TN", "FN"}
    for key in ("threshold", "precision", "recall", "f1", "TP", "FP", "TN", "FN"):
        np.testing.assert_array_equal(np.array([r[key] for r in res], dtype=np.float64), g[key])
    best = max(res, key=lambda x: x["f1"])              # lab3.py:123
    assert best["f1"] == g["f1"].max()


def test_score_classes_and_union_metrics_golden(mm):
    """CLIP/union_dataset.py process_images (:247-260) + calc_combined_metrics (:133-231): all images x
    all classes in one GPU scoring call (D = 512 and D = 768) against the outputs recorded from the
    reference functions; the EN-or-CN union metrics computed from the GPU scores equal the recorded ones."""
    import json
    from golden_inputs import union_inputs
    g = json.loads((GOLDEN / "union_golden.json").read_text())
    u = union_inputs()
    sims = {}
    for key, pos, threshs in (("en", u["en_pos"], u["en_threshs"]), ("cn", u["cn_pos"], u["cn_threshs"])):
        feats, text, labels, paths = u[key]
        sims[key] = mm.score_classes(feats, text, pos, labels, paths)
        for cls, th in zip(pos, threshs):
            want = g["sims"][key][cls]
            got = sims[key][cls]
            assert [[it["true_label"], it["file_path"]] for it in got] == [w[1:] for w in want]
            w = np.array([x[0] for x in want])
            np.testing.assert_allclose([it["similarity"] for it in got], w, atol=1e-5, rtol=0)
            assert np.abs(w - th).min() > 1e-5          # no recorded score sits on its threshold
    got = mm.calc_combined_metrics(sims["en"], sims["cn"], u["en_threshs"], u["cn_threshs"],
                                   u["en_pos"], u["en_neg"], u["cn_pos"], u["cn_neg"])
    assert got == g["combined"]


def test_overlap_grid_find_thresholds_golden(mm):
    """find_thresholds(grid="overlap") = code/main_custom.py:46-91 on the GPU sweep; golden from the reference."""
    from golden_inputs import overlap_grid_inputs
    g = np.load(GOLDEN / "main_custom_golden.npz")
    for name, (pos, neg) in overlap_grid_inputs().items():
        if str(g[f"{name}_raises"]):
            with pytest.raises(ValueError):
                mm.find_thresholds(pos, neg, name, grid="overlap")
        else:
            assert mm.find_thresholds(pos, neg, name, grid="overlap") == g[f"{name}_best_f1"]


# ---- round 2: graph cache keyed on the workspace slot, slots, float64 sweep, new mirrors --------------------
def _graph_stats(mm):
    import ctypes as C
    out = (C.c_int64 * 4)()
    mm._cabi.check(mm._cabi.lib.mmrs_graph_stats(out))
    return {"captures": out[0], "replays": out[1], "patches": out[2], "unpatchable": out[3]}


def test_retained_outputs_capture_once(mm, oracle):
    """1 000 searches whose results (and queries) are all kept alive: every call brings new tensors, yet
    the library captures ONE graph for the slot and only re-points its prep / final-select nodes."""
    g = oracle.synthetic_gallery(60_000, 128, seed=3, dtype=torch.bfloat16)
    gal = mm.DeviceGallery(g)
    base = oracle.synthetic_queries(1000 * 4, 128, seed=8).cuda().view(1000, 4, 128)
    mm.search_topk(base[0], gal, 10)                            # slot + first capture
    s0 = _graph_stats(mm)
    kept = [mm.search_topk(base[t], gal, 10) for t in range(1000)]
    torch.cuda.synchronize()
    s1 = _graph_stats(mm)
    assert s1["captures"] == s0["captures"], (s0, s1)           # nothing was captured again
    assert s1["unpatchable"] == s0["unpatchable"] == 0
    assert s1["replays"] - s0["replays"] == 1000 and s1["patches"] - s0["patches"] >= 999
    assert len({v.data_ptr() for v, _ in kept}) == 1000         # the results really are distinct live tensors
    for t in (0, 1, 500, 999):                                  # and each holds ITS query's answer
        wv, wi = oracle.search_topk(base[t].cpu(), g, 10, mode="bf16")
        assert torch.equal(kept[t][1].cpu(), wi)
    # out=: caller-owned result tensors, host staging keeps the host path on one graph too
    v = torch.empty((4, 10), dtype=torch.float32, device="cuda"); i = torch.empty((4, 10), dtype=torch.int64, device="cuda")
    rv, ri = mm.search_topk(base[7], gal, 10, out=(v, i))
    assert rv is v and ri is i and torch.equal(i, kept[7][1])
    hq = [base[t].cpu() for t in range(20)]
    mm.search_topk(hq[0], gal, 10)
    s2 = _graph_stats(mm)
    host_kept = [mm.search_topk(q, gal, 10) for q in hq]
    assert _graph_stats(mm)["captures"] == s2["captures"]
    assert all(torch.equal(h[1], kept[t][1].cpu()) for t, h in enumerate(host_kept))
    with pytest.raises(ValueError):
        mm.search_topk(base[0], gal, 10, out=(v.cpu(), i.cpu()))


def test_search_slots_are_capped(mm, oracle):
    g = oracle.synthetic_gallery(5_000, 64, seed=3, dtype=torch.bfloat16)
    gal = mm.DeviceGallery(g)
    want = {}
    for nq in range(1, 41):                                     # 40 batch sizes > MAX_SEARCH_SLOTS
        q = oracle.synthetic_queries(nq, 64, seed=nq)
        v, i = mm.search_topk(q.cuda(), gal, 5)
        want[nq] = i.cpu()
    slots = [k for k in gal._workspaces if k[0] == "search"]
    assert len(slots) == gal.MAX_SEARCH_SLOTS
    for nq in (1, 2, 40, 17):                                   # evicted and surviving shapes both still work
        q = oracle.synthetic_queries(nq, 64, seed=nq)
        assert torch.equal(mm.search_topk(q.cuda(), gal, 5)[1].cpu(), want[nq])
    # a search workspace no longer carries the exhaustive buffer (8 bytes per row)
    lib = mm._cabi.lib
    small = lib.mmrs_search_workspace_bytes(100_000_000, 768, 1, 16, 100)
    assert small < 64 << 20 and lib.mmrs_search_exhaustive_workspace_bytes(100_000_000, 768, 16) > 800_000_000


def test_threshold_sweep_float64_scores(mm):
    """float64 scores are compared in float64 (numpy's compare), never rounded to float32 first."""
    rng = np.random.default_rng(3)
    base = rng.normal(20, 3, 5000)
    pos = base[:1000].copy(); neg = base[1000:].copy()
    thr = np.linspace(base.min(), base.max(), 200)
    # plant scores a hair below / above grid points: rounding them to float32 would move them across
    for j in range(10, 190, 7):
        pos[j] = np.nextafter(thr[j], -np.inf)
        neg[j] = np.nextafter(thr[j], np.inf)
    tp, fp = mm.threshold_sweep_counts(pos, neg, thr)
    np.testing.assert_array_equal(tp, [(pos >= t).sum() for t in thr])
    np.testing.assert_array_equal(fp, [(neg >= t).sum() for t in thr])
    tp32, _ = mm.threshold_sweep_counts(pos.astype(np.float32), neg.astype(np.float32), thr)
    assert not np.array_equal(tp32, tp)                           # the float32 path WOULD differ on this data
    # the drop-ins take python lists / float64 arrays like the reference does
    f1 = mm.find_thresholds(list(pos), list(neg), "f64")
    import oracle.oracle as O
    assert f1 == O.find_thresholds(pos, neg)[0]
    got = mm.eval_threshold(pos, neg, float(thr[17]))
    assert tuple(got) == tuple(O.eval_threshold(pos, neg, float(thr[17])))


def test_cls_acc_and_topk_of_scores_golden(mm, oracle):
    """cls_acc (code/utils.py:15-39) on the GPU select kernel against the reference's recorded accuracies."""
    from golden_inputs import topk_inputs
    gold = np.load(GOLDEN / "utils_topk_golden.npz")
    for name, (logits, target) in topk_inputs().items():
        for k in (1, 3):
            v, i = mm.topk_of_scores(logits, k)
            wv, wi = oracle.topk_rows(logits, k)
            assert torch.equal(v, wv) and torch.equal(i, wi)            # values exact, ties by ascending index
            np.testing.assert_array_equal(v.numpy(), gold[f"{name}_topk{k}_values"])
            if name == "random":                                         # no ties: indices equal torch.topk's too
                np.testing.assert_array_equal(i.numpy(), gold[f"{name}_topk{k}_indices"])
                assert mm.cls_acc(logits, target, topk=k) == float(gold[f"{name}_acc_k{k}"])
                assert mm.cls_acc(logits.cuda(), target.cuda(), topk=k) == float(gold[f"{name}_acc_k{k}"])
    logits, target = topk_inputs()["random"]
    keep = target != 2
    want = 100 * float((logits.argmax(1) == target)[keep].float().sum()) / int(keep.sum())
    assert mm.cls_acc(logits, target, 1, exclude_class=2) == want
    assert mm.cls_acc(logits[:3], torch.full((3,), 4), 1, exclude_class=4) == 0.0
    big = torch.randn(257, 1000, generator=torch.Generator().manual_seed(1))
    v, i = mm.topk_of_scores(big.cuda(), 5)
    wv, wi = oracle.topk_rows(big, 5)
    assert torch.equal(i.cpu(), wi) and torch.equal(v.cpu(), wv)
    with pytest.raises(RuntimeError, match="out of range"):
        mm.topk_of_scores(logits, 7)


def test_cosine_similarity_scores_golden(mm):
    """logit_scale * F.cosine_similarity (code/merge_dataset.py:275-278) through the gallery scan, against
    the similarity and the thresholded predictions recorded from the reference's clip_en_predict."""
    from golden_inputs import cosine_inputs
    gold = np.load(GOLDEN / "merge_dataset_golden.npz")
    x, t, logit_scale = cosine_inputs()
    got = mm.cosine_similarity_scores(x, t, logit_scale)
    assert got.shape == (300,) and not got.is_cuda
    np.testing.assert_allclose(got.numpy(), gold["similarity"], atol=1e-3, rtol=0)      # scale 100: 1e-5 on the cosine
    for thr in (-3.0, 0.0, 2.5, 70.0):
        pred = (got.numpy() < thr).astype(np.int64)
        far = np.abs(gold["similarity"] - thr) > 1e-3
        np.testing.assert_array_equal(pred[far], gold[f"preds_{thr:g}"][far])
        assert far.mean() > 0.98
    # torch's clamp semantics: a zero row scores 0 (the x / x.norm() idiom would give NaN)
    xz = x.clone(); xz[5] = 0
    gz = mm.cosine_similarity_scores(xz.cuda(), t.cuda(), 100.0)
    want = 100.0 * torch.nn.functional.cosine_similarity(xz, t)
    assert gz.is_cuda and float(gz[5]) == 0.0
    np.testing.assert_allclose(gz.cpu().numpy(), want.numpy(), atol=1e-3, rtol=0)
    with pytest.raises(ValueError):
        mm.cosine_similarity_scores(x, torch.cat([t, t]), 1.0)


def test_config_c1_literal_random_init_vit_b32(mm, oracle):
    """BASELINE.json configs[0] on its literal inputs: the gallery is 10 000 image features of a random-init
    ViT-B/32 (transformers.CLIPModel(CLIPConfig()), torch.manual_seed(1) -- the reference's seed,
    code/search_image.py:323-324 -- on randn images), the 100 queries are text features of random token ids;
    rows unit-normalised; top-10 cosine in fp32 mode through BOTH fp32 paths (K1 CUDA cores, bf16x3 tensor cores)
    against the torch CPU oracle.  Random-init features are nearly collinear, so many scores are
    near-ties: a differing index is accepted only if the reference's own scores of the rows lie within 1e-5."""
    transformers = pytest.importorskip("transformers")
    torch.manual_seed(1)
    model = transformers.CLIPModel(transformers.CLIPConfig()).eval().cuda()
    gen = torch.Generator(device="cuda").manual_seed(1)
    feats = []
    with torch.no_grad():
        for _ in range(20):
            imgs = torch.randn((500, 3, 224, 224), generator=gen, device="cuda")
            f = model.get_image_features(pixel_values=imgs)
            f = f if isinstance(f, torch.Tensor) else f.pooler_output
            feats.append(f.float())
        ids = torch.randint(0, 49407, (100, 77), generator=gen, device="cuda")
        t = model.get_text_features(input_ids=ids)
        t = (t if isinstance(t, torch.Tensor) else t.pooler_output).float()
    g = torch.cat(feats)
    g = (g / g.norm(dim=-1, keepdim=True)).cpu()                 # build_cache, search_image.py:157
    q = t.cpu()
    del model, feats
    torch.cuda.empty_cache()
    assert g.shape == (10_000, 512) and q.shape == (100, 512) and torch.isfinite(g).all()
    want_v, want_i, ext_v, ext_i = oracle_topk_ext(oracle, q, g, 10, mode="fp32")
    gal = mm.DeviceGallery(g, mode="fp32")
    spread = float(ext_v[:, 0].mean() - ext_v[:, -1].mean())
    for path in ("gemv", "auto"):
        v, i = mm.search_topk(q, gal, 10, path=path)
        np.testing.assert_allclose(v.numpy(), want_v.numpy(), atol=1e-5, rtol=0)
        n_bad = explain_index_mismatches(i.numpy(), ext_i.numpy(), ext_v.numpy(), 1e-5)
        print(f"C1 literal ({path}): {n_bad} near-tie index differences of 1000; top-1..11 score spread {spread:.2e}")
        assert n_bad <= 100
    assert gal._split is not None                                # "auto" at 100 queries ran the tensor-core fp32 path
