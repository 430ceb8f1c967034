"""GPU parity tests (pytest -m gpu): every CUDA path through the C ABI against the CPU oracle on the same seeded
inputs, plus the committed golden fixtures.  Integer / byte / index outputs are compared bit-exactly."""
import os

import numpy as np
import pytest

from conftest import rand_tiles, rand_palettes
from tiler_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _u32(x):
    return np.asarray(x).view(np.uint32) if np.asarray(x).dtype == np.int32 else np.asarray(x)


# ---------------------------------------------------------------- features
def test_features_rgb_bit_exact(tm, oracle):
    tiles = np.concatenate([rand_tiles(3000, 1), rand_tiles(1000, 2, smooth=False),
                            np.zeros((1, 64), np.int32), np.full((1, 64), 0xFFFFFF, np.int32)])
    assert np.array_equal(tm.features_from_rgb(tiles), oracle.features_from_rgb(tiles))


def test_features_pal_bit_exact(tm, oracle):
    rng = np.random.default_rng(3)
    pal = rand_palettes(7, 16, 5)
    idx = rng.integers(0, 16, size=(2000, 64)).astype(np.uint8)
    tp = rng.integers(0, 7, size=2000).astype(np.int32)
    assert np.array_equal(tm.features_from_pal(idx, tp, pal), oracle.features_from_pal(idx, tp, pal))


def test_features_golden(tm):
    g = np.load(os.path.join(GOLD, "oracle_golden.npz"))
    assert np.array_equal(tm.features_from_rgb(g["tiles"]), g["feat_rgb"])
    assert np.array_equal(tm.features_from_pal(g["pal_idx_tk"], g["tile_pal"], g["palettes"]), g["feat_pal"])


def test_features_f64_tolerance(tm, oracle):
    tiles = rand_tiles(200, 4)
    for mode, lab in ((oracle.PVS_WEIGHTED_SPE_DCT, True), (oracle.PVS_WEIGHTED_DCT, False), (oracle.PVS_DCT, False),
                      (oracle.PVS_WAVELETS, False), (oracle.PVS_WAVELETS, True), (oracle.PVS_SPE_DCT, False)):
        got = tm.features_f64(tiles, mode, lab)
        want = np.stack([oracle.tile_features_f64(t, mode, lab) for t in tiles])
        if lab:   # pow() differs by ~1 ulp between libms: relative tolerance 1e-5 of the vector norm
            assert np.max(np.abs(got - want)) <= 1e-5 * np.abs(want).max()
        else:     # YUV path: identical operation order -> bit-exact
            assert np.array_equal(got, want)


def test_mirror_canonicalise(tm, oracle):
    tiles = rand_tiles(500, 8)
    got_tiles, flags = tm.mirror_canonicalise(tiles)
    for i in range(len(tiles)):
        h, v = oracle.mirror_heuristics(tiles[i])
        assert flags[i] == (int(h) | (int(v) << 1))
        t = tiles[i].reshape(8, 8)
        if h: t = t[:, ::-1]
        if v: t = t[::-1, :]
        assert np.array_equal(got_tiles[i].reshape(8, 8), t)


def test_distance_pairs(tm, oracle):
    a = synth.random_features(1000, 1, adversarial=True)
    b = synth.random_features(1000, 2, adversarial=True)
    got = _u32(tm.distance_pairs(a, b))
    want = np.array([oracle.compare_euclidean_dct(a[i], b[i]) for i in range(1000)], dtype=np.uint32)
    assert np.array_equal(got, want)


# ---------------------------------------------------------------- k-NN on tensor cores
@pytest.mark.parametrize("n_dict,n_q,adv", [(64, 7, False), (1000, 300, False), (1000, 300, True), (4133, 777, True)])
def test_knn_k1_exact(tm, oracle, n_dict, n_q, adv):
    d = synth.random_features(n_dict, 10 + n_dict, adv)
    q = synth.random_features(n_q, 20 + n_q, adv)
    knn = tm.KnnShort(d)
    idx, dist = knn.search(q, 1)
    oi, od = oracle.knn_short(d, q, 1)
    assert np.array_equal(_u32(dist), od)
    assert np.array_equal(idx, oi)
    knn.close()


@pytest.mark.parametrize("n_dict,n_q,k,adv", [(1000, 300, 64, False), (1000, 300, 64, True), (5000, 600, 8, True),
                                              (40, 50, 64, False), (9000, 513, 33, False), (30000, 700, 64, True),
                                              (20011, 300, 2, False), (70, 129, 64, True)])
def test_knn_topk_exact(tm, oracle, n_dict, n_q, k, adv):
    d = synth.random_features(n_dict, 30 + n_dict, adv)
    q = synth.random_features(n_q, 40 + n_q, adv)
    knn = tm.KnnShort(d)
    idx, dist = knn.search(q, k)
    oi, od = oracle.knn_short(d, q, k)
    assert np.array_equal(_u32(dist), od)
    assert np.array_equal(idx, oi)          # ordered by (distance, index) on both sides: ties resolve identically
    knn.close()


def test_knn_ties_and_duplicates(tm, oracle):
    # duplicated dictionary rows and queries equal to dictionary rows: zero distances and exact ties
    d = synth.random_features(300, 5)
    d = np.concatenate([d, d[:100], d[:50]])
    q = np.concatenate([d[:64], synth.random_features(64, 6)])
    knn = tm.KnnShort(d)
    for k in (1, 4, 64):
        idx, dist = knn.search(q, k)
        oi, od = oracle.knn_short(d, q, k)
        assert np.array_equal(_u32(dist), od) and np.array_equal(idx, oi)
    knn.close()


def test_knn_topk_heavy_ties_across_cuts(tm, oracle):
    # 40 distinct rows repeated 150 times each: every distance value occurs 150 times, so the streaming cuts and the
    # final selection all have to split ties by dictionary index, exactly like the oracle's (distance, index) order
    base = synth.random_features(40, 17)
    d = np.tile(base, (150, 1))
    q = np.concatenate([base[:20], synth.random_features(200, 18)])
    knn = tm.KnnShort(d)
    for k in (64, 7):
        idx, dist = knn.search(q, k)
        oi, od = oracle.knn_short(d, q, k)
        assert np.array_equal(_u32(dist), od) and np.array_equal(idx, oi)
    knn.close()


def test_knn_device_tensors(tm, oracle):
    import torch
    d = synth.random_features(2000, 7)
    q = synth.random_features(1000, 8)
    knn = tm.KnnShort(torch.from_numpy(d).cuda())
    idx, dist = knn.search(torch.from_numpy(q).cuda(), 64)
    torch.cuda.synchronize()
    oi, od = oracle.knn_short(d, q, 64)
    assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(_u32(dist.cpu().numpy()), od)
    knn.close()


def test_knn_full_dictionary_properties(tm):
    # BASELINE config-B dictionary size: properties that need no CPU brute force.  Queries that ARE dictionary rows
    # must find themselves at distance 0; distances must be non-decreasing; reported distances must equal the
    # pairwise kernel's value for the reported index.
    d = synth.random_features(65536, 99)
    sel = np.random.default_rng(1).choice(65536, size=4096, replace=False)
    q = d[sel]
    knn = tm.KnnShort(d)
    idx, dist = knn.search(q, 64)
    dist = _u32(dist)
    assert np.all(dist[:, 0] == 0)
    assert np.all(np.diff(dist.astype(np.int64), axis=1) >= 0)
    assert np.all((idx >= 0) & (idx < 65536))
    assert np.all([len(set(r)) == 64 for r in idx[:256]])
    chk = _u32(tm.distance_pairs(np.repeat(q, 64, axis=0), d[idx.reshape(-1)])).reshape(-1, 64)
    assert np.array_equal(chk, dist)
    i1, d1 = knn.search(q, 1)
    assert np.array_equal(_u32(d1)[:, 0], dist[:, 0])
    knn.close()


def test_knn_double(tm, oracle):
    rng = np.random.default_rng(3)
    d = rng.normal(0, 50, size=(700, 192))
    q = np.concatenate([rng.normal(0, 50, size=(300, 192)), d[:10]])
    idx, dist = tm.knn_double(d, q)
    oi, od = oracle.knn_double(d, q)
    assert np.array_equal(idx, oi) and np.array_equal(dist, od)


# ---------------------------------------------------------------- dithering
@pytest.mark.parametrize("use_tk", [True, False])
def test_dither_bit_exact(tm, oracle, use_tk):
    tiles = np.concatenate([rand_tiles(300, 11), rand_tiles(100, 12, smooth=False)])
    flags = np.random.default_rng(2).integers(0, 4, size=len(tiles)).astype(np.uint8)
    pal = rand_palettes(5, 16, 13, n_null=2)
    tp = (np.arange(len(tiles)) % 5).astype(np.int32)
    got = tm.dither(tiles, flags, tp, pal, use_thomas_knoll=use_tk)
    want = oracle.dither(tiles, flags, tp, pal, use_tk=use_tk)
    assert np.array_equal(got, want)


def test_dither_duplicate_luma_palette(tm, oracle):
    # colours with identical luma (299*dr + 587*dg + 114*db == 0) force the literal QuickSort-replay path
    pal = rand_palettes(2, 16, 21)
    r, g, b = 120, 100, 90
    for slot, (dr, dg, db) in ((3, (0, 0, 0)), (9, (-41, 17, 20)), (12, (-45, 27, -21))):
        pal[0, slot] = (r + dr) | ((g + dg) << 8) | ((b + db) << 16)
    luma = [299 * (c & 255) + 587 * ((c >> 8) & 255) + 114 * ((c >> 16) & 255) for c in pal[0, [3, 9, 12]]]
    assert len(set(luma)) == 1
    rng = np.random.default_rng(22)
    c = np.clip(np.array([r, g, b]) + rng.normal(0, 30, size=(200, 64, 3)), 0, 255).astype(np.int64)
    tiles = (c[..., 0] | (c[..., 1] << 8) | (c[..., 2] << 16)).astype(np.int32)
    tp = np.zeros(200, np.int32)
    for tk in (True, False):
        assert np.array_equal(tm.dither(tiles, None, tp, pal, use_thomas_knoll=tk), oracle.dither(tiles, None, tp, pal, use_tk=tk))


def test_dither_pairs_and_sizes(tm, oracle):
    tiles = rand_tiles(40, 31)
    for pal_size, n_pal in ((2, 3), (64, 2), (256, 2)):
        pal = rand_palettes(n_pal, pal_size, 32 + pal_size)
        pair_tile = np.repeat(np.arange(40), n_pal).astype(np.int32)
        pair_pal = np.tile(np.arange(n_pal), 40).astype(np.int32)
        got = tm.dither(tiles, None, pair_pal, pal, use_thomas_knoll=True, pair_tile=pair_tile)
        want = oracle.dither(tiles, None, pair_pal, pal, use_tk=True, pair_tile=pair_tile)
        assert np.array_equal(got, want)
    pal = rand_palettes(2, 16, 40)
    for y2 in (1, 2, 8, 16):
        tp = (np.arange(40) % 2).astype(np.int32)
        assert np.array_equal(tm.dither(tiles, None, tp, pal, use_thomas_knoll=False, y2_mixed_colors=y2),
                              oracle.dither(tiles, None, tp, pal, use_tk=False, y2_mixed_colors=y2))


def test_dither_golden(tm):
    g = np.load(os.path.join(GOLD, "oracle_golden.npz"))
    assert np.array_equal(tm.dither(g["tiles"], g["flags"], g["tile_pal"], g["palettes"], use_thomas_knoll=True), g["pal_idx_tk"])
    assert np.array_equal(tm.dither(g["tiles"], g["flags"], g["tile_pal"], g["palettes"], use_thomas_knoll=False), g["pal_idx_yl"])


# ---------------------------------------------------------------- k-means
def test_kmeans_f64_matches_oracle(tm, oracle):
    rng = np.random.default_rng(5)
    centres = rng.normal(0, 200, size=(24, 192))
    x = np.concatenate([c + rng.normal(0, 25, size=(60, 192)) for c in centres])
    rng.shuffle(x)
    init = x[:24].copy()
    labels, cent, inertia, iters = tm.kmeans_fit(x, 24, init=init)
    ol, oc, oin, oit = oracle.kmeans_lloyd(x, init)
    assert iters == oit
    assert np.array_equal(labels, ol)
    assert np.array_equal(cent, oc)                  # ordered segmented sums: bit-identical centroids
    assert abs(inertia - oin) <= 1e-9 * oin


def test_kmeans_seeded_init_matches_oracle(tm, oracle):
    rng = np.random.default_rng(6)
    x = rng.normal(0, 100, size=(1500, 16))
    init = oracle.kmeanspp_init(x, 12, 77)
    labels, cent, inertia, iters = tm.kmeans_fit(x, 12, init=None, seed=77)
    ol, oc, oin, oit = oracle.kmeans_lloyd(x, init)
    assert np.array_equal(labels, ol) and np.allclose(cent, oc, rtol=1e-12, atol=0)


def test_kmeans_empty_clusters_nan(tm, oracle):
    x = np.repeat(np.array([[0.0, 0.0], [10.0, 10.0]]), 50, axis=0)
    init = np.array([[0.0, 0.0], [10.0, 10.0], [1000.0, 1000.0]])
    labels, cent, _, _ = tm.kmeans_fit(x, 3, init=init, nan_empty=True)
    ol, oc, _, _ = oracle.kmeans_lloyd(x, init, nan_empty=True)
    assert np.array_equal(labels, ol)
    assert np.array_equal(np.isnan(cent), np.isnan(oc)) and np.array_equal(np.nan_to_num(cent), np.nan_to_num(oc))


def test_palquant_matches_oracle(tm, oracle):
    tiles = np.concatenate([rand_tiles(500, 50), rand_tiles(100, 51, smooth=False)])
    rng = np.random.default_rng(9)
    n_pal, pal_size = 6, 16
    tp = rng.integers(0, n_pal - 1, size=len(tiles)).astype(np.int32)   # palette 5 stays empty
    tp[:2] = 4                                                            # palette 4: 128 pixels
    got, iters = tm.palquant_kmeans(tiles, tp, n_pal, pal_size, seed=1234)
    for p in range(n_pal):
        px = tiles[tp == p].reshape(-1)
        want, n = oracle.quantize_palette(px, pal_size, seed=1234)
        assert np.array_equal(got[p], want), p


# ---------------------------------------------------------------- matcher
def _dictionary(oracle, n_dict, n_pal, pal_size, seed):
    tiles = rand_tiles(n_dict, seed)
    pal = rand_palettes(n_pal, pal_size, seed + 1)
    tp = (np.random.default_rng(seed + 2).integers(0, n_pal, size=n_dict)).astype(np.int32)
    idx = oracle.dither(tiles, None, tp, pal, use_tk=True)
    return tiles, pal, tp, idx


@pytest.mark.parametrize("extended", [True, False])
def test_matcher_matches_oracle(tm, oracle, extended):
    tiles, pal, tp, idx = _dictionary(oracle, 600, 6, 16, 60)
    q = rand_tiles(200, 66)
    m = tm.Matcher(idx, tp, pal, extended=extended)
    dict_feat = oracle.features_from_pal(idx, tp, pal)
    assert np.array_equal(m.dict_features(), dict_feat)
    t, p, e = m.match_rgb(q)
    qf = oracle.features_from_rgb(q)
    ot, op, oe = oracle.match_tiles(qf, dict_feat, idx, tp, pal, k=64, extended=extended)
    assert np.array_equal(_u32(e), oe)
    assert np.array_equal(t, ot) and np.array_equal(p, op)
    m.close()


def test_matcher_small_dictionary(tm, oracle):
    tiles, pal, tp, idx = _dictionary(oracle, 40, 3, 16, 70)     # fewer tiles than k = 64
    q = rand_tiles(64, 71)
    m = tm.Matcher(idx, tp, pal, extended=True)
    t, p, e = m.match_rgb(q)
    dict_feat = oracle.features_from_pal(idx, tp, pal)
    ot, op, oe = oracle.match_tiles(oracle.features_from_rgb(q), dict_feat, idx, tp, pal, k=64, extended=True)
    assert np.array_equal(t, ot) and np.array_equal(p, op) and np.array_equal(_u32(e), oe)
    m.close()


def test_matcher_host_batch_pipelined_upload(tm, oracle):
    """tm_match_tiles_rgb with a large HOST batch uploads it in pieces on a second stream while earlier pieces are matched
    (csrc/abi.cu); the result must equal the one-piece device-resident call, and the oracle on a sample of rows."""
    import torch
    tiles, pal, tp, idx = _dictionary(oracle, 300, 4, 16, 75)
    n = 8 * 148 * 128 + 1234                       # above the threshold, ragged last piece
    q = rand_tiles(4096, 76)
    q = np.ascontiguousarray(np.tile(q, (n // 4096 + 1, 1))[:n])
    q[:, 0] ^= (np.arange(n, dtype=np.int64) % 251).astype(q.dtype)   # rows differ between the pieces
    m = tm.Matcher(idx, tp, pal, extended=True)
    th, ph, eh = m.match_rgb(q, 16)                                   # host buffers: pipelined upload
    td, pd_, ed = m.match_rgb(torch.from_numpy(q).cuda(), 16)         # device-resident: one piece
    assert np.array_equal(th, td.cpu().numpy()) and np.array_equal(ph, pd_.cpu().numpy())
    assert np.array_equal(_u32(eh), _u32(ed.cpu().numpy()))
    rows = np.r_[0:64, n // 4 - 32:n // 4 + 32, n - 64:n]
    dict_feat = oracle.features_from_pal(idx, tp, pal)
    ot, op, oe = oracle.match_tiles(oracle.features_from_rgb(q[rows]), dict_feat, idx, tp, pal, k=16, extended=True)
    assert np.array_equal(th[rows], ot) and np.array_equal(ph[rows], op) and np.array_equal(_u32(eh)[rows], oe)
    m.close()


# ---------------------------------------------------------------- drop-in symbols
def test_dropin_ann_short(tm, oracle):
    d = synth.random_features(500, 80)
    q = synth.random_features(20, 81)
    tree = tm.AnnKdTreeShort(d)
    oi1, od1 = oracle.knn_short(d, q, 1)
    oi, od = oracle.knn_short(d, q, 64)
    for i in range(20):
        idx, err = tree.search(q[i])
        assert idx == oi1[i, 0] and err == od1[i, 0]
        idxs, errs = tree.search_multi(q[i], 64)
        assert np.array_equal(idxs, oi[i]) and np.array_equal(errs, od[i])
    tree.destroy()


def test_dropin_ann_double_yakmo_bico(tm, oracle):
    rng = np.random.default_rng(4)
    pts = rng.normal(0, 30, size=(256, 192))
    tree = tm.AnnKdTree(pts)
    q = rng.normal(0, 30, size=(16, 192))
    oi, od = oracle.knn_double(pts, q)
    for i in range(16):
        idx, err = tree.search(q[i])
        assert idx == oi[i] and err == od[i]
    tree.destroy()
    # yakmo: fixed point of Lloyd from the library's own seeding; check against the oracle's Lloyd from the same init
    x = np.concatenate([rng.normal(m, 2.0, size=(100, 8)) for m in (0.0, 30.0, 60.0, 90.0)])
    y = tm.Yakmo(4, 1, 300, 1, 0, 0, 0)
    y.load_train_data(x)
    labels = y.train_on_data()
    cent = y.get_centroids()
    assert sorted(np.bincount(labels, minlength=4).tolist()) == [100, 100, 100, 100]
    ol, oc, _, _ = oracle.kmeans_lloyd(x, cent)
    assert np.array_equal(ol, labels)                # returned centroids are a Lloyd fixed point with these labels
    y.destroy()
    b = tm.Bico(8, len(x), 16, 32, 16, 0x42381337)
    for row in x:
        b.insert_line(row, 1.0)
    c, w = b.get_results()
    assert 1 <= len(c) <= 16 and abs(w.sum() - len(x)) < 1e-6
    b.destroy()


# ---------------------------------------------------------------- k-means on int16 tile vectors (tensor-core search)
@pytest.mark.parametrize("n,k,adv", [(6000, 96, False), (6000, 96, True), (3000, 3, False), (2500, 700, True)])
def test_kmeans_i16_tensorcore_matches_oracle(tm, oracle, n, k, adv):
    rng = np.random.default_rng(n + k)
    centres = synth.random_features(k, 1000 + k, adv).astype(np.float64)
    x = np.clip(np.rint(centres[rng.integers(0, k, size=n)] + rng.normal(0, 25, size=(n, 192))), -32768, 32767).astype(np.int16)
    init = x[rng.permutation(n)[:k]].astype(np.float64)
    labels, cent, inertia, iters, amb = tm.kmeans_fit_i16(x, k, init, max_iter=12, nan_empty=True)
    ol, oc, oin, oit = oracle.kmeans_lloyd(x.astype(np.float64), init, max_iter=12, nan_empty=True)
    assert iters == oit
    assert np.array_equal(labels, ol)                      # exact f64 arg-min, certified or re-checked
    assert np.array_equal(np.isnan(cent), np.isnan(oc)) and np.array_equal(np.nan_to_num(cent), np.nan_to_num(oc))
    assert abs(inertia - oin) <= 1e-9 * max(oin, 1.0)
    assert amb <= 0.2 * n * (iters + 1)


def test_kmeans_i16_near_ties_fall_back(tm, oracle):
    # 80 centroids a fraction of an LSB apart: they all round to the same int16 vector, so neither the 4-candidate nor the
    # 64-candidate certificate can separate them and the brute-force f64 kernel must decide
    rng = np.random.default_rng(4)
    base = synth.random_features(1, 5)[0].astype(np.float64)
    init = np.stack([base + 0.005 * i for i in range(80)])
    x = np.clip(np.rint(base + rng.normal(0, 3, size=(2000, 192))), -32768, 32767).astype(np.int16)
    labels, cent, inertia, iters, amb = tm.kmeans_fit_i16(x, 80, init, max_iter=0)
    ol, oc, oin, oit = oracle.kmeans_lloyd(x.astype(np.float64), init, max_iter=0)
    assert np.array_equal(labels, ol) and amb > 0


def test_kmeans_update_blocked_order_large_clusters(tm, oracle):
    # clusters far above the 128-member block of the oracle's fixed summation order: centroids must still be bit-identical
    rng = np.random.default_rng(11)
    x = np.concatenate([rng.normal(m, 40.0, size=(1500, 192)) for m in (0.0, 300.0, -250.0)]) + rng.random((4500, 192)) * 1e-3
    init = x[[10, 1600, 3100]].copy()
    labels, cent, inertia, iters = tm.kmeans_fit(x, 3, init, max_iter=25)[:4]
    ol, oc, oin, oit = oracle.kmeans_lloyd(x, init, max_iter=25)
    assert iters == oit and np.array_equal(labels, ol) and np.array_equal(cent, oc)
    assert np.bincount(labels).max() > 1000


# ---------------------------------------------------------------- dlquant: against the REFERENCE's own C code (oracle/_ref)
def _dl_images():
    rng = np.random.default_rng(77)
    imgs = []
    a = rng.integers(0, 256, size=(64 * 64, 3), dtype=np.uint8); imgs.append(a)                       # noise: many cells
    b = rand_tiles(96, 5).reshape(-1); imgs.append(np.stack([b & 255, (b >> 8) & 255, (b >> 16) & 255], 1).astype(np.uint8))
    c = rng.integers(0, 256, size=(40, 3), dtype=np.uint8); imgs.append(np.repeat(c, 30, axis=0))      # few distinct colours
    g = np.load(os.path.join(GOLD, "dlquant_ref.npz"))["img"]; imgs.append(g)
    return imgs


@pytest.mark.parametrize("which", [3, 1])
def test_dlquant_matches_reference_c(tm, oracle, which):
    if oracle.ref_dlquant() is None:
        pytest.skip("oracle/_ref not built")
    imgs = _dl_images()
    name = "dl3quant" if which == 3 else "dl1quant"
    for quant_to, bpc in ((16, 5), (64, 5), (5, 4)):
        pal, cnt = tm.dlquant_batch(imgs, quant_to, bpc, which=which)
        for i, im in enumerate(imgs):
            rc, want = oracle.ref_dl3quant(im, len(im), 1, quant_to, bpc, which=name)
            assert rc == 0
            assert np.array_equal(pal[i], want), (name, i, quant_to, bpc)


def test_dlquant_golden_and_dropin(tm):
    g = np.load(os.path.join(GOLD, "dlquant_ref.npz"))
    rc, pal, full = tm.dlquant_dropin(g["img"], 48, 48, 16, 5, which=3)
    assert rc == 0 and np.array_equal(pal, g["dl3_16"]) and not full[:, 16:].any()
    rc, pal, full = tm.dlquant_dropin(g["img"], 48, 48, 16, 5, which=1)
    assert rc == 0 and np.array_equal(pal, g["dl1_16"])


# ---------------------------------------------------------------- motion search + Reconstruct
def _small_clip(w, h, n, seed):
    clip = synth.make_clip(w, h, n, cut_every=0, seed=seed, n_sprites=6)
    return synth.pack_rgb(clip)          # [n, h, w] int32


def test_sliding_features_bit_exact(tm, oracle):
    frame = _small_clip(96, 64, 1, 11)[0]
    got = tm.sliding_features(frame)
    want = oracle.sliding_features(frame)
    assert got.shape == ((64 - 7) * (96 - 7), 192)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("w,h,radius", [(96, 64, 32), (160, 96, 32), (64, 64, 5), (264, 136, 32), (264, 136, 17), (136, 72, 1)])
def test_motion_search_bit_exact(tm, oracle, w, h, radius):
    frames = _small_clip(w, h, 2, 12)
    tw, th = w // 8, h // 8
    tiles = synth.frame_to_tiles(frames[1])
    cur = oracle.features_from_rgb(tiles)
    dcts = oracle.sliding_features(frames[0])
    gx, gy, ge = tm.motion_search(cur, tw, th, dcts, radius)
    ox, oy, oe = oracle.motion_search(cur, tw, th, dcts, radius)
    assert np.array_equal(_u32(ge), oe)
    assert np.array_equal(gx, ox) and np.array_equal(gy, oy)
    # a static region must be found at (0, 0) with error 0 when the frame repeats
    gx0, gy0, ge0 = tm.motion_search(oracle.features_from_rgb(synth.frame_to_tiles(frames[0])), tw, th, dcts, radius)
    assert np.all(_u32(ge0) == 0) and np.all(gx0 == 0) and np.all(gy0 == 0)


def test_motion_search_persistent_ctas(tm, oracle, monkeypatch):
    """Frames with more 8x16 tile blocks than CTAs (1080p has 255): every CTA walks several blocks.  Forced here with 2 CTAs."""
    monkeypatch.setenv("TM_MOTION_CTAS", "2")
    w, h = 264, 136                                   # 33 x 17 tiles -> 3 x 3 tile blocks, 4-5 per CTA
    frames = _small_clip(w, h, 2, 19)
    cur = oracle.features_from_rgb(synth.frame_to_tiles(frames[1]))
    dcts = oracle.sliding_features(frames[0])
    gx, gy, ge = tm.motion_search(cur, w // 8, h // 8, dcts, 32)
    ox, oy, oe = oracle.motion_search(cur, w // 8, h // 8, dcts, 32)
    assert np.array_equal(_u32(ge), oe) and np.array_equal(gx, ox) and np.array_equal(gy, oy)


def test_predict_motion_frame_uses_natural_orientation(tm, oracle):
    frames = _small_clip(96, 64, 2, 13)
    tw, th = 12, 8
    canon, flags = tm.mirror_canonicalise(synth.frame_to_tiles(frames[1]))
    gx, gy, ge = tm.predict_motion_frame(frames[0], canon, flags, tw, th, 32)
    cur = oracle.features_from_rgb_mirrored(canon, flags)
    assert np.array_equal(cur, oracle.features_from_rgb(synth.frame_to_tiles(frames[1])))   # un-mirroring restores the source tile
    ox, oy, oe = oracle.motion_search(cur, tw, th, oracle.sliding_features(frames[0]), 32)
    assert np.array_equal(_u32(ge), oe) and np.array_equal(gx, ox) and np.array_equal(gy, oy)


def _build_small_dictionary(tm, oracle, frames, n_dict, n_pal, pal_size, seed):
    """Dictionary + palettes for a small clip through the library's own stages (their parity is tested above)."""
    tiles = np.concatenate([synth.frame_to_tiles(f) for f in frames])
    canon, flags = tm.mirror_canonicalise(tiles)
    sel = np.linspace(0, len(canon) - 1, n_dict).astype(np.int64)
    dt, dfl = np.ascontiguousarray(canon[sel]), np.ascontiguousarray(flags[sel])
    rng = np.random.default_rng(seed)
    dpal = rng.integers(0, n_pal, size=n_dict).astype(np.int32)
    pal, _ = tm.palquant_kmeans(dt, dpal, n_pal, pal_size, seed=seed)
    didx = tm.dither(dt, dfl, dpal, pal)
    return canon.reshape(len(frames), -1, 64), flags.reshape(len(frames), -1), didx, dpal, pal


@pytest.mark.parametrize("extended", [True, False])
def test_reconstruct_sequence_bit_exact(tm, oracle, extended):
    w, h, n_frames = 96, 64, 4
    frames = _small_clip(w, h, n_frames, 14)
    tw, th = w // 8, h // 8
    canon, flags, didx, dpal, pal = _build_small_dictionary(tm, oracle, frames, 160, 4, 16, 5)
    m = tm.Matcher(didx, dpal, pal, extended=extended)
    got = m.reconstruct_sequence(canon, flags, tw, th, radius=32)
    dict_feat = oracle.features_from_pal(didx, dpal, pal)
    want = oracle.reconstruct_sequence(canon, flags, tw, th, dict_feat, didx, dpal, pal, radius=32, extended=extended)
    for key in ("is_pred", "pred_x", "pred_y", "tile_idx", "pal_idx"):
        assert np.array_equal(got[key], want[key]), key
    assert np.array_equal(_u32(got["err"]), want["err"])
    assert np.array_equal(got["recon"], want["recon"])
    assert got["is_pred"][0].sum() == 0            # first frame of the sequence: no motion prediction (:1496)
    assert got["is_pred"][1:].sum() > 0            # a translating clip must predict some tiles
    psnr = np.array([oracle.euclidean_to_psnr(int(e)) for e in want["err"].reshape(-1)], dtype=np.float32)
    assert np.allclose(got["psnr"].reshape(-1), psnr, rtol=1e-6, atol=1e-5)
    # decoded-frame PSNR of the reconstruction against the source, GPU reduction vs numpy
    src = np.stack([synth.tiles_to_frame(synth.frame_to_tiles(f), h, w) for f in frames])
    a, b = src.astype(np.int64), want["recon"].astype(np.int64)
    se = sum((((a >> s) & 255) - ((b >> s) & 255)) ** 2 for s in (0, 8, 16)).sum()
    assert abs(tm.mse_rgb(src, got["recon"]) - se / (3.0 * a.size)) < 1e-9
    m.close()


def test_reconstruct_frame_by_frame_equals_sequence(tm, oracle):
    """tm_reconstruct_frame chained by the host (the Pascal host's own frame loop) = tm_reconstruct_sequence."""
    w, h, n_frames = 96, 64, 4
    frames = _small_clip(w, h, n_frames, 16)
    canon, flags, didx, dpal, pal = _build_small_dictionary(tm, oracle, frames, 160, 4, 16, 7)
    m = tm.Matcher(didx, dpal, pal, extended=True)
    seq = m.reconstruct_sequence(canon, flags, 12, 8, radius=32)
    back = None
    for f in range(n_frames):
        r = m.reconstruct_frame(canon[f], flags[f], 12, 8, back=back, radius=32)
        for key in ("is_pred", "pred_x", "pred_y", "tile_idx", "pal_idx", "err", "recon"):
            assert np.array_equal(r[key], seq[key][f]), (f, key)
        back = r["recon"]
    m.close()


def test_reconstruct_sequence_device_tensors(tm, oracle):
    import torch
    w, h, n_frames = 64, 64, 3
    frames = _small_clip(w, h, n_frames, 15)
    canon, flags, didx, dpal, pal = _build_small_dictionary(tm, oracle, frames, 100, 2, 16, 6)
    m = tm.Matcher(didx, dpal, pal, extended=True)
    host = m.reconstruct_sequence(canon, flags, 8, 8)
    dev = m.reconstruct_sequence(torch.from_numpy(canon).cuda(), torch.from_numpy(flags).cuda(), 8, 8)
    for key in ("is_pred", "pred_x", "pred_y", "tile_idx", "pal_idx", "recon"):
        assert np.array_equal(dev[key].cpu().numpy(), host[key]), key
    m.close()


# ---------------------------------------------------------------- Reduce and the whole encode
def test_tile_classes_exact(tm):
    rng = np.random.default_rng(9)
    base = rand_tiles(5000, 31)
    tiles = base[rng.integers(0, 5000, size=60000)]            # heavy duplication
    tiles[::7, 63] ^= 1                                         # near-duplicates differing in the last pixel only
    cls, n_cls = tm.tile_classes(tiles)
    rows = np.ascontiguousarray(tiles).view(np.dtype((np.void, 256))).reshape(-1)
    uniq, inv = np.unique(rows, return_inverse=True)
    assert n_cls == len(uniq)
    # same partition: class ids are a relabelling of numpy's
    pairs = np.unique(np.stack([cls, inv.reshape(-1)], axis=1), axis=0)
    assert len(pairs) == n_cls and len(np.unique(pairs[:, 0])) == n_cls and len(np.unique(pairs[:, 1])) == n_cls
    assert cls.min() == 0 and cls.max() == n_cls - 1


def test_encode_end_to_end_small_clip(tm, oracle):
    from tiler_b200 import gtm
    from tiler_b200.encoder import TilingEncoder, euclidean_to_psnr, psnr_rgb
    w, h, n = 96, 64, 6
    frames = synth.pack_rgb(synth.make_clip(w, h, n, cut_every=3, seed=33, n_sprites=4))
    tw, th = w // 8, h // 8
    seqs = [(0, 2), (3, 5)]
    enc = TilingEncoder(palette_size=16, palette_count=2)
    res = enc.encode(frames, seqs, tile_count=150, radius=32)
    # --- PredictMotion + Reduce against a numpy/oracle restatement
    tiles = np.stack([synth.frame_to_tiles(f) for f in frames])
    canon, flags = tm.mirror_canonicalise(tiles.reshape(-1, 64))
    canon, flags = canon.reshape(n, -1, 64), flags.reshape(n, -1)
    psnr = np.empty((n, tw * th), np.float32)
    for f in range(n):
        prev = frames[f - 1] if f > 0 else frames[1]
        cur = oracle.features_from_rgb_mirrored(canon[f], flags[f])
        _, _, e = oracle.motion_search(cur, tw, th, oracle.sliding_features(prev), 32)
        psnr[f] = [oracle.euclidean_to_psnr(int(v)) for v in e]
    got_psnr, _, _ = enc.predict_motion(frames, canon, flags, tw, th, 32)
    assert np.array_equal(got_psnr, psnr)
    eff = psnr.astype(np.float64); eff[0] /= 10.0; eff[3] /= 10.0                # STCGREval: Single promoted to Double (:4028-4031)
    unpred = ~(eff.reshape(-1) > enc.reduce_threshold)
    rows = np.ascontiguousarray(canon.reshape(-1, 64)).view(np.dtype((np.void, 256))).reshape(-1)
    assert len(enc.tiles) == len(np.unique(rows[unpred]))                       # MakeTilesUnique(True) count at the chosen threshold
    assert abs(len(enc.tiles) - 150) <= 12                                      # the search lands near the requested tile count
    uc = enc.use_count
    assert (np.diff(uc) <= 0).all() and uc.sum() == unpred.sum()                # ReindexTiles(True): use count descending
    # --- the stream decodes to exactly the frames Reconstruct drew, and they resemble the source
    decoded, hdr = gtm.decode_gtm(res["gtm"])
    assert hdr["frame_count"] == n and hdr["kf_count"] == 2 and (hdr["width"], hdr["height"]) == (w, h)
    assert np.array_equal(decoded, res["recon"])
    assert psnr_rgb(decoded, frames) > 18.0
    assert abs(psnr_rgb(decoded, frames) - 10 * np.log10(255.0 ** 2 / tm.mse_rgb(frames, decoded))) < 1e-9
    tmap = res["tilemap"]
    assert not tmap["is_pred"][0].any() and not tmap["is_pred"][3].any()        # first frame of each keyframe sequence
    assert tmap["is_pred"][[1, 2, 4, 5]].any()
    used = tmap["tile_idx"][tmap["tile_idx"] >= 0]
    assert used.max() == len(res["tiles"]) - 1 and np.array_equal(np.bincount(used, minlength=len(res["tiles"])), res["use_count"])


def test_knn_full_range_int16_wrapping(tm, oracle):
    """The reference accumulates the distance in a Cardinal: it wraps mod 2^32.  Full-range int16 vectors (norms ~ 7e10) and
    single extreme rows among small ones must give the oracle's wrapped distances and indices for k = 1, 4 and 64."""
    rng = np.random.default_rng(77)
    d_small = synth.random_features(3000, 5)
    q_small = synth.random_features(300, 6)
    d_big = rng.integers(-32768, 32768, size=(3000, 192)).astype(np.int16)
    q_big = rng.integers(-32768, 32768, size=(300, 192)).astype(np.int16)
    d_small = np.clip(d_small, -1500, 1500).astype(np.int16); q_small = np.clip(q_small, -1500, 1500).astype(np.int16)   # norms < 2^29
    d_one = d_small.copy(); d_one[1234] = 32767          # one row with norm 192 * 32767^2 > 2^29
    q_one = q_small.copy(); q_one[7] = -32768
    for d, q in ((d_big, q_big), (d_one, q_small), (d_small, q_one), (d_small, q_small)):
        knn = tm.KnnShort(d)
        for k in (1, 4, 64):
            idx, dist = knn.search(q, k)
            oi, od = oracle.knn_short(d, q, k)
            assert np.array_equal(_u32(dist), od), k
            assert np.array_equal(idx, oi), k
        knn.close()


def test_encode_matches_oracle_pipeline(tm, oracle):
    """The north star's end-to-end check: the whole encode on the GPU against the same pipeline restated on the CPU
    (oracle/pipeline.py: host bookkeeping from the reference, compute from oracle/tm_oracle.c; it imports nothing from the
    product) on the same synthetic clip, two keyframe sequences.  Stated tolerance: decoded-frame PSNR within 0.05 dB; the
    stages are deterministic and bit-exact, so dictionary, palettes, tilemap and frames are compared for equality."""
    from oracle import pipeline as P
    from tiler_b200 import gtm
    from tiler_b200.encoder import TilingEncoder, psnr_rgb
    w, h, n = 96, 64, 6
    frames = synth.pack_rgb(synth.make_clip(w, h, n, cut_every=3, seed=41, n_sprites=4))
    seqs = [(0, 2), (3, 5)]
    enc = TilingEncoder(palette_size=16, palette_count=2, seed=0x42381337)
    res = enc.encode(frames, seqs, tile_count=150, radius=32)
    ref = P.encode(frames, seqs, 150, 2, 16, 0x42381337, radius=32)
    assert enc.reduce_threshold == ref["threshold"]
    assert np.array_equal(res["palettes"], ref["palettes"])
    assert np.array_equal(res["tiles"], ref["tiles"]) and np.array_equal(res["use_count"], ref["use_count"])
    for key in ("tile_idx", "pal_idx", "pred_x", "pred_y", "is_pred"):
        assert np.array_equal(np.asarray(res["tilemap"][key]).reshape(n, -1), np.asarray(ref[key]).reshape(n, -1)), key
    got_frames, _ = gtm.decode_gtm(res["gtm"])
    assert np.array_equal(got_frames, ref["recon"]) and np.array_equal(res["recon"], ref["recon"])
    assert abs(psnr_rgb(got_frames, frames) - psnr_rgb(ref["recon"], frames)) <= 0.05
