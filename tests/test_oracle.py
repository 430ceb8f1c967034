"""CPU tests: the oracle against everything in the reference that pins this path:
the Test properties (tilingencoder.pas:3847-3902), the literal tables (utils.pas:47-109), the repo's QuickSort
(extern.pas:370-418), and the reference's own dlquant C code built into oracle/_ref."""
import os

import numpy as np
import pytest

from conftest import rand_tiles, rand_palettes

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_tables_match_reference_literals(oracle):
    dm = oracle.dithering_map()
    assert dm[:8].tolist() == [0, 48, 12, 60, 3, 51, 15, 63] and dm[-8:].tolist() == [42, 26, 38, 22, 41, 25, 37, 21]
    assert sorted(dm.tolist()) == list(range(64))
    sn = oracle.dct_snake()
    assert sn[:8].tolist() == [0, 1, 5, 6, 14, 15, 27, 28] and sorted(sn.tolist()) == list(range(64))
    w = oracle.dct_weights()
    assert w[0, 0, 0] == 1.6193873005 and w[2, 7, 7] == 0.285345396658 and np.allclose(w[0], w[0].T)
    vi = oracle.vec_inv()
    assert vi[0] == 0 and vi[4] == 65536 and vi[12] == 65536 // 3 and vi[1023] == 65536 // 255


def test_reference_selftest_colour_roundtrips(oracle):
    # TTilingEncoder.Test: 10001 random colours survive RGB->LAB->RGB and RGB->YUV->RGB (tilingencoder.pas:3857-3867)
    rng = np.random.default_rng(7)
    for c in rng.integers(0, (1 << 24) - 1, size=10001):
        c = int(c)
        r, g, b = c & 255, (c >> 8) & 255, (c >> 16) & 255
        assert oracle.lab_to_rgb(*oracle.rgb_to_lab(r, g, b)) == c
        assert oracle.yuv_to_rgb(*oracle.rgb_to_yuv(r, g, b)) == c


def test_reference_selftest_dct_roundtrip(oracle):
    # T[i,j] = ToRGB(i*8, j*32, i*j): DCT -> inverse DCT reproduces the tile bit-exactly (tilingencoder.pas:3872-3893)
    T = np.array([[((i * j) & 255) << 16 | ((j * 32) & 255) << 8 | ((i * 8) & 255) for j in range(8)] for i in range(8)],
                 dtype=np.int32).reshape(64)
    for mode in (oracle.PVS_DCT, oracle.PVS_WEIGHTED_DCT):
        d = oracle.tile_features_f64(T, mode, False)
        assert np.array_equal(oracle.inv_tile_features_f64(d, mode, False), T)


def test_features_basic_properties(oracle):
    white = np.full((1, 64), 0xFFFFFF, dtype=np.int32)
    f = oracle.features_from_rgb(white)[0]
    assert f[0] == 13214 and np.count_nonzero(f) == 1          # Y-DC = 64*255*0.5*1.6193873
    tiles = rand_tiles(64, 3)
    f = oracle.features_from_rgb(tiles)
    # H-mirror negates odd-u coefficients, V-mirror odd-v; check through the mirrored-read path
    sn = oracle.dct_snake()
    for t in range(8):
        fm = oracle.tile_features_i16(rgb=tiles[t], hmirror=True)
        sign = np.ones(192)
        for c in range(3):
            for v in range(8):
                for u in range(8):
                    if u & 1:
                        sign[c * 64 + sn[v * 8 + u]] = -1
        assert np.max(np.abs(fm - sign * f[t])) <= 1


def test_distance_matches_definition(oracle):
    rng = np.random.default_rng(5)
    a = rng.integers(-13000, 13000, size=(50, 192)).astype(np.int16)
    b = rng.integers(-13000, 13000, size=(50, 192)).astype(np.int16)
    for i in range(50):
        want = int(((a[i].astype(np.int64) - b[i]) ** 2).sum() % (1 << 32))
        assert oracle.compare_euclidean_dct(a[i], b[i]) == want
        if np.abs(a[i].astype(np.int64) - b[i]).max() < 32768:
            assert oracle.compare_euclidean_dct(a[i], b[i], sse=True) == want
    assert abs(oracle.euclidean_to_psnr(0) - 51.1411) < 1e-3   # cPsnrMaxValue (utils.pas:111)


def test_quicksort_restatement(oracle):
    # sorts correctly; equal keys keep the multiset; pinned sequence for a duplicate-key case
    rng = np.random.default_rng(11)
    key = rng.integers(0, 1000, size=16).astype(np.int32)
    for _ in range(200):
        d = rng.integers(0, 16, size=64).astype(np.uint8)
        s = oracle.quicksort_bytes_by_key(d, key)
        assert np.all(np.diff(key[s]) >= 0) and sorted(s.tolist()) == sorted(d.tolist())
    key2 = np.array([5, 5, 1, 9], dtype=np.int32)
    d = np.array([0, 1, 2, 3, 1, 0, 3, 2, 0, 1], dtype=np.uint8)
    s = oracle.quicksort_bytes_by_key(d, key2)
    assert np.all(np.diff(key2[s]) >= 0)


def test_knn_oracle_vs_numpy(oracle):
    from tiler_b200 import synth
    d = synth.random_features(500, 1)
    q = synth.random_features(40, 2)
    idx, dist = oracle.knn_short(d, q, 8)
    full = ((q[:, None, :].astype(np.int64) - d[None, :, :]) ** 2).sum(-1)
    order = np.lexsort((np.broadcast_to(np.arange(500), full.shape), full), axis=-1)[:, :8]
    assert np.array_equal(idx, order)
    assert np.array_equal(dist, np.take_along_axis(full, order, 1).astype(np.uint32))
    i2, d2 = oracle.knn_short(d, q, 8, sse=True)
    assert np.array_equal(i2, idx) and np.array_equal(d2, dist)


def test_dither_oracle_sanity(oracle):
    tiles = rand_tiles(6, 9)
    pal = rand_palettes(2, 16, 4, n_null=3)
    for tk in (True, False):
        out = oracle.dither(tiles, None, np.array([0, 1, 0, 1, 0, 1], np.int32), pal, use_tk=tk)
        assert out.max() < 13                                    # null colours are never chosen (Remap)
    # a tile of one palette colour dithers to that colour everywhere
    flat = np.full((1, 64), pal[0, 5], dtype=np.int32)
    out = oracle.dither(flat, None, np.array([0], np.int32), pal, use_tk=True)
    assert np.all(pal[0][out[0]] == pal[0, 5])


def test_kmeans_and_palette_oracle(oracle):
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.normal(m, 1.0, size=(200, 5)) for m in (0.0, 10.0, 20.0)])
    init = x[[0, 200, 400]].copy()
    labels, cent, inertia, it = oracle.kmeans_lloyd(x, init)
    assert sorted(np.bincount(labels).tolist()) == [200, 200, 200] and it >= 1
    px = rand_tiles(40, 6).reshape(-1)
    pal, n = oracle.quantize_palette(px, 16, seed=3)
    assert n == 16 and np.all(pal != oracle.NULL_COLOR)
    hsv = [oracle.rgb_to_hsv(int(c)) for c in pal]
    keys = [(v, s, h) for h, s, v in hsv]
    assert keys == sorted(keys)                                   # sorted by (Val, Sat, Hue), utils.pas:741-748
    pal2, n2 = oracle.quantize_palette(px[:5], 16, seed=3)
    assert n2 == 5 and np.all(pal2[5:] == oracle.NULL_COLOR)


def test_reference_dlquant_builds_and_runs(oracle):
    if oracle.ref_dlquant() is None:
        pytest.skip("oracle/_ref not built (reference checkout absent at build time)")
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, size=(32 * 32, 3), dtype=np.uint8)
    rc, pal = oracle.ref_dl3quant(img, 32, 32, 16)
    assert rc == 0 and pal.shape == (16, 3)
    g = np.load(os.path.join(GOLD, "dlquant_ref.npz"))
    rc, pal = oracle.ref_dl3quant(g["img"], 48, 48, 16)
    assert np.array_equal(pal, g["dl3_16"])
    rc, pal = oracle.ref_dl3quant(g["img"], 48, 48, 16, which="dl1quant")
    assert np.array_equal(pal, g["dl1_16"])


def test_oracle_against_committed_golden(oracle):
    g = np.load(os.path.join(GOLD, "oracle_golden.npz"))
    assert np.array_equal(oracle.features_from_rgb(g["tiles"]), g["feat_rgb"])
    assert np.array_equal(oracle.features_from_pal(g["pal_idx_tk"], g["tile_pal"], g["palettes"]), g["feat_pal"])
    assert np.array_equal(oracle.dither(g["tiles"], g["flags"], g["tile_pal"], g["palettes"], use_tk=True), g["pal_idx_tk"])
    assert np.array_equal(oracle.dither(g["tiles"], g["flags"], g["tile_pal"], g["palettes"], use_tk=False), g["pal_idx_yl"])
    idx, dist = oracle.knn_short(g["feat_pal"], g["feat_rgb"], 8)
    assert np.array_equal(idx, g["knn_idx"]) and np.array_equal(dist, g["knn_dist"])


# ---------------------------------------------------------------- motion search + Reconstruct restatement
def test_motion_search_properties(oracle):
    from tiler_b200 import synth
    rng = np.random.default_rng(3)
    w, h = 64, 48
    base = rng.integers(0, 256, size=(h + 16, w + 16, 3)).astype(np.uint8)
    prev = synth.pack_rgb(base[8:8 + h, 8:8 + w])
    cur = synth.pack_rgb(base[8 + 2:8 + 2 + h, 8 - 3:8 - 3 + w])   # cur(y, x) = prev(y + 2, x - 3)
    dcts = oracle.sliding_features(prev)
    assert dcts.shape == ((h - 7) * (w - 7), 192)
    # sliding features at tile-aligned offsets are the tile features
    tiles_prev = synth.frame_to_tiles(prev)
    f_prev = oracle.features_from_rgb(tiles_prev)
    pw = w - 7
    for t in (0, 5, 13):
        ty, tx = divmod(t, w // 8)
        assert np.array_equal(dcts[ty * 8 * pw + tx * 8], f_prev[t])
    px, py, err = oracle.motion_search(oracle.features_from_rgb(synth.frame_to_tiles(cur)), w // 8, h // 8, dcts, 32)
    inner = [t for t in range((w // 8) * (h // 8)) if 1 <= t % (w // 8) < w // 8 - 1 and 1 <= t // (w // 8) < h // 8 - 1]
    assert all(px[t] == -3 and py[t] == 2 and err[t] == 5 for t in inner)   # exact match, error = Manhattan penalty only
    # identical frame: zero vector, zero error everywhere
    px0, py0, err0 = oracle.motion_search(f_prev, w // 8, h // 8, dcts, 32)
    assert not px0.any() and not py0.any() and not err0.any()
    # radius 1 -> window dy-1 .. dy+0 (tilingencoder.pas:1213-1216 after Dec(ARadius))
    px1, py1, _ = oracle.motion_search(oracle.features_from_rgb(synth.frame_to_tiles(cur)), w // 8, h // 8, dcts, 1)
    assert px1.min() >= -1 and px1.max() <= 0 and py1.min() >= -1 and py1.max() <= 0


def test_reconstruct_sequence_oracle_semantics(oracle):
    from tiler_b200 import synth
    frames = synth.pack_rgb(synth.make_clip(64, 48, 3, seed=21, n_sprites=3, noise=0.0))
    tw, th = 8, 6
    tiles = np.stack([synth.frame_to_tiles(f) for f in frames])
    flags = np.zeros(tiles.shape[:2], np.uint8)
    # dictionary = 80 tiles of frame 0 "dithered" to 16 grey levels
    pal = np.array([[(v * 17) * 0x010101 for v in range(16)]], dtype=np.int32)
    sel = tiles[0][:48]
    luma = ((sel & 255) * 299 + ((sel >> 8) & 255) * 587 + ((sel >> 16) & 255) * 114) // 1000
    didx = np.clip((luma + 8) // 17, 0, 15).astype(np.uint8)
    dpal = np.zeros(len(didx), np.int32)
    dfeat = oracle.features_from_pal(didx, dpal, pal)
    r = oracle.reconstruct_sequence(tiles, flags, tw, th, dfeat, didx, dpal, pal, radius=32, extended=True)
    assert not r["is_pred"][0].any() and (r["tile_idx"][0] >= 0).all()        # start frame: k-NN only (:1496)
    # every reconstructed pixel of frame 0 is a palette colour; predicted tiles of later frames copy the previous recon
    assert np.isin(r["recon"][0], pal).all()
    for f in (1, 2):
        for t in np.nonzero(r["is_pred"][f])[0][:10]:
            y, x = (t // tw) * 8, (t % tw) * 8
            sy, sx = y + r["pred_y"][f, t], x + r["pred_x"][f, t]
            assert np.array_equal(r["recon"][f, y:y + 8, x:x + 8], r["recon"][f - 1, sy:sy + 8, sx:sx + 8])
        dead = r["is_pred"][f].astype(bool) & (r["tile_idx"][f] < 0)
        assert (r["err"][f][dead] <= 192).all()                                # dead band: mpErr <= cTileDCTSize (:1534)
