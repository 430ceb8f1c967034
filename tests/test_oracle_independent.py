"""Second, independent derivations that pin the C oracle (oracle/tm_oracle.c): plain float64 numpy / pure Python written from
the reference's formulas, not from the oracle's code -- a transcription error shared by the oracle and the CUDA kernels
(which were written side by side) would show up here.  CPU only."""
import math

import numpy as np

from conftest import rand_tiles, rand_palettes

NULL = np.int32(-65281)


def _numpy_features_f64(rgb):
    """ConvertToCpnPixels + ComputeCpnPixelsPsyVisFeatures from the formulas alone (utils.pas:478-490, :100-109,
    tilingencoder.pas:1709, :3103-3131), everything in float64: no f32 storage, no asm summation order.  -> unrounded."""
    from oracle import oracle as O
    rgb = np.asarray(rgb, dtype=np.int64).reshape(-1, 8, 8)
    r, g, b = (rgb & 255).astype(np.float64), ((rgb >> 8) & 255).astype(np.float64), ((rgb >> 16) & 255).astype(np.float64)
    y = r * 0.299 + g * 0.587 + b * 0.114
    planes = np.stack([y, (b - y) * 0.492, (r - y) * 0.877], axis=1)            # [n, 3, y, x]
    k = np.arange(8)
    c = np.cos((k[None, :] + 0.5) * k[:, None] * math.pi / 8.0)                   # c[u, x]
    ratio = np.ones((8, 8)); ratio[0, :] = ratio[:, 0] = math.sqrt(0.5); ratio[0, 0] = 0.5   # cDCTUVRatio[v, u]
    z = np.einsum("ncyx,vy,ux->ncvu", planes, c, c) * ratio[None, None] * O.dct_weights()[None]
    snake = O.dct_snake().reshape(8, 8)                                          # position of (v, u) in the zig-zag
    out = np.empty((rgb.shape[0], 192))
    for cpn in range(3):
        out[:, cpn * 64 + snake.reshape(-1)] = z[:, cpn].reshape(-1, 64)
    return out


def test_features_i16_against_float64_numpy_dct(oracle):
    tiles = np.concatenate([rand_tiles(700, 901), rand_tiles(300, 902, smooth=False)])
    got = oracle.features_from_rgb(tiles).astype(np.int64)
    ref = _numpy_features_f64(tiles)
    d = np.abs(got - np.rint(ref))
    assert d.max() <= 1                                                          # +-1 LSB: the asm's f32 products and pair adds
    assert (d != 0).mean() <= 1e-3
    assert np.abs(got - ref).max() <= 0.5 + 2e-2                                 # and never further than the rounding plus f32 noise
    # Y DC of a white tile: 64 * 255 * 0.5 * 1.6193873005 (no 2/N normalisation)
    white = oracle.features_from_rgb(np.full((1, 64), 0xFFFFFF, np.int32))[0]
    assert white[0] == round(64 * 255 * 0.5 * 1.6193873005) and not white[1:64].any()


def _tk_plan_python(pal, col):
    """DeviseBestMixingPlanThomasKnoll + PreparePlan + ColorCompare restated in plain Python integers
    (tilingencoder.pas:2268-2337, 2565-2612): returns the 64 picks (compacted indices) BEFORE the luma sort."""
    def tdiv(a, b):                                  # Pascal div truncates toward zero
        q = abs(a) // abs(b)
        return q if (a >= 0) == (b > 0) else -q
    cols = [(int(c) & 255, (int(c) >> 8) & 255, (int(c) >> 16) & 255) for c in pal if int(c) != int(NULL)]
    remap = [i for i, c in enumerate(pal) if int(c) != int(NULL)]
    luma = [r * 299 + g * 587 + b * 114 for r, g, b in cols]
    s = [int(col) & 255, (int(col) >> 8) & 255, (int(col) >> 16) & 255]
    e = [0, 0, 0]
    picks = []
    for c in range(64):
        t = [s[i] + tdiv(e[i] * 9, 100) for i in range(3)]
        least, chosen = None, c % len(cols)
        for idx, (r, g, b) in enumerate(cols):
            l1 = t[0] * 299 + t[1] * 587 + t[2] * 114
            ld = tdiv(l1 - luma[idx], 1000)
            pen = 13 * ((t[0] - r) ** 2 + (t[1] - g) ** 2 + (t[2] - b) ** 2) + ((ld * ld) << 5)
            if least is None or pen < least:
                least, chosen = pen, idx
        picks.append(chosen)
        for i in range(3):
            e[i] += s[i] - cols[chosen][i]
    return picks, luma, remap


def test_thomas_knoll_against_pure_python_plan(oracle):
    """One tile (64 pixels), a 16-colour palette with DISTINCT lumas and two null entries: with distinct lumas the sorted
    pick list is unique, so the comparison does not depend on the reference's non-stable QuickSort."""
    rng = np.random.default_rng(5150)
    while True:
        pal = rand_palettes(1, 16, int(rng.integers(1 << 30)), n_null=2)[0]
        cols = [(int(c) & 255, (int(c) >> 8) & 255, (int(c) >> 16) & 255) for c in pal[:14]]
        if len({r * 299 + g * 587 + b * 114 for r, g, b in cols}) == 14:
            break
    tile = rand_tiles(1, 77)[0]
    dmap = oracle.dithering_map()
    want = np.empty(64, np.uint8)
    for p in range(64):
        picks, luma, remap = _tk_plan_python(pal, tile[p])
        picks.sort(key=lambda i: luma[i])
        want[p] = remap[picks[dmap[p]]]                 # cDitheringMap[((y and 7) shl 3) or (x and 7)] = position p for an 8x8 tile
    got = oracle.dither(tile[None], np.zeros(1, np.uint8), np.zeros(1, np.int32), pal[None], use_tk=True)[0]
    assert np.array_equal(got, want)
    # mirrored tile: dithered in natural orientation, stored mirrored (DitherTile, :2696-2697, :2721-2722)
    nat = tile.reshape(8, 8)
    stored = nat[::-1, ::-1].reshape(64)               # H and V mirrored storage of the same natural tile
    got_m = oracle.dither(stored[None], np.full(1, 3, np.uint8), np.zeros(1, np.int32), pal[None], use_tk=True)[0]
    assert np.array_equal(got_m.reshape(8, 8)[::-1, ::-1].reshape(64), want)


def test_pipeline_quicksort_is_the_reference_procedure(oracle):
    """oracle/pipeline.py's generic QuickSort and the C oracle's byte QuickSort are two restatements of extern.pas:370-418;
    tiler_b200.encoder._quicksort_desc is a third (host bookkeeping of the product).  All must agree on tie placement."""
    from oracle import pipeline as P
    from tiler_b200.encoder import _quicksort_desc
    rng = np.random.default_rng(12)
    for n in (1, 2, 3, 17, 64, 200):
        key = rng.integers(0, 6, size=256).astype(np.int32)          # few distinct keys: many ties
        data = rng.integers(0, 256, size=n).astype(np.uint8)
        want = oracle.quicksort_bytes_by_key(data, key)
        items = [int(v) for v in data]
        P.quicksort(items, lambda a, b: int(key[a] > key[b]) - int(key[a] < key[b]))
        assert items == [int(v) for v in want]
        counts = rng.integers(0, 5, size=n)
        recs = [[int(c), i] for i, c in enumerate(counts)]
        P.quicksort(recs, lambda a, b: (b[0] > a[0]) - (b[0] < a[0]))  # ComparePaletteUseCount: descending
        assert [r[1] for r in recs] == list(_quicksort_desc(counts))


def test_pipeline_golden_ratio_search_and_host_search_agree():
    from oracle import pipeline as P
    from tiler_b200.encoder import golden_ratio_search
    vals = np.sort(np.random.default_rng(3).uniform(0, 51, size=5000))
    f = lambda x: float(np.searchsorted(vals, x, side="right"))
    for target in (1.0, 250.0, 2500.0, 4999.0, 6000.0):
        seen = []
        x_ref = P.golden_ratio_search(lambda x: (seen.append(x), f(x))[1], 0.0, P.C_PSNR_MAX, target, 1e-6, 0.5)
        x_got, x_last = golden_ratio_search(f, 0.0, P.C_PSNR_MAX, target)
        assert x_got == x_ref and x_last == seen[-1]
        if target <= 4999:
            assert abs(f(seen[-1]) - target) <= 0.5


def test_coreset_stand_in_properties(oracle):
    rng = np.random.default_rng(8)
    x = rng.normal(size=(500, 12)) * 10
    w = rng.integers(1, 9, size=500).astype(np.float64)
    c, cw = oracle.coreset_weighted(x, w, 40, seed=0x42381337)
    assert 1 <= len(c) <= 40 and abs(cw.sum() - w.sum()) < 1e-9          # a weighted summary of at most k points
    assert np.allclose((c * cw[:, None]).sum(0), (x * w[:, None]).sum(0), rtol=1e-9)   # weighted means: first moment preserved
    c2, cw2 = oracle.coreset_weighted(x[:30], w[:30], 40, seed=1)
    assert np.array_equal(c2, x[:30]) and np.array_equal(cw2, w[:30])   # n <= k: every point is its own summary


def test_pipeline_reduce_and_reindex_semantics(oracle):
    from oracle import pipeline as P
    rng = np.random.default_rng(21)
    base = rand_tiles(40, 5)
    canon = base[rng.integers(0, 40, size=(3, 30))]                     # 3 frames x 30 tiles, many exact duplicates
    flags = rng.integers(0, 4, size=(3, 30)).astype(np.uint8)
    psnr = rng.uniform(5, 50, size=(3, 30)).astype(np.float32)
    tiles, tfl, use, tmap, x = P.reduce(canon, flags, psnr, [0], 20)
    unpred = tmap >= 0
    eff = psnr.astype(np.float64); eff[0] /= 10.0
    assert np.array_equal(unpred, ~(eff > x))
    assert use.sum() == unpred.sum() and (np.diff(use) <= 0).all() and abs(len(tiles) - 20) <= 3
    assert all(np.array_equal(tiles[tmap[f, t]], canon[f, t]) for f in range(3) for t in range(30) if tmap[f, t] >= 0)
    didx = rng.integers(0, 4, size=(len(tiles), 64)).astype(np.uint8)
    didx[1] = didx[0]                                                    # identical after dithering: merged by Reindex
    ft, fu, fm = P.reindex(didx, tmap)
    assert len(ft) == len(np.unique(didx[np.unique(tmap[tmap >= 0])], axis=0)) and fu.sum() == unpred.sum()
    assert all(np.array_equal(ft[fm[f, t]], didx[tmap[f, t]]) for f in range(3) for t in range(30) if tmap[f, t] >= 0)
    assert (np.diff(fu) <= 0).all()


def test_host_reindex_and_tile_order_against_pipeline_restatement():
    """The product's host bookkeeping (packed-key sorts, one-gather remap, use counts handed in from the device) against the
    statement-by-statement restatements in oracle/pipeline.py, on inputs with duplicate tiles, rows that tie on the packed sort
    key, unreferenced tiles and tilemap items without a tile."""
    from oracle import pipeline as P
    from tiler_b200 import gtm
    from tiler_b200.encoder import _reindex_order
    rng = np.random.default_rng(33)
    for n_dict, hi in ((700, 16), (300, 2), (5, 3), (1, 4)):
        didx = rng.integers(0, hi, size=(n_dict, 64)).astype(np.uint8)
        if n_dict >= 300:
            didx[10:40] = didx[100:130]                                  # exact duplicates
            didx[50:120, :8] = didx[0, :8]                               # ties on the first 8 bytes, different tails
        tmap = rng.integers(-1, n_dict, size=(4, 500)).astype(np.int32)
        tmap[rng.random(tmap.shape) < 0.3] = -1
        want = P.reindex(didx, tmap)
        got = gtm.reindex(didx, tmap)
        use_dev = np.bincount(tmap[tmap >= 0].reshape(-1), minlength=n_dict)   # what the encoder counts on the device
        got2 = gtm.reindex(didx, tmap, tile_use=use_dev)
        for g in (got, got2):
            assert np.array_equal(g[0], want[0]) and np.array_equal(g[1], want[1]) and np.array_equal(g[2], want[2])
    none = gtm.reindex(didx, np.full((2, 3), -1, np.int32))
    assert none[0].shape == (0, 64) and len(none[1]) == 0 and (none[2] == -1).all()
    # ReindexTiles(True): use count descending, then the 64 pixels as unsigned dwords (first difference decides)
    for n, hi in ((900, 1 << 24), (400, 3), (1, 9)):
        rgb = rng.integers(0, hi, size=(n, 64), dtype=np.int64).astype(np.int32)
        if n >= 400:
            rgb[5:25] = rgb[200:220]
            rgb[30:90, :2] = rgb[0, :2]
            rgb[100:110, 0] = -7                                         # 0xFFFFFFF9: sorts last as an unsigned dword
        use = rng.integers(1, 4, size=n)
        keys = [(-int(use[i]), tuple(int(v) & 0xFFFFFFFF for v in rgb[i]), i) for i in range(n)]
        assert _reindex_order(rgb, use).tolist() == [k[2] for k in sorted(keys)]


def test_optimize_palettes_host_port_matches_pipeline_port():
    """OptimizePalettes + the Powell minimiser (tilingencoder.pas:4246-4432, powell.pas) exist twice: C++ host code of the product
    (libtm_gtm.so) and the Python restatement in oracle/pipeline.py.  The minimiser's path depends on every floating-point detail
    (and on FreePascal's reference semantics of dynamic arrays), so equal permutations on several shapes pin both to each other."""
    from oracle import pipeline as P
    from tiler_b200 import gtm
    for n_pal, pal_size, n_null in ((1, 16, 0), (2, 2, 0), (3, 8, 2), (16, 16, 1), (5, 64, 7)):
        pal = rand_palettes(n_pal, pal_size, 100 + n_pal, n_null=n_null)
        got, it_got = gtm.optimize_palettes(pal, n_threads=3)
        want, it_want = P.optimize_palettes(pal)
        assert it_got == it_want and np.array_equal(got, want)
        assert np.array_equal(np.sort(got, axis=1), np.sort(pal, axis=1))      # a permutation inside each palette, nothing else
    single, _ = gtm.optimize_palettes(rand_palettes(8, 16, 7), n_threads=1)
    multi, _ = gtm.optimize_palettes(rand_palettes(8, 16, 7), n_threads=8)
    assert np.array_equal(single, multi)                                       # palettes of a pass are independent: thread count is free


def test_wavelet_features_against_numpy_haar(oracle):
    """pvsWavelets (WaveletGS, tilingencoder.pas:2727-2762): the oracle against a numpy Haar pyramid written from the definition
    (orthonormal 2x2 averaging / differencing on the 8x8, 4x4 and 2x2 low-pass corners); orthonormal => energy preserved."""
    tiles = rand_tiles(50, 314)
    snake = oracle.dct_snake()
    for t in tiles:
        got = oracle.tile_features_f64(t, oracle.PVS_WAVELETS, False)
        r, g, b = (t & 255).astype(np.float64), ((t >> 8) & 255).astype(np.float64), ((t >> 16) & 255).astype(np.float64)
        y = (r * 0.299 + g * 0.587 + b * 0.114).astype(np.float32).astype(np.float64)
        planes = [y, ((b - y) * 0.492).astype(np.float32).astype(np.float64), ((r - y) * 0.877).astype(np.float32).astype(np.float64)]
        for c, p in enumerate(planes):
            a = p.reshape(8, 8).copy()
            d = 8
            while d >= 2:
                blk = a[:d, :d]
                lo, hi = (blk[:, 0::2] + blk[:, 1::2]) / np.sqrt(2), (blk[:, 0::2] - blk[:, 1::2]) / np.sqrt(2)
                rows = np.concatenate([lo, hi], axis=1)
                lo, hi = (rows[0::2] + rows[1::2]) / np.sqrt(2), (rows[0::2] - rows[1::2]) / np.sqrt(2)
                a[:d, :d] = np.concatenate([lo, hi], axis=0)
                d //= 2
            want = np.empty(64); want[snake] = a.reshape(64)
            assert np.allclose(got[c * 64:(c + 1) * 64], want, rtol=1e-12, atol=1e-9)
            assert abs((got[c * 64:(c + 1) * 64] ** 2).sum() - (p ** 2).sum()) <= 1e-6 * max(1.0, (p ** 2).sum())
