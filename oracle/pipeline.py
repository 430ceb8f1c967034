"""CPU restatement of the encoder's HOST-SIDE steps between the compute stages, written from the reference alone.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): imported by tests/ to check tiler_b200.encoder; it imports
nothing from tiler_b200.  Plain Python / numpy loops and sorts that follow the Pascal statement by statement -- small
clips only.  Every function cites the reference lines it follows.
"""
import math

import numpy as np

from . import oracle as O

C_INV_PHI = 2.0 / (1.0 + math.sqrt(5.0))                                   # cInvPhi (utils.pas:42-43)
C_PSNR_MAX = 10.0 * math.log(255.0 * 255.0 / 0.5) / math.log(10.0)          # cPsnrMaxValue (utils.pas:111)
C_PSYV_EPSILON = 1e-6                                                      # cPsyVEpsilon


def quicksort(items, compare):
    """extern.pas:370-418: the repo's own QuickSort (middle pivot, Hoare partition, the pivot INDEX follows swaps), in
    place on a Python list.  Not stable: ties land where this exact procedure puts them."""
    def qs(first, last):
        if last <= first:
            return
        while True:
            i, j, p = first, last, (first + last) >> 1
            while True:
                while compare(items[i], items[p]) < 0:
                    i += 1
                while compare(items[j], items[p]) > 0:
                    j -= 1
                if i <= j:
                    items[i], items[j] = items[j], items[i]
                    if p == i:
                        p = j
                    elif p == j:
                        p = i
                    i += 1
                    j -= 1
                if i > j:
                    break
            if first < j:
                qs(first, j)
            first = i
            if i >= last:
                break
    qs(0, len(items) - 1)
    return items


def golden_ratio_search(func, min_x, max_x, objective_y, eps_x, eps_y):
    """GoldenRatioSearch (utils.pas:1044-1072), recursion unrolled.  The encoder's state after the search is whatever the
    LAST call of func left behind; the returned abscissa is not re-evaluated (:1048-1052 returns MinX without calling)."""
    while True:
        if abs(min_x - max_x) <= eps_x:                       # SameValue(MinX, MaxX, EpsilonX)
            return min_x
        t = (1.0 - C_INV_PHI) if min_x < max_x else C_INV_PHI
        x = min_x + (max_x - min_x) * t                       # lerp
        y = func(x)
        if abs(y - objective_y) <= eps_y:                     # CompareValue(y, ObjectiveY, EpsilonY) = 0
            return x
        if y < objective_y:
            min_x = x
        else:
            max_x = x


def pad_frames(frames):
    """Tilemap size rounds up to whole tiles (:1776) and the screen IS the tilemap (ReframeUI, :2631-2638); pixels beyond
    the image stay 0 (AllocMem'd tiles, :1310)."""
    n, h, w = frames.shape
    th, tw = (h - 1) // 8 + 1, (w - 1) // 8 + 1
    out = np.zeros((n, th * 8, tw * 8), dtype=np.int32)
    out[:, :h, :w] = frames
    return out, tw, th


def load(frames):
    """TFrame.LoadFromImage + AsyncLoadFromImage (:1289-1411): frame -> 8x8 tiles -> mirror canonicalisation."""
    frames, tw, th = pad_frames(np.asarray(frames, dtype=np.int32))
    n = frames.shape[0]
    nt = tw * th
    tiles = frames.reshape(n, th, 8, tw, 8).transpose(0, 1, 3, 2, 4).reshape(n, nt, 64)
    canon = np.empty_like(tiles)
    flags = np.zeros((n, nt), np.uint8)
    for f in range(n):
        for t in range(nt):
            hm, vm = O.mirror_heuristics(tiles[f, t])
            px = tiles[f, t].reshape(8, 8)
            if hm:
                px = px[:, ::-1]
            if vm:
                px = px[::-1, :]
            canon[f, t] = px.reshape(64)
            flags[f, t] = int(hm) | (int(vm) << 1)
    return frames, canon, flags, tw, th


def predict_motion(frames, canon, flags, tw, th, radius):
    """TTilingEncoder.PredictMotion (:1964-1991): frame f against the previous SOURCE frame; frame 0 against frame 1 (the
    loop starts at -Min(1, High) so the buffer first holds frame 1); a one-frame clip against the zeroed buffer."""
    n = frames.shape[0]
    psnr = np.empty((n, tw * th), np.float32)
    for f in range(n):
        prev = frames[f - 1] if f > 0 else (frames[1] if n > 1 else np.zeros_like(frames[0]))
        cur = O.features_from_rgb_mirrored(canon[f], flags[f])
        _, _, e = O.motion_search(cur, tw, th, O.sliding_features(prev), radius)
        psnr[f] = [O.euclidean_to_psnr(int(v)) for v in e]
    return psnr


def reduce(canon, flags, psnr, seq_starts, tile_count):
    """TTilingEncoder.Reduce (:1908-1926): SolveTileCount = GoldenRatioSearch over STCGREval (:4014-4046), which marks a
    tile predicted when PSNR > x (PSNR / 10.0 > x on the first frame of its keyframe sequence: Single promoted to Double,
    compared with the Double x), transfers the unpredicted tiles (:4048-4103) and merges exact duplicates
    (MakeTilesUnique(True), :4720-4781); then ReindexTiles(True) (:4626-4696).
    -> dictionary tiles, their mirror flags, use counts, tilemap TileIdx [n, nt] (-1 = predicted), threshold."""
    n, nt = canon.shape[:2]
    flat = canon.reshape(-1, 64)
    fl = flags.reshape(-1)
    p64 = psnr.astype(np.float64)
    starts = set(int(s) for s in seq_starts)
    eff = np.stack([p64[f] / 10.0 if f in starts else p64[f] for f in range(n)]).reshape(-1)
    keys = [flat[i].astype(np.uint32).tobytes() for i in range(len(flat))]
    state = {}

    def stcgr_eval(x):
        pred = eff > x
        groups = {}                                            # MakeTilesUnique: identical RGB pixels -> one tile
        for i in np.flatnonzero(~pred):
            groups.setdefault(keys[i], []).append(int(i))
        state["pred"], state["groups"], state["x"] = pred, groups, x
        return float(len(groups))                              # GetTileCount(True)

    target = min(int(tile_count), len(flat))
    golden_ratio_search(stcgr_eval, 0.0, C_PSNR_MAX, float(target), C_PSYV_EPSILON, 0.5)
    groups = state["groups"]
    # ReindexTiles(True): CompareTileUseCountRev (:582-599) = use count descending, then CompareDWord over the 64 pixels
    # (unsigned dwords, first difference decides).  Keys are unique after MakeTilesUnique, so any sort gives this order.
    items = [(-len(m), tuple(flat[m[0]].astype(np.uint32).tolist()), m) for m in groups.values()]
    items.sort(key=lambda it: (it[0], it[1]))
    tile_idx = np.full(n * nt, -1, np.int32)
    rep = []
    for new, (_, _, members) in enumerate(items):
        tile_idx[members] = new
        rep.append(members[0])      # which duplicate survives is unspecified in the reference (non-stable sort): first in frame order here
    rep = np.asarray(rep, dtype=np.int64)
    use = np.asarray([-it[0] for it in items], dtype=np.int32)
    return flat[rep].copy(), fl[rep].copy(), use, tile_idx.reshape(n, nt), state["x"]


def palettize(dict_tiles, use_count, palette_count, seed, dithering_mode=O.PVS_WEIGHTED_SPE_DCT, coreset_iters=8):
    """DoPalettization (:4105-4245): LAB features (ComputeTilePsyVisFeatures, DitheringMode) -> coreset of 8 x PaletteCount
    points, tiles inserted with weight UseCount (:4149-4173; BICO itself is unpinned: oracle.coreset_weighted defines the
    stand-in) -> nearest coreset point per tile (ANN, eps 0, :4183-4188) -> k-means of the coreset points, UNWEIGHTED
    (yakmo, :4198-4207; k-means++ draw order unpinned: our seeded generator) -> identity mapping when the coreset has
    <= PaletteCount points (:4214-4219), no k-means when PaletteCount = 1 (:4209-4212) -> palettes re-indexed by
    descending tile count with the repo's QuickSort (:4221-4244).  -> PalIdx_Initial per tile."""
    feats = np.stack([O.tile_features_f64(t, dithering_mode, True) for t in dict_tiles])
    n_core_req = palette_count << 3
    core, _ = O.coreset_weighted(feats, np.asarray(use_count, dtype=np.float64), n_core_req, seed, coreset_iters)
    ann, _ = O.knn_double(core, feats)
    if len(core) > palette_count:
        if palette_count > 1:
            yk, _, _, _ = O.kmeans_lloyd(core, O.kmeanspp_init(core, palette_count, seed), max_iter=300)
        else:
            yk = np.zeros(len(core), np.int32)
    else:
        yk = np.arange(len(core), dtype=np.int32)
    pals = [[0, p] for p in range(palette_count)]              # [UseCount, PalIdx_Initial]
    for a in ann:
        pals[int(yk[a])][0] += 1
    quicksort(pals, lambda a, b: (b[0] > a[0]) - (b[0] < a[0]))   # ComparePaletteUseCount (utils.pas:750-753)
    lut = np.empty(palette_count, np.int32)
    for new, (_, old) in enumerate(pals):
        lut[old] = new
    return lut[yk[ann]].astype(np.int32)


def quantize(dict_tiles, tile_pal, palette_count, palette_size, seed):
    """DoQuantization / QuantizeUsingYakmo per palette (:4434-4564)."""
    return np.stack([O.quantize_palette(dict_tiles[tile_pal == p].reshape(-1), palette_size, seed=seed)[0]
                     for p in range(palette_count)])


# ------------------------------------------------------------------ OptimizePalettes (tilingencoder.pas:4246-4432) over powell.pas
def _sign(x):
    return 1.0 if x > 0 else (-1.0 if x < 0 else 0.0)


def _bracket(f, xa, xb):
    """powell.pas:56-146."""
    gold, small, grow = (1 + math.sqrt(5)) / 2, 1e-21, 110
    fa, fb = f(xa), f(xb)
    if fa < fb:
        xa, xb, fa, fb = xb, xa, fb, fa
    xc = xb + gold * (xb - xa)
    fc = f(xc)
    it = 0
    while fc < fb:
        tmp1 = (xb - xa) * (fb - fc)
        tmp2 = (xb - xc) * (fb - fa)
        val = tmp2 - tmp1
        denom = 2 * small if abs(val) < small else 2 * val
        w = xb - ((xb - xc) * tmp2 - (xb - xa) * tmp1) / denom
        wlim = xb + grow * (xc - xb)
        if it > 1000:
            raise RuntimeError("bracket: Too many iterations")
        it += 1
        fw = 0
        if (w - xc) * (xb - w) > 0:
            fw = f(w)
            if fw < fc:
                xa, xb, fa, fb = xb, w, fb, fw
                break
            elif fw > fb:
                xc, fc = w, fw
                break
            w = xc + gold * (xc - xb)
            fw = f(w)
        elif (w - wlim) * (wlim - xc) >= 0:
            w = wlim
            fw = f(w)
        elif (w - wlim) * (xc - w) > 0:
            fw = f(w)
            if fw < fc:
                xb, xc = xc, w
                w = xc + gold * (xc - xb)
                fb, fc = fc, fw
                fw = f(w)
        else:
            w = xc + gold * (xc - xb)
            fw = f(w)
        xa, xb, xc = xb, xc, w
        fa, fb, fc = fb, fc, fw
    if xa > xc:
        xa, xc = xc, xa
    return xa, xb, xc


def _brent_helper(f, a, x, b, fx, xtol, maxiter):
    """powell.pas:148-260."""
    cg = (3 - math.sqrt(5)) / 2
    if a > b:
        a, b = b, a
    w = v = x
    fw = fv = fx
    deltax = rat = 0.0
    it = 0
    while it < maxiter:
        xmid = 0.5 * (a + b)
        if abs(x - xmid) <= 2 * xtol - 0.5 * (b - a):
            break
        if abs(deltax) <= xtol:
            deltax = a - x if x >= xmid else b - x
            rat = cg * deltax
        else:
            tmp1 = (x - w) * (fx - fv)
            tmp2 = (x - v) * (fx - fw)
            p = (x - v) * tmp2 - (x - w) * tmp1
            tmp2 = 2 * (tmp2 - tmp1)
            if tmp2 > 0:
                p = -p
            tmp2 = abs(tmp2)
            dx_temp = deltax
            deltax = rat
            if p > tmp2 * (a - x) and p < tmp2 * (b - x) and abs(p) < abs(0.5 * tmp2 * dx_temp):
                rat = p / tmp2
                u = x + rat
                if u - a < xtol or b - u < xtol:
                    rat = _sign(xmid - x) * xtol
            else:
                deltax = a - x if x >= xmid else b - x
                rat = cg * deltax
        u = x + rat if abs(rat) > xtol else x + _sign(rat) * xtol
        fu = f(u)
        if fu > fx:
            if u < x:
                a = u
            else:
                b = u
            if fu <= fw or w == x:
                v, w, fv, fw = w, u, fw, fu
            elif fu <= fv or v == x or v == w:
                v, fv = u, fu
        else:
            if u >= x:
                a = x
            else:
                b = x
            v, w, x = w, x, u
            fv, fw, fx = fw, fx, fu
        it += 1
    return x, fx


def _linesearch_powell(f, p, xi, xtol):
    """powell.pas:294-324; p and xi are Python lists modified IN PLACE (var parameters)."""
    n = len(p)
    sqsos = math.sqrt(sum(v * v for v in xi))
    atol = 1.0
    if sqsos != 0:
        atol = 5 * xtol / sqsos
    atol = min(0.1, atol)
    along = lambda t: f([p[i] + t * xi[i] for i in range(n)])
    a, b, c = _bracket(along, 0.0, 1.0)
    alpha, fret = _brent_helper(along, a, b, c, along(b), atol, 100)
    for i in range(n):
        xi[i] = xi[i] * alpha
        p[i] = p[i] + xi[i]
    return fret


def powell_minimize(f, x, scale, xtol, ftol, maxiter):
    """powell.pas:326-385.  Lists are references like FreePascal dynamic arrays: `direc[n - 1] = direc1` shares the array."""
    n = len(x)
    direc1 = [0.0] * n
    tmp = [0.0] * n
    direc = [[scale if i == j else 0.0 for j in range(n)] for i in range(n)]
    fval = f(x)
    x1 = list(x)
    it = 0
    while True:
        fx = fval
        bigind, delta = 0, 0.0
        for i in range(n):
            fx2 = fval
            fval = _linesearch_powell(f, x, direc[i], xtol)
            if fx2 - fval > delta:
                delta, bigind = fx2 - fval, i
        it += 1
        if fx - fval <= ftol or it >= maxiter:
            break
        for i in range(n):
            direc1[i] = x[i] - x1[i]
            tmp[i] = x[i] + direc1[i]
            x1[i] = x[i]
        fx2 = f(tmp)
        if fx > fx2:
            t = 2 * (fx + fx2 - 2 * fval)
            temp = fx - fval - delta
            t = t * temp * temp
            temp = fx - fx2
            t = t - delta * temp * temp
            if t < 0:
                fval = _linesearch_powell(f, x, direc1, xtol)
                direc[bigind] = direc[n - 1]
                direc[n - 1] = direc1
    return fval


def _optimize_one(args):
    """DoPal (:4320-4392) for one palette of one pass: PowellMinimize over PowellOP (:4264-4305)."""
    row_a, acc, mean, S = args
    rgb = lambda c: (c & 255, (c >> 8) & 255, (c >> 16) & 255)
    last = [None]

    def op(x):
        perm = [(0, 0)] + [(int(round(x[c - 1] * 1000)), c) for c in range(1, S)]     # (Count, Index); Python round = half to even
        perm.sort()
        sd = [0, 0, 0]
        row = []
        for c in range(S):
            col = row_a[perm[c][1]]
            row.append(col)
            for k, v in enumerate(rgb(col)):
                sd[k] += (acc[c][k] + v - mean[k]) ** 2
        last[0] = row
        return -((299 * math.sqrt(sd[0] / S) + 587 * math.sqrt(sd[1] / S) + 114 * math.sqrt(sd[2] / S)) / 1000)
    x = [float(c) for c in range(1, S)]
    powell_minimize(op, x, 1.0, 1.0, 1.0, 0x7FFFFFFF)
    f = -op(x)
    return last[0], f


def optimize_palettes(palettes):
    """TTilingEncoder.OptimizePalettes (:4307-4432) with PowellOP (:4264-4305), statement by statement.  The palettes of a pass
    are independent (the reference runs them on its thread pool, :4415); large palette sets use worker processes."""
    pal = [[int(c) & 0xFFFFFFFF for c in row] for row in np.asarray(palettes)]
    P, S = len(pal), len(pal[0])
    rgb = lambda c: (c & 255, (c >> 8) & 255, (c >> 16) & 255)
    mean = [0, 0, 0]
    for row in pal:
        for c in row:
            for k, v in enumerate(rgb(c)):
                mean[k] += v
    mean = [m // S for m in mean]
    pool = None
    if P >= 64:
        import multiprocessing as mp
        import os
        pool = mp.get_context("fork").Pool(min(os.cpu_count() or 1, 32))
    prev_fsum = fsum = 0.0
    iteration = 0
    try:
        while True:
            prev_fsum = max(fsum, prev_fsum)
            iteration += 1
            # "accumulate the whole palette except the one that will be permutated" (:4357-4377): column sums over all palettes
            # minus the palette's own colours (integers: the same values as the reference's loop over the other palettes)
            tot = [[sum(rgb(pal[p][c])[k] for p in range(P)) for k in range(3)] for c in range(S)]
            jobs = [(pal[a], [[tot[c][k] - rgb(pal[a][c])[k] for k in range(3)] for c in range(S)], mean, S) for a in range(P)]
            res = pool.map(_optimize_one, jobs, chunksize=8) if pool else [_optimize_one(j) for j in jobs]
            pal = [r[0] for r in res]
            fsum = sum(r[1] for r in res) / P
            if fsum <= prev_fsum:
                break
    finally:
        if pool:
            pool.close()
    out = np.array(pal, dtype=np.uint32).astype(np.int64)
    return np.where(out >= 1 << 31, out - (1 << 32), out).astype(np.int32), iteration


def reindex(dict_idx, tile_idx):
    """TTilingEncoder.Reindex (:1993-2038): MakeTilesUnique(False) merges dictionary tiles with identical palette indices
    into the first of the sorted run, use counts are recounted from every tilemap item with TileIdx >= 0, then
    ReindexTiles(False): drop unused tiles, order by (use count descending, CompareByte on the 64 indices), remap."""
    dict_idx = np.asarray(dict_idx, dtype=np.uint8).reshape(-1, 64)
    tmap = np.asarray(tile_idx, dtype=np.int64)
    first_of = {}
    merge = np.arange(len(dict_idx))
    for i in range(len(dict_idx)):
        k = dict_idx[i].tobytes()
        merge[i] = first_of.setdefault(k, i)
    use = np.zeros(len(dict_idx), np.int64)
    flat = tmap.reshape(-1)
    merged = np.where(flat >= 0, merge[np.maximum(flat, 0)], -1)
    for t in merged[merged >= 0]:
        use[t] += 1
    keep = [i for i in range(len(dict_idx)) if use[i] > 0]
    keep.sort(key=lambda i: (-int(use[i]), dict_idx[i].tobytes()))
    new_of = np.full(len(dict_idx), -1, np.int64)
    for new, i in enumerate(keep):
        new_of[i] = new
    out = np.where(merged >= 0, new_of[np.maximum(merged, 0)], -1).astype(np.int32).reshape(tmap.shape)
    return dict_idx[keep].copy(), use[keep].astype(np.int32), out


def merge_tiles(dict_tiles, use_count, clusters, best, tile_idx):
    """InitMergeTiles / MergeTiles / FinishMergeTiles (:4783-4840) followed by ReindexTiles(True) (:4626-4696), statement by
    statement: clusters[i] = cluster of tile i, best[c] = the tile cluster c keeps."""
    n = len(dict_tiles)
    use = [int(u) for u in use_count]
    active = [True] * n
    merge_index = [-1] * n                                                    # InitMergeTiles
    members = {}
    for i, c in enumerate(clusters):
        members.setdefault(int(c), []).append(i)
    for c, idxs in members.items():                                           # MergeTiles(idxs, len, best[c], nil, nil)
        b = int(best[c])
        for t in idxs:
            if t == b:
                continue
            use[b] += use[t]
            active[t] = False
            use[t] = 0
            merge_index[t] = b
    tmap = np.asarray(tile_idx, dtype=np.int64).copy()
    flat = tmap.reshape(-1)
    for i in range(len(flat)):                                                # FinishMergeTiles
        t = flat[i]
        if t >= 0 and merge_index[t] >= 0:
            flat[i] = merge_index[t]
    keep = [i for i in range(n) if active[i] and use[i] > 0]                  # ReindexTiles(True)
    keep.sort(key=lambda i: (-use[i], tuple(np.asarray(dict_tiles[i]).astype(np.uint32).tolist())))
    new_of = {old: new for new, old in enumerate(keep)}
    out = np.array([new_of[t] if t >= 0 else -1 for t in flat], dtype=np.int32).reshape(tmap.shape)
    return np.asarray(dict_tiles)[keep].copy(), np.array([use[i] for i in keep], dtype=np.int32), out


def encode(frames, seqs, tile_count, palette_count, palette_size, seed, radius=32, use_tk=True, y2_mixed=4, extended=True,
           optimize=True):
    """TTilingEncoder.Run (:5529-5554) up to, but not including, the stream writer: Load -> PredictMotion -> Reduce ->
    PreparePalettes (with OptimizePalettes unless optimize=False) -> Dither -> Reconstruct -> Reindex."""
    frames_p, canon, flags, tw, th = load(frames)
    psnr = predict_motion(frames_p, canon, flags, tw, th, radius)
    dtiles, dflags, use, _, x = reduce(canon, flags, psnr, [s for s, _ in seqs], tile_count)
    tpal = palettize(dtiles, use, palette_count, seed)
    pal = quantize(dtiles, tpal, palette_count, palette_size, seed)
    if optimize:
        pal, _ = optimize_palettes(pal)
    didx = O.dither(dtiles, dflags, tpal, pal, use_tk=use_tk, y2_mixed_colors=y2_mixed)
    dfeat = O.features_from_pal(didx, tpal, pal)
    parts = [O.reconstruct_sequence(canon[s0:s1 + 1], flags[s0:s1 + 1], tw, th, dfeat, didx, tpal, pal, radius=radius, extended=extended)
             for s0, s1 in seqs]
    tm = {k: np.concatenate([p[k] for p in parts]) for k in ("tile_idx", "pal_idx", "pred_x", "pred_y", "is_pred", "recon", "err")}
    ftiles, fuse, fmap = reindex(didx, tm["tile_idx"])
    return {"tiles": ftiles, "use_count": fuse, "tile_idx": fmap, "pal_idx": tm["pal_idx"], "pred_x": tm["pred_x"], "pred_y": tm["pred_y"],
            "is_pred": tm["is_pred"], "recon": tm["recon"], "err": tm["err"], "palettes": pal, "mirror": flags, "tile_pal": tpal,
            "dict_tiles_rgb": dtiles, "threshold": x, "psnr": psnr, "tw": tw, "th": th}
