"""Host-side mirror of the reference's operator interface for the hot path, over the C ABI of libtm_gpu.so.

Names follow the reference (tilingencoder.pas / extern.pas): tiles, palettes, features (TDCT), tilemap items.  Every
function accepts either numpy arrays (host buffers: copied to the GPU and back inside the call) or torch CUDA tensors
(device buffers: nothing is copied, work is enqueued on torch's current stream and results are CUDA tensors).
torch is used for device memory and streams only.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import TmError, check  # noqa: F401

DCT = 192           # cTileDCTSize (utils.pas:40)
TILE_PX = 64
NULL_COLOR = np.int32(-65281)  # cDitheringNullColor 0xFFFF00FF (utils.pas:45)
PVS_DCT, PVS_WEIGHTED_DCT, PVS_WAVELETS, PVS_SPE_DCT, PVS_WEIGHTED_SPE_DCT = range(5)  # TPsyVisMode (tilingencoder.pas:21)

try:  # torch is optional for pure host use
    import torch
except Exception:  # pragma: no cover
    torch = None

_NP2T = {}
if torch is not None:
    _NP2T = {np.dtype(np.int16): torch.int16, np.dtype(np.int32): torch.int32, np.dtype(np.uint8): torch.uint8,
             np.dtype(np.float64): torch.float64, np.dtype(np.int64): torch.int64, np.dtype(np.uint32): torch.int32,
             np.dtype(np.float32): torch.float32}


def _is_dev(x):
    return torch is not None and isinstance(x, torch.Tensor) and x.is_cuda


class _Call:
    """Collects the arguments of one ABI call; decides host vs device from the first array argument."""

    def __init__(self, *arrays):
        self.dev = any(_is_dev(a) for a in arrays)
        self.keep = []
        if self.dev:
            self.device = next(a.device for a in arrays if _is_dev(a))
            torch.cuda.set_device(self.device)
            check(_lib.lib().tm_set_stream(C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        else:
            check(_lib.lib().tm_set_stream(None))

    def inp(self, x, dtype, shape=None):
        if x is None:
            return None
        if self.dev:
            if not _is_dev(x):
                x = torch.as_tensor(np.ascontiguousarray(x, dtype=dtype)).to(self.device)
            t = x.contiguous()
            want = _NP2T[np.dtype(dtype)]
            if t.dtype != want:
                raise TypeError(f"expected {want}, got {t.dtype}")
            self.keep.append(t)
            return C.c_void_p(t.data_ptr())
        a = np.ascontiguousarray(x, dtype=dtype)
        self.keep.append(a)
        return C.c_void_p(a.ctypes.data)

    def out(self, shape, dtype):
        if self.dev:
            t = torch.empty(shape, dtype=_NP2T[np.dtype(dtype)], device=self.device)
            self.keep.append(t)
            return t, C.c_void_p(t.data_ptr())
        a = np.empty(shape, dtype=dtype)
        return a, C.c_void_p(a.ctypes.data)


def _n_rows(x, width):
    n = int(np.prod(x.shape)) // width
    return n


def device_count():
    return int(_lib.lib().tm_device_count())


def kernel_launches():
    return int(_lib.lib().tm_kernel_launches())


def profile_enable(on=True):
    check(_lib.lib().tm_profile_enable(int(on)))


def profile_read(name):
    """-> (total_ms, launches) of the kernels recorded under `name` since the last read."""
    ms, n = C.c_double(), C.c_int64()
    check(_lib.lib().tm_profile_read(name.encode(), C.byref(ms), C.byref(n)))
    return ms.value, n.value


def synchronize():
    check(_lib.lib().tm_synchronize())


FEATURES_EXACT, FEATURES_FAST = 0, 1


def set_feature_mode(mode):
    """Arithmetic of the sliding-window features (DoDCTs): FEATURES_EXACT = DCTInner_asm's summation order, bit-exact (default);
    FEATURES_FAST = separable f64 DCT, <= 1 LSB away on < 1e-3 of the coefficients.  Returns the previous mode."""
    prev = int(_lib.lib().tm_get_feature_mode())
    check(_lib.lib().tm_set_feature_mode(int(mode)))
    return prev


# ------------------------------------------------------------------ features (tilingencoder.pas:3049-3182)
def features_from_rgb(rgb):
    """ConvertToCpnPixels + ComputeCpnPixelsPsyVisFeatures(pvsWeightedDCT): RGB tiles [n,64] int32 -> int16 [n,192]."""
    c = _Call(rgb)
    n = _n_rows(rgb, 64)
    out, po = c.out((n, DCT), np.int16)
    check(_lib.lib().tm_features_from_rgb(c.inp(rgb, np.int32), n, po))
    return out


def features_from_rgb_mirrored(rgb, flags):
    """Features of stored tiles read through mirror flags (bit 0 H, bit 1 V): ConvertToCpnPixels with AHMirror / AVMirror."""
    c = _Call(rgb, flags)
    n = _n_rows(rgb, 64)
    out, po = c.out((n, DCT), np.int16)
    check(_lib.lib().tm_features_from_rgb_mirrored(c.inp(rgb, np.int32), c.inp(flags, np.uint8), n, po))
    return out


def features_from_pal(pal_idx, tile_pal, palettes):
    """PrepareReconstruct.DoPsyV (tilingencoder.pas:4570-4583): indexed tiles + their palettes -> int16 [n,192]."""
    c = _Call(pal_idx, tile_pal, palettes)
    n = _n_rows(pal_idx, 64)
    n_pal, pal_size = palettes.shape
    out, po = c.out((n, DCT), np.int16)
    check(_lib.lib().tm_features_from_pal(c.inp(pal_idx, np.uint8), c.inp(tile_pal, np.int32), c.inp(palettes, np.int32),
                                          int(pal_size), int(n_pal), n, po))
    return out


def features_f64(rgb, mode=PVS_WEIGHTED_SPE_DCT, use_lab=True):
    """ComputeTilePsyVisFeatures (tilingencoder.pas:3133-3182): f64 [n,192]."""
    c = _Call(rgb)
    n = _n_rows(rgb, 64)
    out, po = c.out((n, DCT), np.float64)
    check(_lib.lib().tm_features_f64(c.inp(rgb, np.int32), n, int(mode), int(use_lab), po))
    return out


def mirror_canonicalise(rgb):
    """TFrame.AsyncLoadFromImage mirror step (tilingencoder.pas:1393-1411). Returns (flipped tiles, flags bit0=H bit1=V)."""
    c = _Call(rgb)
    n = _n_rows(rgb, 64)
    if c.dev:
        tiles = rgb.clone().contiguous().view(n, 64)
        c.keep.append(tiles)
        pt = C.c_void_p(tiles.data_ptr())
    else:
        tiles = np.array(rgb, dtype=np.int32, copy=True).reshape(n, 64)
        pt = C.c_void_p(tiles.ctypes.data)
    flags, pf = c.out((n,), np.uint8)
    check(_lib.lib().tm_mirror_canonicalise(pt, n, pf))
    return tiles, flags


def distance_pairs(a, b):
    """CompareEuclideanDCTPtr (utils.pas:541-557) for n vector pairs -> uint32 (int32 bit pattern on device)."""
    c = _Call(a, b)
    n = _n_rows(a, DCT)
    out, po = c.out((n,), np.uint32)
    check(_lib.lib().tm_distance_pairs(c.inp(a, np.int16), c.inp(b, np.int16), n, po))
    return out


# ------------------------------------------------------------------ k-NN
class KnnShort:
    """Exact k-NN over int16[192] rows: the ANN_short.dll kd-tree (extern.pas:182-185), batched."""

    def __init__(self, feat):
        c = _Call(feat)
        self.n = _n_rows(feat, DCT)
        h = C.c_void_p()
        check(_lib.lib().tm_knn_short_create(c.inp(feat, np.int16), self.n, C.byref(h)))
        self._h = h

    def search(self, q, k=1, sorted=True):
        """-> (idx [n_q,k] int32, dist [n_q,k] uint32), rows ordered by (distance, index) when sorted."""
        c = _Call(q)
        n_q = _n_rows(q, DCT)
        idx, pi = c.out((n_q, k), np.int32)
        dist, pd = c.out((n_q, k), np.uint32)
        check(_lib.lib().tm_knn_short_batch(self._h, c.inp(q, np.int16), n_q, int(k), pi, pd, int(sorted)))
        return idx, dist

    def close(self):
        if self._h:
            _lib.lib().tm_knn_short_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def knn_double(dict_pts, q):
    """Exact NN of f64 vectors: ann_kdtree_search with eps 0 (extern.pas:180), batched. -> (idx, dist)."""
    c = _Call(dict_pts, q)
    n_dict, dim = dict_pts.shape
    n_q = q.shape[0]
    idx, pi = c.out((n_q,), np.int32)
    dist, pd = c.out((n_q,), np.float64)
    check(_lib.lib().tm_knn_double_batch(c.inp(dict_pts, np.float64), n_dict, int(dim), c.inp(q, np.float64), n_q, pi, pd))
    return idx, dist


# ------------------------------------------------------------------ dithering (tilingencoder.pas:1873-1907)
def dither(rgb, mirror_flags, pair_pal, palettes, use_thomas_knoll=True, y2_mixed_colors=4, pair_tile=None):
    """TTilingEncoder.Dither generalised to (tile, palette) pair lists. -> uint8 [n_pairs,64] palette indices."""
    c = _Call(rgb, pair_pal, palettes)
    n_tiles = _n_rows(rgb, 64)
    n_pairs = int(np.prod(pair_pal.shape))
    n_pal, pal_size = palettes.shape
    out, po = c.out((n_pairs, 64), np.uint8)
    check(_lib.lib().tm_dither(c.inp(rgb, np.int32), c.inp(mirror_flags, np.uint8), n_tiles, c.inp(pair_tile, np.int32),
                               c.inp(pair_pal, np.int32), n_pairs, c.inp(palettes, np.int32), int(pal_size), int(n_pal),
                               int(use_thomas_knoll), int(y2_mixed_colors), po))
    return out


# ------------------------------------------------------------------ k-means (yakmo contract, extern.pas:198-203)
def kmeans_fit(x, k, init=None, seed=1, max_iter=300, nan_empty=False):
    """Lloyd from explicit initial centroids (or seeded k-means++). -> labels, centroids, inertia, iterations."""
    c = _Call(x, init)
    n, dim = x.shape
    labels, pl = c.out((n,), np.int32)
    cent, pc = c.out((k, dim), np.float64)
    inertia, iters = C.c_double(), C.c_int()
    check(_lib.lib().tm_kmeans_fit(c.inp(x, np.float64), n, int(dim), int(k), int(max_iter), c.inp(init, np.float64),
                                   C.c_uint64(seed), int(nan_empty), pl, pc, C.byref(inertia), C.byref(iters)))
    return labels, cent, inertia.value, iters.value


def coreset_weighted(x, w, k, seed, max_iter=8):
    """The BICO stand-in (bico_create / insert_line / get_results, tilingencoder.pas:4149-4172) in one call: <= k weighted
    summary points of the weighted rows x [n, dim].  -> (centroids [m, dim], weights [m]) as numpy arrays (the host sizes
    its next step from m, :4168)."""
    c = _Call(x, w)
    n, dim = x.shape
    cent = np.empty((int(k), int(dim)), dtype=np.float64)
    wts = np.empty(int(k), dtype=np.float64)
    m = C.c_int64()
    check(_lib.lib().tm_coreset_weighted(c.inp(x, np.float64), c.inp(w, np.float64), n, int(dim), int(k), int(max_iter), C.c_uint64(seed),
                                         C.c_void_p(cent.ctypes.data), C.c_void_p(wts.ctypes.data), C.byref(m)))
    return cent[:m.value].copy(), wts[:m.value].copy()


def kmeans_fit_i16(x, k, init, max_iter=300, nan_empty=False):
    """Lloyd on int16 feature rows [n,192]: tensor-core candidate search + exact f64 decision (tm_kmeans_fit_i16).
    -> labels, centroids (f64), inertia, iterations, number of points that needed the exact f64 fallback."""
    c = _Call(x, init)
    n = _n_rows(x, DCT)
    labels, pl = c.out((n,), np.int32)
    cent, pc = c.out((k, DCT), np.float64)
    inertia, iters, amb = C.c_double(), C.c_int(), C.c_int64()
    check(_lib.lib().tm_kmeans_fit_i16(c.inp(x, np.int16), n, int(k), int(max_iter), c.inp(init, np.float64), int(nan_empty),
                                       pl, pc, C.byref(inertia), C.byref(iters), C.byref(amb)))
    return labels, cent, inertia.value, iters.value, amb.value


def kmeans_partial_step_i16(x, centroids, labels):
    """Sharded variant of kmeans_fit_i16's step (see kmeans_partial_step)."""
    c = _Call(x, centroids, labels)
    n = _n_rows(x, DCT)
    k = centroids.shape[0]
    sums, ps = c.out((k, DCT), np.float64)
    counts, pc = c.out((k,), np.int64)
    changed, inertia = C.c_int64(), C.c_double()
    if c.dev:
        lab = labels.contiguous()
        c.keep.append(lab)
        plab = C.c_void_p(lab.data_ptr())
    else:
        lab = np.ascontiguousarray(labels, dtype=np.int32)
        plab = C.c_void_p(lab.ctypes.data)
    check(_lib.lib().tm_kmeans_partial_step_i16(c.inp(x, np.int16), n, int(k), c.inp(centroids, np.float64), plab, ps, pc,
                                                C.byref(changed), C.byref(inertia)))
    return lab, sums, counts, changed.value, inertia.value


class KmeansI16Shard:
    """tm_kmeans_i16_create / _step / _destroy: one rank's shard of the sharded Lloyd loop.  The points are split into limb
    rows once; step() takes and returns device tensors (or numpy arrays) and copies nothing to the host."""

    def __init__(self, x, k):
        c = _Call(x)
        self._x = c.inp(x, np.int16)          # keeps the rows alive: the handle references device rows in place
        self._keep = c.keep
        self.n, self.k, self.dev = _n_rows(x, DCT), int(k), c.dev
        self.device = c.device if c.dev else None
        h = C.c_void_p()
        check(_lib.lib().tm_kmeans_i16_create(self._x, self.n, self.k, C.byref(h)))
        self._h = h

    def step(self, centroids, labels, sums=None, counts=None, stats=None, inertia=None):
        """labels updated in place; -> (sums [k,192] f64, counts [k] int64).  stats: int64[2] accumulator tensor (changed,
        exact-scan points), inertia: f64[1] tensor, both optional."""
        c = _Call(centroids, labels)
        if sums is None:
            sums, _ = c.out((self.k, DCT), np.float64)
        if counts is None:
            counts, _ = c.out((self.k,), np.int64)

        def ptr(a):
            if a is None:
                return None
            return C.c_void_p(a.data_ptr()) if _is_dev(a) else C.c_void_p(a.ctypes.data)
        check(_lib.lib().tm_kmeans_i16_step(self._h, c.inp(centroids, np.float64), ptr(labels), ptr(sums), ptr(counts), ptr(stats), ptr(inertia)))
        return sums, counts

    def close(self):
        if self._h:
            _lib.lib().tm_kmeans_i16_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def kmeans_partial_step(x, centroids, labels):
    """One Lloyd step on a shard: assignment + per-cluster partial sums/counts (to be all-reduced across GPUs)."""
    c = _Call(x, centroids, labels)
    n, dim = x.shape
    k = centroids.shape[0]
    sums, ps = c.out((k, dim), np.float64)
    counts, pc = c.out((k,), np.int64)
    changed, inertia = C.c_int64(), C.c_double()
    if c.dev:
        lab = labels.contiguous()
        c.keep.append(lab)
        plab = C.c_void_p(lab.data_ptr())
    else:
        lab = np.ascontiguousarray(labels, dtype=np.int32)
        plab = C.c_void_p(lab.ctypes.data)
    check(_lib.lib().tm_kmeans_partial_step(c.inp(x, np.float64), n, int(dim), int(k), c.inp(centroids, np.float64), plab,
                                            ps, pc, C.byref(changed), C.byref(inertia)))
    return lab, sums, counts, changed.value, inertia.value


def kmeans_finish_step(sums, counts, centroids, nan_empty=False):
    c = _Call(sums, counts, centroids)
    k, dim = sums.shape
    if c.dev:
        cent = centroids.clone().contiguous()
        c.keep.append(cent)
        pc = C.c_void_p(cent.data_ptr())
    else:
        cent = np.array(centroids, dtype=np.float64, copy=True)
        pc = C.c_void_p(cent.ctypes.data)
    check(_lib.lib().tm_kmeans_finish_step(c.inp(sums, np.float64), c.inp(counts, np.int64), int(k), int(dim), int(nan_empty), pc))
    return cent


# ------------------------------------------------------------------ palette colour quantisation (:4434-4564)
def palquant_kmeans(rgb, tile_pal, n_pal, pal_size, init=None, seed=1):
    """QuantizeUsingYakmo + DoQuantization for every palette. -> palettes int32 [n_pal,pal_size], iterations."""
    c = _Call(rgb, tile_pal)
    n_tiles = _n_rows(rgb, 64)
    out, po = c.out((n_pal, pal_size), np.int32)
    iters = C.c_int()
    check(_lib.lib().tm_palquant_kmeans(c.inp(rgb, np.int32), c.inp(tile_pal, np.int32), n_tiles, int(n_pal), int(pal_size),
                                        c.inp(init, np.float64), C.c_uint64(seed), po, C.byref(iters)))
    return out, iters.value


# ------------------------------------------------------------------ dlquant (dlquant/quantizer.c; extern.pas:195-196)
def dlquant_batch(images, quant_to, lookup_bpc=5, which=3):
    """images: list of uint8 arrays [n_px, 3] (R,G,B).  -> palettes uint8 [n_img, quant_to, 3], counts int32 [n_img]."""
    imgs = [np.ascontiguousarray(im, dtype=np.uint8).reshape(-1, 3) for im in images]
    off = np.zeros(len(imgs) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(im) for im in imgs])
    flat = np.concatenate(imgs, axis=0)
    pal = np.empty((len(imgs), quant_to, 3), dtype=np.uint8)
    cnt = np.empty(len(imgs), dtype=np.int32)
    check(_lib.lib().tm_set_stream(None))
    fn = _lib.lib().tm_dl3quant_batch if which == 3 else _lib.lib().tm_dl1quant_batch
    check(fn(C.c_void_p(flat.ctypes.data), C.c_void_p(off.ctypes.data), len(imgs), int(quant_to), int(lookup_bpc),
             C.c_void_p(pal.ctypes.data), C.c_void_p(cnt.ctypes.data)))
    return pal, cnt


def dlquant_dropin(rgb888, width, height, quant_to, lookup_bpc=5, which=3):
    """dl3quant / dl1quant with the DLL's own signature -> (return code, palette [quant_to, 3])."""
    buf = np.ascontiguousarray(rgb888, dtype=np.uint8).reshape(-1).copy()
    userpal = np.full((3, 65536), 255, dtype=np.uint8)
    fn = _lib.lib().dl3quant if which == 3 else _lib.lib().dl1quant
    rc = fn(C.c_void_p(buf.ctypes.data), int(width), int(height), int(quant_to), int(lookup_bpc), C.c_void_p(userpal.ctypes.data))
    return rc, userpal[:, :quant_to].T.copy(), userpal


# ------------------------------------------------------------------ matcher (tilingencoder.pas:4566-4613, 1464-1659)
class Matcher:
    """PrepareReconstruct + the k-NN / extended-palette part of TFrame.Reconstruct.DoXY."""

    K_EPU = 64  # cEpuKnnK (tilingencoder.pas:1433)

    def __init__(self, dict_idx, dict_pal, palettes, extended=True):
        c = _Call(dict_idx, dict_pal, palettes)
        self.n_dict = _n_rows(dict_idx, 64)
        self.n_pal, self.pal_size = palettes.shape
        self.extended = bool(extended)
        h = C.c_void_p()
        check(_lib.lib().tm_matcher_create(c.inp(dict_idx, np.uint8), c.inp(dict_pal, np.int32), self.n_dict,
                                           c.inp(palettes, np.int32), int(self.pal_size), int(self.n_pal), int(extended),
                                           C.byref(h)))
        self._h = h

    def _run(self, fn, x, width, dtype, k, out=None):
        c = _Call(x)
        n_q = _n_rows(x, width)
        if out is not None and not c.dev:
            # caller-provided HOST result arrays (tile int32, pal int32, err uint32; contiguous, n_q each) -- like the C ABI itself,
            # where the host owns the output buffers: page-locked ones make the read-back a plain DMA instead of a staged copy
            tile, pal, err = out
            for a, dt in ((tile, np.int32), (pal, np.int32), (err, np.uint32)):
                if not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags.c_contiguous and a.size == n_q):
                    raise ValueError("out: three contiguous numpy arrays (int32, int32, uint32) of n rows")
            pt, pp, pe = (C.c_void_p(a.ctypes.data) for a in (tile, pal, err))
        else:
            tile, pt = c.out((n_q,), np.int32)
            pal, pp = c.out((n_q,), np.int32)
            err, pe = c.out((n_q,), np.uint32)
        check(fn(self._h, c.inp(x, dtype), n_q, int(k), pt, pp, pe))
        return tile, pal, err

    def match_rgb(self, rgb, k=None, out=None):
        """Source tiles as RGB [n,64] -> (TileIdx, PalIdx, err) per tile.  out: optional host result arrays to fill (see _run)."""
        return self._run(_lib.lib().tm_match_tiles_rgb, rgb, 64, np.int32, k or (self.K_EPU if self.extended else 1), out=out)

    def match_rgb_mirrors(self, rgb, k=None):
        """Mirror-variant search (tm_match_tiles_rgb_mirrors). -> (tile_idx, pal_idx, err, variant); variant bit 0 / 1 = extra
        H / V mirror of the winning match."""
        c = _Call(rgb)
        n = _n_rows(rgb, 64)
        kk = int(k) if k else (self.K_EPU if self.extended else 1)
        tile, pt = c.out((n,), np.int32)
        pal, pp = c.out((n,), np.int32)
        err, pe = c.out((n,), np.uint32)
        var, pv = c.out((n,), np.uint8)
        check(_lib.lib().tm_match_tiles_rgb_mirrors(self._h, c.inp(rgb, np.int32), n, kk, pt, pp, pe, pv))
        return tile, pal, err, var

    def match_feat(self, feat, k=None):
        return self._run(_lib.lib().tm_match_tiles_feat, feat, DCT, np.int16, k or (self.K_EPU if self.extended else 1))

    def reconstruct_sequence(self, canon_tiles, flags, tw, th, radius=32, k=None, want_recon=True):
        """TTilingEncoder.Reconstruct over one keyframe sequence (tilingencoder.pas:1928-1962, 1430-1679).
        canon_tiles [n_frames, th*tw, 64] stored (canonicalised) tiles, flags [n_frames, th*tw].
        -> dict(tile_idx, pal_idx, pred_x, pred_y, is_pred, err, psnr, recon [n_frames, th*8, tw*8])."""
        c = _Call(canon_tiles, flags)
        n_frames = int(canon_tiles.shape[0])
        nt = tw * th
        shp = (n_frames, nt)
        tile, p_tile = c.out(shp, np.int32)
        pal, p_pal = c.out(shp, np.int32)
        px, p_px = c.out(shp, np.int32)
        py, p_py = c.out(shp, np.int32)
        isp, p_isp = c.out(shp, np.uint8)
        err, p_err = c.out(shp, np.uint32)
        psnr, p_psnr = c.out(shp, np.float32)
        recon, p_recon = c.out((n_frames, th * 8, tw * 8), np.int32) if want_recon else (None, None)
        check(_lib.lib().tm_reconstruct_sequence(self._h, c.inp(canon_tiles, np.int32), c.inp(flags, np.uint8), n_frames, int(tw), int(th),
                                                 int(radius), int(k or (self.K_EPU if self.extended else 1)), p_tile, p_pal, p_px, p_py,
                                                 p_isp, p_err, p_psnr, p_recon))
        return {"tile_idx": tile, "pal_idx": pal, "pred_x": px, "pred_y": py, "is_pred": isp, "err": err, "psnr": psnr, "recon": recon}

    def reconstruct_frame(self, canon_tiles, flags, tw, th, back=None, radius=32, k=None):
        """One frame of TFrame.Reconstruct (tilingencoder.pas:1430-1679): back = previous reconstructed frame [th*8, tw*8] or
        None on the first frame of a keyframe sequence.  -> dict like reconstruct_sequence, recon = this frame [th*8, tw*8]."""
        c = _Call(canon_tiles, flags) if back is None else _Call(canon_tiles, flags, back)
        nt = tw * th
        tile, p_tile = c.out((nt,), np.int32)
        pal, p_pal = c.out((nt,), np.int32)
        px, p_px = c.out((nt,), np.int32)
        py, p_py = c.out((nt,), np.int32)
        isp, p_isp = c.out((nt,), np.uint8)
        err, p_err = c.out((nt,), np.uint32)
        psnr, p_psnr = c.out((nt,), np.float32)
        front, p_front = c.out((th * 8, tw * 8), np.int32)
        check(_lib.lib().tm_reconstruct_frame(self._h, c.inp(canon_tiles, np.int32), c.inp(flags, np.uint8), int(tw), int(th), int(radius),
                                              int(k or (self.K_EPU if self.extended else 1)), c.inp(back, np.int32), p_front, p_tile, p_pal,
                                              p_px, p_py, p_isp, p_err, p_psnr))
        return {"tile_idx": tile, "pal_idx": pal, "pred_x": px, "pred_y": py, "is_pred": isp, "err": err, "psnr": psnr, "recon": front}

    def dict_features(self):
        out = np.empty((self.n_dict, DCT), dtype=np.int16)
        check(_lib.lib().tm_matcher_dict_features(self._h, C.c_void_p(out.ctypes.data)))
        return out

    def close(self):
        if self._h:
            _lib.lib().tm_matcher_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------ motion search (tilingencoder.pas:1154-1282, 1496-1532)
def sliding_features(frame):
    """DoDCTs: frame buffer [h, w] packed RGB -> features of every 8x8 window, int16 [(h-7)*(w-7), 192]."""
    c = _Call(frame)
    h, w = int(frame.shape[0]), int(frame.shape[1])
    out, po = c.out(((h - 7) * (w - 7), DCT), np.int16)
    check(_lib.lib().tm_sliding_features(c.inp(frame, np.int32), w, h, po))
    return out


def motion_search(cur_feat, tw, th, dcts, radius=32):
    """Window scan of DoXY -> (pred_x, pred_y, err) per tile; err carries the Manhattan penalty."""
    c = _Call(cur_feat, dcts)
    nt = tw * th
    px, ppx = c.out((nt,), np.int32)
    py, ppy = c.out((nt,), np.int32)
    err, pe = c.out((nt,), np.uint32)
    check(_lib.lib().tm_motion_search(c.inp(cur_feat, np.int16), int(tw), int(th), c.inp(dcts, np.int16), int(radius), ppx, ppy, pe))
    return px, py, err


def predict_motion_frame(prev_frame, canon_tiles, flags, tw, th, radius=32):
    """One frame of TTilingEncoder.PredictMotion: previous frame pixels [th*8, tw*8] + this frame's stored tiles."""
    c = _Call(prev_frame, canon_tiles, flags)
    nt = tw * th
    px, ppx = c.out((nt,), np.int32)
    py, ppy = c.out((nt,), np.int32)
    err, pe = c.out((nt,), np.uint32)
    check(_lib.lib().tm_predict_motion_frame(c.inp(prev_frame, np.int32), c.inp(canon_tiles, np.int32), c.inp(flags, np.uint8), int(tw),
                                             int(th), int(radius), ppx, ppy, pe))
    return px, py, err


def tile_classes(rgb):
    """Exact duplicate classes of RGB tiles [n,64] (MakeTilesUnique(True), tilingencoder.pas:4720-4781) -> (class_id [n], n_classes)."""
    c = _Call(rgb)
    n = _n_rows(rgb, 64)
    cls, pc = c.out((n,), np.int32)
    cnt = C.c_int()
    check(_lib.lib().tm_tile_classes(c.inp(rgb, np.int32), n, pc, C.byref(cnt)))
    return cls, cnt.value


def reduce_class_min(class_id, eff_psnr, n_classes):
    """Per duplicate class the smallest effective PSNR of its members, sorted ascending (STCGREval's count for any threshold)."""
    c = _Call(class_id, eff_psnr)
    n = int(np.prod(class_id.shape))
    out, po = c.out((int(n_classes),), np.float64)
    check(_lib.lib().tm_reduce_class_min(c.inp(class_id, np.int32), c.inp(eff_psnr, np.float64), n, int(n_classes), po))
    return out


def reduce_apply(class_id, eff_psnr, n_classes, x):
    """IsPredicted := PSNR > x for every tile -> (use_count [n_classes], first unpredicted member [n_classes], unpredicted [n])."""
    c = _Call(class_id, eff_psnr)
    n = int(np.prod(class_id.shape))
    use, pu = c.out((int(n_classes),), np.int32)
    rep, pr = c.out((int(n_classes),), np.int32)
    unp, pn = c.out((n,), np.uint8)
    check(_lib.lib().tm_reduce_apply(c.inp(class_id, np.int32), c.inp(eff_psnr, np.float64), n, int(n_classes), float(x), pu, pr, pn))
    return use, rep, unp


def reduce_remap(class_id, unpredicted, new_of_class):
    """Tilemap TileIdx after TransferTiles + ReindexTiles: new_of_class[class] where unpredicted, else -1."""
    c = _Call(class_id, unpredicted, new_of_class)
    n = int(np.prod(class_id.shape))
    out, po = c.out((n,), np.int32)
    check(_lib.lib().tm_reduce_remap(c.inp(class_id, np.int32), c.inp(unpredicted, np.uint8), c.inp(new_of_class, np.int32), n,
                                     int(np.prod(new_of_class.shape)), po))
    return out


def mse_rgb(a, b):
    c = _Call(a, b)
    n = int(np.prod(a.shape))
    out = C.c_double()
    check(_lib.lib().tm_mse_rgb(c.inp(a, np.int32), c.inp(b, np.int32), n, C.byref(out)))
    return out.value


# ------------------------------------------------------------------ drop-in symbols, exercised the way extern.pas binds them
class AnnKdTreeShort:
    """ann_kdtree_short_create / _search / _search_multi / _destroy with the Pascal calling pattern (row pointers)."""

    def __init__(self, rows, bucket=32, split=0):
        self._rows = np.ascontiguousarray(rows, dtype=np.int16)
        n, dim = self._rows.shape
        ptrs = (C.c_void_p * n)(*[self._rows[i].ctypes.data for i in range(n)])
        self._h = _lib.lib().ann_kdtree_short_create(ptrs, n, dim, bucket, split)
        if not self._h:
            raise TmError(-1, _lib.lib().tm_last_error().decode())

    def search(self, q):
        q = np.ascontiguousarray(q, dtype=np.int16)
        err = C.c_uint32()
        idx = _lib.lib().ann_kdtree_short_search(self._h, C.c_void_p(q.ctypes.data), 0, C.byref(err))
        return int(idx), int(err.value)

    def search_multi(self, q, k):
        q = np.ascontiguousarray(q, dtype=np.int16)
        idxs = np.empty(k, dtype=np.int32)
        errs = np.empty(k, dtype=np.uint32)
        _lib.lib().ann_kdtree_short_search_multi(self._h, C.c_void_p(idxs.ctypes.data), C.c_void_p(errs.ctypes.data), k,
                                                 C.c_void_p(q.ctypes.data), 0)
        return idxs, errs

    def rendezvous_stats(self):
        """(queries answered, batched launches) of the per-query searches on this handle."""
        q, b = C.c_int64(), C.c_int64()
        check(_lib.lib().tm_rendezvous_stats(self._h, C.byref(q), C.byref(b)))
        return q.value, b.value

    def destroy(self):
        if self._h:
            _lib.lib().ann_kdtree_short_destroy(self._h)
            self._h = None


class AnnKdTree:
    """ann_kdtree_create / _search / _destroy (ANN.dll, doubles)."""

    def __init__(self, rows, bucket=32, split=0):
        self._rows = np.ascontiguousarray(rows, dtype=np.float64)
        n, dim = self._rows.shape
        ptrs = (C.c_void_p * n)(*[self._rows[i].ctypes.data for i in range(n)])
        self._h = _lib.lib().ann_kdtree_create(ptrs, n, dim, bucket, split)
        if not self._h:
            raise TmError(-1, _lib.lib().tm_last_error().decode())

    def search(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        err = C.c_double()
        idx = _lib.lib().ann_kdtree_search(self._h, C.c_void_p(q.ctypes.data), 0.0, C.byref(err))
        return int(idx), float(err.value)

    def destroy(self):
        if self._h:
            _lib.lib().ann_kdtree_destroy(self._h)
            self._h = None


class Yakmo:
    """yakmo_create / load_train_data / train_on_data / get_centroids / destroy (extern.pas:198-203)."""

    def __init__(self, k, restarts=1, max_iter=300, init_type=1, init_seed=0, normalize=0, verbose=0):
        self.k = k
        self._h = _lib.lib().yakmo_create(k, restarts, max_iter, init_type, init_seed, normalize, verbose)

    def load_train_data(self, data):
        data = np.ascontiguousarray(data, dtype=np.float64)
        self.rows, self.cols = data.shape
        ptrs = (C.c_void_p * self.rows)(*[data[i].ctypes.data for i in range(self.rows)])
        _lib.lib().yakmo_load_train_data(self._h, self.rows, self.cols, ptrs)

    def train_on_data(self):
        labels = np.full(self.rows, -1, dtype=np.int32)
        _lib.lib().yakmo_train_on_data(self._h, C.c_void_p(labels.ctypes.data))
        return labels

    def get_centroids(self):
        cent = np.zeros((self.k, self.cols), dtype=np.float64)
        ptrs = (C.c_void_p * self.k)(*[cent[i].ctypes.data for i in range(self.k)])
        _lib.lib().yakmo_get_centroids(self._h, ptrs)
        return cent

    def destroy(self):
        if self._h:
            _lib.lib().yakmo_destroy(self._h)
            self._h = None


class Bico:
    """bico_create / insert_line / get_results / destroy (extern.pas:218-223)."""

    def __init__(self, dim, n, k, nrandproj, coresetsize, seed):
        self.dim, self.coreset = dim, coresetsize
        self._h = _lib.lib().bico_create(dim, n, k, nrandproj, coresetsize, seed)

    def insert_line(self, row, weight):
        row = np.ascontiguousarray(row, dtype=np.float64)
        _lib.lib().bico_insert_line(self._h, C.c_void_p(row.ctypes.data), float(weight))

    def get_results(self):
        cent = np.zeros((self.coreset, self.dim), dtype=np.float64)
        w = np.zeros(self.coreset, dtype=np.float64)
        m = _lib.lib().bico_get_results(self._h, C.c_void_p(cent.ctypes.data), C.c_void_p(w.ctypes.data))
        return cent[:m], w[:m]

    def destroy(self):
        if self._h:
            _lib.lib().bico_destroy(self._h)
            self._h = None
