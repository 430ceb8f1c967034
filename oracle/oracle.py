"""ctypes binding of the CPU oracle (oracle/liboracle.so) and of the reference's own dlquant (oracle/_ref).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under tiler_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

DCT = 192
NULL_COLOR = np.int32(-65281)  # 0xffff00ff as int32 (cDitheringNullColor, utils.pas:45)

PVS_DCT, PVS_WEIGHTED_DCT, PVS_WAVELETS, PVS_SPE_DCT, PVS_WEIGHTED_SPE_DCT = range(5)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "tm_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    ref = os.path.join(_HERE, "_ref", "libdlquant_ref.so")
    if (force or not os.path.exists(ref)) and os.path.exists("/root/reference/dlquant/quantizer.c"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def lib():
    global _LIB
    if _LIB is None:
        build()
        _LIB = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        L = _LIB
        L.tmo_dithering_map.restype = C.POINTER(C.c_uint8)
        L.tmo_dct_snake.restype = C.POINTER(C.c_uint8)
        L.tmo_dct_weights.restype = C.POINTER(C.c_double)
        L.tmo_dct_lut_f32.restype = C.POINTER(C.c_float)
        L.tmo_dct_lut_f64.restype = C.POINTER(C.c_double)
        L.tmo_vec_inv.restype = C.POINTER(C.c_uint32)
        L.tmo_compare_euclidean_dct.restype = C.c_uint32
        L.tmo_compare_euclidean_dct_sse.restype = C.c_uint32
        L.tmo_euclidean_to_psnr.restype = C.c_float
        L.tmo_color_compare.restype = C.c_int64
        L.tmo_color_compare.argtypes = [C.c_int64] * 6
        L.tmo_yuv_to_rgb.argtypes = [C.c_float] * 3
        L.tmo_lab_to_rgb.argtypes = [C.c_float] * 3
        L.tmo_kmeans_lloyd.restype = C.c_int
    return _LIB


def ref_dlquant():
    """The reference's own dlquant (dlquant/quantizer.c) compiled into oracle/_ref; None if not built."""
    global _REF
    if _REF is None:
        build()
        p = os.path.join(_HERE, "_ref", "libdlquant_ref.so")
        if not os.path.exists(p):
            return None
        _REF = C.CDLL(p)
    return _REF


def set_num_threads(n):
    lib().tmo_set_num_threads(int(n))


def num_threads():
    return int(lib().tmo_num_threads())


# ---- tables ----
def dithering_map():
    return np.ctypeslib.as_array(lib().tmo_dithering_map(), (64,)).copy()


def dct_snake():
    return np.ctypeslib.as_array(lib().tmo_dct_snake(), (64,)).copy()


def dct_weights():
    return np.ctypeslib.as_array(lib().tmo_dct_weights(), (3, 8, 8)).copy()


def dct_lut_f32(special=False):
    return np.ctypeslib.as_array(lib().tmo_dct_lut_f32(int(special)), (4096,)).copy()


def dct_lut_f64(special=False):
    return np.ctypeslib.as_array(lib().tmo_dct_lut_f64(int(special)), (4096,)).copy()


def vec_inv():
    return np.ctypeslib.as_array(lib().tmo_vec_inv(), (1024,)).copy()


# ---- colour ----
def rgb_to_yuv(r, g, b):
    y, u, v = C.c_float(), C.c_float(), C.c_float()
    lib().tmo_rgb_to_yuv(int(r), int(g), int(b), C.byref(y), C.byref(u), C.byref(v))
    return y.value, u.value, v.value


def yuv_to_rgb(y, u, v):
    return int(lib().tmo_yuv_to_rgb(C.c_float(y), C.c_float(u), C.c_float(v)))


def rgb_to_lab(r, g, b):
    y, u, v = C.c_float(), C.c_float(), C.c_float()
    lib().tmo_rgb_to_lab(int(r), int(g), int(b), C.byref(y), C.byref(u), C.byref(v))
    return y.value, u.value, v.value


def lab_to_rgb(l, a, b):
    return int(lib().tmo_lab_to_rgb(C.c_float(l), C.c_float(a), C.c_float(b)))


def rgb_to_hsv(col):
    h, s, v = C.c_uint8(), C.c_uint8(), C.c_uint8()
    lib().tmo_rgb_to_hsv(C.c_int32(int(col)), C.byref(h), C.byref(s), C.byref(v))
    return h.value, s.value, v.value


# ---- features ----
def features_from_rgb(rgb):
    rgb = np.ascontiguousarray(rgb, dtype=np.int32).reshape(-1, 64)
    out = np.empty((rgb.shape[0], DCT), dtype=np.int16)
    lib().tmo_features_from_rgb_batch(_p(rgb, C.c_int32), C.c_int64(rgb.shape[0]), _p(out, C.c_int16))
    return out


def features_from_pal(pal_idx, tile_pal, palettes):
    pal_idx = np.ascontiguousarray(pal_idx, dtype=np.uint8).reshape(-1, 64)
    tile_pal = np.ascontiguousarray(tile_pal, dtype=np.int32)
    palettes = np.ascontiguousarray(palettes, dtype=np.int32)
    out = np.empty((pal_idx.shape[0], DCT), dtype=np.int16)
    lib().tmo_features_from_pal_batch(_p(pal_idx, C.c_uint8), _p(tile_pal, C.c_int32), _p(palettes, C.c_int32),
                                      int(palettes.shape[1]), C.c_int64(pal_idx.shape[0]), _p(out, C.c_int16))
    return out


def tile_features_i16(rgb=None, pal_idx=None, palette=None, hmirror=False, vmirror=False):
    out = np.empty(DCT, dtype=np.int16)
    from_pal = rgb is None
    a = None if from_pal else np.ascontiguousarray(rgb, dtype=np.int32)
    b = np.ascontiguousarray(pal_idx, dtype=np.uint8) if from_pal else None
    c = np.ascontiguousarray(palette, dtype=np.int32) if from_pal else None
    lib().tmo_tile_features_i16(_p(a, C.c_int32), _p(b, C.c_uint8), _p(c, C.c_int32), int(from_pal),
                                int(hmirror), int(vmirror), _p(out, C.c_int16))
    return out


def tile_features_f64(rgb, mode, use_lab, hmirror=False, vmirror=False):
    out = np.empty(DCT, dtype=np.float64)
    a = np.ascontiguousarray(rgb, dtype=np.int32)
    lib().tmo_tile_features_f64(_p(a, C.c_int32), None, None, int(mode), 0, int(use_lab), int(hmirror), int(vmirror),
                                _p(out, C.c_double))
    return out


def inv_tile_features_f64(dct, mode, use_lab):
    out = np.empty(64, dtype=np.int32)
    d = np.ascontiguousarray(dct, dtype=np.float64)
    lib().tmo_inv_tile_features_f64(_p(d, C.c_double), int(mode), int(use_lab), _p(out, C.c_int32))
    return out


def mirror_heuristics(rgb):
    a = np.ascontiguousarray(rgb, dtype=np.int32)
    h, v = C.c_int(), C.c_int()
    lib().tmo_mirror_heuristics(_p(a, C.c_int32), C.byref(h), C.byref(v))
    return bool(h.value), bool(v.value)


# ---- distance / k-NN ----
def compare_euclidean_dct(a, b, sse=False):
    a = np.ascontiguousarray(a, dtype=np.int16)
    b = np.ascontiguousarray(b, dtype=np.int16)
    f = lib().tmo_compare_euclidean_dct_sse if sse else lib().tmo_compare_euclidean_dct
    return int(f(_p(a, C.c_int16), _p(b, C.c_int16)))


def euclidean_to_psnr(d):
    return float(lib().tmo_euclidean_to_psnr(C.c_uint32(int(d))))


def knn_short(dict_feat, q, k, sse=False):
    d = np.ascontiguousarray(dict_feat, dtype=np.int16).reshape(-1, DCT)
    q = np.ascontiguousarray(q, dtype=np.int16).reshape(-1, DCT)
    idx = np.empty((q.shape[0], k), dtype=np.int32)
    dist = np.empty((q.shape[0], k), dtype=np.uint32)
    lib().tmo_knn_short(_p(d, C.c_int16), C.c_int64(d.shape[0]), _p(q, C.c_int16), C.c_int64(q.shape[0]), int(k),
                        _p(idx, C.c_int32), _p(dist, C.c_uint32), int(sse))
    return idx, dist


def knn_double(dict_pts, q):
    d = np.ascontiguousarray(dict_pts, dtype=np.float64)
    q = np.ascontiguousarray(q, dtype=np.float64)
    idx = np.empty(q.shape[0], dtype=np.int32)
    dist = np.empty(q.shape[0], dtype=np.float64)
    lib().tmo_knn_double(_p(d, C.c_double), C.c_int64(d.shape[0]), int(d.shape[1]), _p(q, C.c_double),
                         C.c_int64(q.shape[0]), _p(idx, C.c_int32), _p(dist, C.c_double))
    return idx, dist


# ---- dithering ----
def dither(rgb, mirror_flags, tile_pal, palettes, use_tk=True, y2_mixed_colors=4, pair_tile=None):
    rgb = np.ascontiguousarray(rgb, dtype=np.int32).reshape(-1, 64)
    palettes = np.ascontiguousarray(palettes, dtype=np.int32)
    tile_pal = np.ascontiguousarray(tile_pal, dtype=np.int32)
    mf = None if mirror_flags is None else np.ascontiguousarray(mirror_flags, dtype=np.uint8)
    pt = None if pair_tile is None else np.ascontiguousarray(pair_tile, dtype=np.int32)
    n = tile_pal.shape[0]
    out = np.empty((n, 64), dtype=np.uint8)
    lib().tmo_dither_batch(_p(rgb, C.c_int32), _p(mf, C.c_uint8), _p(tile_pal, C.c_int32), C.c_int64(n),
                           _p(pt, C.c_int32), _p(palettes, C.c_int32), int(palettes.shape[1]), int(palettes.shape[0]),
                           int(use_tk), int(y2_mixed_colors), _p(out, C.c_uint8))
    return out


def color_compare(r1, g1, b1, r2, g2, b2):
    return int(lib().tmo_color_compare(r1, g1, b1, r2, g2, b2))


def quicksort_bytes_by_key(data, key):
    d = np.ascontiguousarray(data, dtype=np.uint8).copy()
    k = np.ascontiguousarray(key, dtype=np.int32)
    lib().tmo_quicksort_bytes_by_key(_p(d, C.c_uint8), C.c_int64(0), C.c_int64(len(d) - 1), _p(k, C.c_int32))
    return d


# ---- k-means ----
def kmeans_lloyd(x, init, max_iter=300, nan_empty=False):
    x = np.ascontiguousarray(x, dtype=np.float64)
    cent = np.ascontiguousarray(init, dtype=np.float64).copy()
    labels = np.empty(x.shape[0], dtype=np.int32)
    inertia = C.c_double()
    it = lib().tmo_kmeans_lloyd(_p(x, C.c_double), C.c_int64(x.shape[0]), int(x.shape[1]), int(cent.shape[0]),
                                int(max_iter), _p(cent, C.c_double), _p(labels, C.c_int32), C.byref(inertia),
                                int(nan_empty))
    return labels, cent, inertia.value, int(it)


def kmeanspp_init(x, k, seed):
    x = np.ascontiguousarray(x, dtype=np.float64)
    cent = np.empty((k, x.shape[1]), dtype=np.float64)
    lib().tmo_kmeanspp_init(_p(x, C.c_double), C.c_int64(x.shape[0]), int(x.shape[1]), int(k), C.c_uint64(seed),
                            _p(cent, C.c_double))
    return cent


def coreset_weighted(x, w, k, seed, max_iter=8):
    """tmo_coreset_weighted: the BICO stand-in.  -> (centroids [m, dim], weights [m])."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    cent = np.empty((int(k), x.shape[1]), dtype=np.float64)
    wts = np.empty(int(k), dtype=np.float64)
    f = lib().tmo_coreset_weighted
    f.restype = C.c_int64
    m = f(_p(x, C.c_double), _p(w, C.c_double), C.c_int64(x.shape[0]), int(x.shape[1]), C.c_int64(k), int(max_iter),
          C.c_uint64(seed), _p(cent, C.c_double), _p(wts, C.c_double))
    return cent[:m].copy(), wts[:m].copy()


def quantize_palette(pixels, pal_size, init=None, seed=1):
    px = np.ascontiguousarray(pixels, dtype=np.int32).reshape(-1)
    out = np.empty(pal_size, dtype=np.int32)
    ini = None if init is None else np.ascontiguousarray(init, dtype=np.float64)
    n = lib().tmo_quantize_palette(_p(px, C.c_int32), C.c_int64(px.shape[0]), int(pal_size), _p(ini, C.c_double),
                                   C.c_uint64(seed), _p(out, C.c_int32))
    return out, int(n)


# ---- matcher ----
class Match(C.Structure):
    _fields_ = [("tile_idx", C.c_int32), ("pal_idx", C.c_int32), ("err", C.c_uint32)]


def match_tiles(q_feat, dict_feat, dict_idx, dict_pal, palettes, k=64, extended=True):
    q = np.ascontiguousarray(q_feat, dtype=np.int16).reshape(-1, DCT)
    d = np.ascontiguousarray(dict_feat, dtype=np.int16).reshape(-1, DCT)
    di = np.ascontiguousarray(dict_idx, dtype=np.uint8).reshape(-1, 64)
    dp = np.ascontiguousarray(dict_pal, dtype=np.int32)
    pal = np.ascontiguousarray(palettes, dtype=np.int32)
    out = np.empty((q.shape[0], 3), dtype=np.int32)
    lib().tmo_match_tiles(_p(q, C.c_int16), C.c_int64(q.shape[0]), _p(d, C.c_int16), _p(di, C.c_uint8),
                          _p(dp, C.c_int32), C.c_int64(d.shape[0]), _p(pal, C.c_int32), int(pal.shape[1]),
                          int(pal.shape[0]), int(k), int(extended), out.ctypes.data_as(C.c_void_p))
    return out[:, 0].copy(), out[:, 1].copy(), out[:, 2].view(np.uint32).copy()


def sliding_features(frame):
    f = np.ascontiguousarray(frame, dtype=np.int32)
    h, w = f.shape
    out = np.empty(((h - 7) * (w - 7), DCT), dtype=np.int16)
    lib().tmo_sliding_features(_p(f, C.c_int32), int(w), int(h), _p(out, C.c_int16))
    return out


def motion_search(cur_feat, tw, th, dcts, radius=32):
    cf = np.ascontiguousarray(cur_feat, dtype=np.int16)
    d = np.ascontiguousarray(dcts, dtype=np.int16)
    nt = tw * th
    px, py, err = np.empty(nt, np.int32), np.empty(nt, np.int32), np.empty(nt, np.uint32)
    lib().tmo_motion_search(_p(cf, C.c_int16), int(tw), int(th), _p(d, C.c_int16), int(radius), _p(px, C.c_int32), _p(py, C.c_int32),
                            _p(err, C.c_uint32))
    return px, py, err


def features_from_rgb_mirrored(rgb, flags):
    """Features of stored tiles read through their mirror flags (ConvertToCpnPixels with AHMirror/AVMirror)."""
    t = np.ascontiguousarray(rgb, dtype=np.int32).reshape(-1, 64)
    fl = np.ascontiguousarray(flags, dtype=np.uint8).reshape(-1)
    out = np.empty((t.shape[0], DCT), dtype=np.int16)
    L = lib()
    for i in range(t.shape[0]):
        L.tmo_tile_features_i16(_p(t[i], C.c_int32), None, None, 0, int(fl[i] & 1), int((fl[i] >> 1) & 1),
                                out[i].ctypes.data_as(C.POINTER(C.c_int16)))
    return out


def reconstruct_sequence(canon_tiles, flags, tw, th, dict_feat, dict_idx, dict_pal, palettes, radius=32, extended=True):
    t = np.ascontiguousarray(canon_tiles, dtype=np.int32)
    n_frames = t.shape[0]
    nt = tw * th
    fl = np.ascontiguousarray(flags, dtype=np.uint8)
    d = np.ascontiguousarray(dict_feat, dtype=np.int16).reshape(-1, DCT)
    di = np.ascontiguousarray(dict_idx, dtype=np.uint8).reshape(-1, 64)
    dp = np.ascontiguousarray(dict_pal, dtype=np.int32)
    pal = np.ascontiguousarray(palettes, dtype=np.int32)
    shp = (n_frames, nt)
    r = {"tile_idx": np.empty(shp, np.int32), "pal_idx": np.empty(shp, np.int32), "pred_x": np.empty(shp, np.int32),
         "pred_y": np.empty(shp, np.int32), "is_pred": np.empty(shp, np.uint8), "err": np.empty(shp, np.uint32),
         "recon": np.empty((n_frames, th * 8, tw * 8), np.int32)}
    ps = C.c_double()
    lib().tmo_reconstruct_sequence(_p(t, C.c_int32), _p(fl, C.c_uint8), int(n_frames), int(tw), int(th), _p(d, C.c_int16),
                                   _p(di, C.c_uint8), _p(dp, C.c_int32), C.c_int64(d.shape[0]), _p(pal, C.c_int32),
                                   int(pal.shape[1]), int(pal.shape[0]), int(radius), int(extended), _p(r["tile_idx"], C.c_int32),
                                   _p(r["pal_idx"], C.c_int32), _p(r["pred_x"], C.c_int32), _p(r["pred_y"], C.c_int32),
                                   _p(r["is_pred"], C.c_uint8), _p(r["err"], C.c_uint32), _p(r["recon"], C.c_int32), C.byref(ps))
    r["psnr_sum"] = ps.value
    return r


# ---- reference dlquant (oracle/_ref) ----
def ref_dl3quant(rgb888, w, h, quant_to, lookup_bpc=5, which="dl3quant"):
    r = ref_dlquant()
    if r is None:
        raise RuntimeError("oracle/_ref/libdlquant_ref.so not built (needs /root/reference at build time)")
    buf = np.ascontiguousarray(rgb888, dtype=np.uint8).reshape(-1).copy()
    pal = np.zeros((3, 65536), dtype=np.uint8)
    rc = getattr(r, which)(_p(buf, C.c_uint8), int(w), int(h), int(quant_to), int(lookup_bpc), _p(pal, C.c_uint8))
    return rc, pal[:, :quant_to].T.copy()
