"""ctypes loader for libtm_gpu.so (the C ABI declared in include/tm_gpu.h).

The library is built in-tree by tiler_b200/csrc/Makefile (see __graft_entry__.build()).  There is no Python or CPU
implementation behind these bindings: if the shared object is missing, or no sm_100 device is present, calls fail
loudly (TmError) instead of falling back.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TM_LIB_PATH") or os.path.join(_HERE, "libtm_gpu.so")   # TM_LIB_PATH: another build of the same library

TM_OK = 0
ERR_NAMES = {1: "TM_ERR_ARG", 2: "TM_ERR_CUDA", 3: "TM_ERR_DRIVER", 4: "TM_ERR_NOGPU", 5: "TM_ERR_NOMEM"}


class TmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


_lib = None

_vp, _i32, _i64, _u64, _dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double

# name -> (restype, argtypes); every tm_* compute entry point returns an int status
SIGNATURES = {
    "tm_version": (C.c_int, []),
    "tm_device_count": (C.c_int, []),
    "tm_set_device": (C.c_int, [_i32]),
    "tm_set_stream": (C.c_int, [_vp]),
    "tm_last_error": (C.c_char_p, []),
    "tm_kernel_launches": (C.c_int64, []),
    "tm_synchronize": (C.c_int, []),
    "tm_set_feature_mode": (C.c_int, [_i32]),
    "tm_get_feature_mode": (C.c_int, []),
    "tm_profile_enable": (C.c_int, [_i32]),
    "tm_profile_read": (C.c_int, [C.c_char_p, C.POINTER(_dbl), C.POINTER(_i64)]),
    "tm_features_from_rgb": (C.c_int, [_vp, _i64, _vp]),
    "tm_features_from_pal": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i64, _vp]),
    "tm_features_f64": (C.c_int, [_vp, _i64, _i32, _i32, _vp]),
    "tm_mirror_canonicalise": (C.c_int, [_vp, _i64, _vp]),
    "tm_distance_pairs": (C.c_int, [_vp, _vp, _i64, _vp]),
    "tm_knn_short_create": (C.c_int, [_vp, _i64, C.POINTER(_vp)]),
    "tm_knn_short_destroy": (C.c_int, [_vp]),
    "tm_knn_short_batch": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _i32]),
    "tm_knn_double_batch": (C.c_int, [_vp, _i64, _i32, _vp, _i64, _vp, _vp]),
    "tm_dither": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _vp, _i32, _i32, _i32, _i32, _vp]),
    "tm_kmeans_fit": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _u64, _i32, _vp, _vp, C.POINTER(_dbl), C.POINTER(_i32)]),
    "tm_kmeans_fit_i16": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _i32, _vp, _vp, C.POINTER(_dbl), C.POINTER(_i32), C.POINTER(_i64)]),
    "tm_kmeans_partial_step_i16": (C.c_int, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, C.POINTER(_i64), C.POINTER(_dbl)]),
    "tm_kmeans_i16_create": (C.c_int, [_vp, _i64, _i32, C.POINTER(_vp)]),
    "tm_kmeans_i16_destroy": (C.c_int, [_vp]),
    "tm_kmeans_i16_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tm_kmeans_partial_step": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, C.POINTER(_i64), C.POINTER(_dbl)]),
    "tm_kmeans_finish_step": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp]),
    "tm_coreset_weighted": (C.c_int, [_vp, _vp, _i64, _i32, _i64, _i32, _u64, _vp, _vp, C.POINTER(_i64)]),
    "tm_palquant_kmeans": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _u64, _vp, C.POINTER(_i32)]),
    "tm_dl3quant_batch": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "tm_dl1quant_batch": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "dl3quant": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp]),
    "dl1quant": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp]),
    "tm_matcher_create": (C.c_int, [_vp, _vp, _i64, _vp, _i32, _i32, _i32, C.POINTER(_vp)]),
    "tm_matcher_destroy": (C.c_int, [_vp]),
    "tm_match_tiles_rgb": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "tm_match_tiles_feat": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "tm_match_tiles_rgb_mirrors": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "tm_features_from_rgb_mirrored": (C.c_int, [_vp, _vp, _i64, _vp]),
    "tm_matcher_dict_features": (C.c_int, [_vp, _vp]),
    "tm_sliding_features": (C.c_int, [_vp, _i32, _i32, _vp]),
    "tm_motion_search": (C.c_int, [_vp, _i32, _i32, _vp, _i32, _vp, _vp, _vp]),
    "tm_predict_motion_frame": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "tm_reconstruct_sequence": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tm_reconstruct_frame": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tm_mse_rgb": (C.c_int, [_vp, _vp, _i64, C.POINTER(_dbl)]),
    "tm_tile_classes": (C.c_int, [_vp, _i64, _vp, C.POINTER(_i32)]),
    "tm_reduce_class_min": (C.c_int, [_vp, _vp, _i64, _i64, _vp]),
    "tm_reduce_apply": (C.c_int, [_vp, _vp, _i64, _i64, _dbl, _vp, _vp, _vp]),
    "tm_reduce_remap": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp]),
    # drop-in exports (extern.pas:178-223)
    "ann_kdtree_short_create": (_vp, [_vp, _i32, _i32, _i32, _i32]),
    "ann_kdtree_short_destroy": (None, [_vp]),
    "ann_kdtree_short_search": (C.c_int, [_vp, _vp, C.c_uint32, C.POINTER(C.c_uint32)]),
    "ann_kdtree_short_search_multi": (None, [_vp, _vp, _vp, _i32, _vp, C.c_uint32]),
    "tm_rendezvous_stats": (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "ann_kdtree_create": (_vp, [_vp, _i32, _i32, _i32, _i32]),
    "ann_kdtree_destroy": (None, [_vp]),
    "ann_kdtree_search": (C.c_int, [_vp, _vp, _dbl, C.POINTER(_dbl)]),
    "yakmo_create": (_vp, [C.c_uint32, C.c_uint32, _i32, _i32, _i32, _i32, _i32]),
    "yakmo_destroy": (None, [_vp]),
    "yakmo_set_num_threads": (None, [_i32]),
    "yakmo_load_train_data": (None, [_vp, C.c_uint32, C.c_uint32, _vp]),
    "yakmo_train_on_data": (None, [_vp, _vp]),
    "yakmo_get_centroids": (None, [_vp, _vp]),
    "bico_create": (_vp, [_i64, _i64, _i64, _i64, _i64, _i32]),
    "bico_destroy": (None, [_vp]),
    "bico_set_num_threads": (None, [_i32]),
    "bico_set_rebuild_properties": (None, [_vp, C.c_uint32, _dbl, _dbl]),
    "bico_insert_line": (None, [_vp, _vp, _dbl]),
    "bico_get_results": (C.c_int64, [_vp, _vp, _vp]),
}


def lib():
    """Load libtm_gpu.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TmError(-1, f"{LIB_PATH} not built: run `make -C tiler_b200/csrc` (or __graft_entry__.build())")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != TM_OK:
        raise TmError(rc, lib().tm_last_error().decode("utf-8", "replace"))
