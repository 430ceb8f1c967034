"""GTM stream I/O for verification: writes what TTilingEncoder.SaveStream writes
(tilingencoder.pas:5177-5482) and decodes it the way LoadStream (:4880-5175) / gtm.player.js do, so that an encode can
be checked end to end (decoded-frame PSNR) without the FreePascal host or a browser.

Host-side CPU code over libtm_gtm.so (csrc/gtm_host.cpp).  It is not on the GPU product path: in the reference the
bitstream writer stays in the Pascal host and the decoder is the player.
"""
import ctypes as C
import os
import struct

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

GTM_LC, GTM_LP, GTM_PB, GTM_DICT = 8, 0, 2, 1 << 22      # LZCompress, extern.pas:420-440 (props byte 0x62, 4 MiB)
ENCODER_VERSION = 4                                        # tilingencoder.pas:5343
NULL_COLOR = -65281                                        # cDitheringNullColor


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libtm_gtm.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} not built: run `make -C tiler_b200/csrc`")
        L = C.CDLL(path)
        L.tmh_lzma_encode.restype = C.c_int64
        L.tmh_lzma_encode.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_int64]
        L.tmh_lzma_encode_mt.restype = C.c_int64
        L.tmh_lzma_encode_mt.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_int64, C.c_int]
        L.tmh_lzma_decode.restype = C.c_int64
        L.tmh_lzma_decode.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        L.tmh_gtm_write_frames.restype = C.c_int64
        L.tmh_gtm_write_frames.argtypes = [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                                              C.c_void_p, C.c_int64]
        L.tmh_gtm_decoder_create.restype = C.c_void_p
        L.tmh_gtm_decoder_create.argtypes = []
        L.tmh_gtm_decoder_destroy.argtypes = [C.c_void_p]
        L.tmh_gtm_decoder_dims.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64)]
        L.tmh_gtm_decode.restype = C.c_int64
        L.tmh_gtm_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        L.tmh_optimize_palettes.restype = C.c_int
        L.tmh_optimize_palettes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        _LIB = L
    return _LIB


# ------------------------------------------------------------------ LZMA ("alone" container, end marker)
def lzma_encode(data, lc=GTM_LC, lp=GTM_LP, pb=GTM_PB, dict_size=GTM_DICT, n_threads=0):
    """n_threads: parser threads inside this one stream (0 = library default); the bytes do not depend on it."""
    src = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, np.uint8)
    cap = int(src.size + src.size // 4 + 1024)
    out = np.empty(cap, dtype=np.uint8)
    n = lib().tmh_lzma_encode_mt(src.ctypes.data, src.size, lc, lp, pb, dict_size, out.ctypes.data, cap, int(n_threads))
    if n < 0:
        raise RuntimeError(f"tmh_lzma_encode failed ({n})")
    return out[:n].tobytes()


def lzma_decode(data, offset=0, max_out=None):
    """Decodes one stream starting at data[offset] -> (bytes, input bytes consumed)."""
    src = np.frombuffer(data, dtype=np.uint8)[offset:]
    cap = int(max_out or max(1 << 20, 16 * src.size))
    while True:
        out = np.empty(cap, dtype=np.uint8)
        used = C.c_int64()
        n = lib().tmh_lzma_decode(src.ctypes.data, src.size, out.ctypes.data, cap, C.byref(used))
        if n == -1 and max_out is None:
            cap *= 4
            continue
        if n < 0:
            raise RuntimeError(f"tmh_lzma_decode failed ({n})")
        return out[:n].tobytes(), used.value


# ------------------------------------------------------------------ OptimizePalettes (tilingencoder.pas:4246-4432, powell.pas)
def optimize_palettes(palettes, n_threads=0):
    """Reorders the colours inside each palette (host code in the reference too).  -> (new palettes [n_pal, pal_size], passes)."""
    pal = np.ascontiguousarray(palettes, dtype=np.int32).copy()
    it = lib().tmh_optimize_palettes(pal.ctypes.data, int(pal.shape[0]), int(pal.shape[1]), int(n_threads))
    if it < 0:
        raise ValueError("tmh_optimize_palettes: bad argument")
    return pal, int(it)


# ------------------------------------------------------------------ dictionary bookkeeping (Reindex, tilingencoder.pas:1992-2040)
def reindex(tiles_idx, tile_idx_map, tile_use=None):
    """MakeTilesUnique(False) + use counts + ReindexTiles(False) (tilingencoder.pas:2010-2037, 4720-4781, 4626-4696):
    merges dictionary tiles with identical palette indices, counts how often each is referenced by a tilemap item
    (TileIdx >= 0, predicted items included, as the reference does), drops unused tiles, orders by (use count descending,
    palette-index bytes ascending) and remaps the tilemap.  -> (tiles [n,64] uint8, use_count [n], remapped tile_idx).
    tile_use (optional): references per dictionary tile, already counted (the encoder counts them on the device while the
    tilemap is still there); without it they are counted here."""
    tiles_idx = np.ascontiguousarray(tiles_idx, dtype=np.uint8).reshape(-1, 64)
    tmap = np.ascontiguousarray(tile_idx_map, dtype=np.int32)
    uniq, inverse = _unique_rows64(tiles_idx)                            # lexicographic on bytes = CompareByte order
    flat = tmap.reshape(-1)
    if tile_use is None:
        tile_use = np.bincount(flat + 1, minlength=len(tiles_idx) + 1)[1:]
    use = np.bincount(inverse, weights=np.asarray(tile_use, dtype=np.float64), minlength=len(uniq)).astype(np.int64)
    keep = np.nonzero(use > 0)[0]
    order = keep[np.lexsort((keep, -use[keep]))]                         # keep is already in byte order
    new_of_cls = np.full(len(uniq), -1, dtype=np.int32)
    new_of_cls[order] = np.arange(len(order), dtype=np.int32)
    # one gather over the whole tilemap: index -1 (no tile) reads the extra last slot of the table
    new_of_tile = np.append(new_of_cls[inverse], np.int32(-1)).astype(np.int32)
    out_map = new_of_tile[flat].reshape(tmap.shape)
    tiles_out = np.ascontiguousarray(uniq[order])
    return tiles_out, use[order].astype(np.int32), out_map


def _unique_rows64(rows):
    """np.unique(rows, axis=0, return_inverse=True) for uint8 [n, 64] rows in byte-lexicographic order: a sort on the first 8
    bytes as one big-endian uint64, the 64-byte memcmp sort only for the rows that tie on it.  -> (unique rows, inverse)."""
    n = len(rows)
    if n == 0:
        return rows.reshape(0, 64), np.zeros(0, dtype=np.int64)
    key = rows[:, :8].copy().view(">u8").reshape(-1)
    o = np.argsort(key, kind="stable")
    ks = key[o]
    eq = ks[1:] == ks[:-1]
    tie = np.zeros(n, dtype=bool)
    tie[1:] |= eq
    tie[:-1] |= eq
    if tie.any():
        idx = o[tie]
        o[tie] = idx[np.argsort(rows[idx].view(np.dtype((np.void, 64))).reshape(-1), kind="stable")]
    srt = rows[o]
    new = np.ones(n, dtype=bool)
    new[1:] = (srt[1:] != srt[:-1]).any(axis=1)
    inverse = np.empty(n, dtype=np.int64)
    inverse[o] = np.cumsum(new) - 1
    return srt[new], inverse


# ------------------------------------------------------------------ writer (SaveStream)
def _cmd(code, data):
    return struct.pack("<H", ((data << 4) | code) & 0xFFFF)


def write_gtm(path_or_none, tm, tiles_idx, use_count, palettes, tw, th, sequences, fps=24.0, settings_text="",
              emit_skip_blocks=True, only=None, exchange=None):
    """tm: dict of per-frame arrays [n_frames, th*tw] (tile_idx, pal_idx, pred_x, pred_y, is_pred, mirror) with tile_idx
    already re-indexed; tiles_idx / use_count: the final dictionary; palettes [n_pal, pal_size] int32; sequences: list of
    (start_frame, end_frame) inclusive.  Returns the file bytes (and writes them when a path is given).
    Multi-process use: `only` = the sequence indices whose chunks THIS process serialises and compresses, `exchange` = a
    callable that all-gathers {sequence index: (chunk bytes, keyframe info)} dictionaries across the processes."""
    n_frames = int(tm["tile_idx"].shape[0])
    nt = tw * th
    L = lib()
    arr = {k: np.ascontiguousarray(tm[k], dtype=dt) for k, dt in (("tile_idx", np.int32), ("pal_idx", np.int32), ("pred_x", np.int32),
                                                                  ("pred_y", np.int32), ("is_pred", np.uint8), ("mirror", np.uint8))}
    tiles_idx = np.ascontiguousarray(tiles_idx, dtype=np.uint8).reshape(-1, 64)
    use_count = np.ascontiguousarray(use_count, dtype=np.int32)
    n_tiles = len(tiles_idx)
    pal_size = int(palettes.shape[1])
    # ---- first chunk preamble: settings, dimensions, tile set, palettes (:5331-5335, 5318-5329, 5292-5316, 5270-5290)
    pre = bytearray()
    txt = settings_text.encode("latin-1")
    pre += _cmd(15, 0) + struct.pack("<I", len(txt)) + txt
    pre += _cmd(14, 0) + struct.pack("<HHII", tw, th, int(round(1e9 / fps)), n_tiles)
    reused = 0
    ones = np.nonzero(use_count == 1)[0]
    if len(ones):
        reused = int(ones[0])                     # tiles are sorted by use count: everything before the first single-use tile
    if reused > 0:
        pre += _cmd(13, pal_size) + struct.pack("<II", 0, reused - 1) + tiles_idx[:reused].tobytes()
    for p in range(palettes.shape[0]):
        cols = np.asarray(palettes[p], dtype=np.int64).copy()
        cols[cols == NULL_COLOR] = 0xFFFFFF
        cols = (cols & 0xFFFFFF) | 0xFF000000
        pre += _cmd(12, 0) + struct.pack("<H", p) + cols.astype("<u4").tobytes()
    # ---- chunks: one LZMA stream per keyframe sequence, compressed in parallel (the C calls release the GIL)
    def one_chunk(job):
        ki, (f0, f1) = job
        nf = f1 - f0 + 1
        cap = nf * nt * 70 + 64
        buf = np.empty(cap, dtype=np.uint8)
        sl = slice(f0, f1 + 1)
        views = [np.ascontiguousarray(arr[k][sl]) for k in ("tile_idx", "pal_idx", "pred_x", "pred_y", "is_pred", "mirror")]
        n = L.tmh_gtm_write_frames(*[v.ctypes.data for v in views], nf, nt, tiles_idx.ctypes.data, use_count.ctypes.data, n_tiles,
                                   int(emit_skip_blocks), 1, buf.ctypes.data, cap)
        assert 0 < n <= cap
        raw = (bytes(pre) if ki == 0 else b"") + buf[:n].tobytes()
        # chunk 0 carries the tile set and is the long pole of Save: it gets the full set of parser threads, the other chunks
        # (compressed at the same time by the pool below) two each, so that its coder thread is not descheduled by theirs
        comp = lzma_encode(raw, n_threads=(0 if ki == 0 or n_jobs <= 2 else 2))
        return comp, (ki, f0, len(raw), len(comp), int(round(1000.0 * f0 / fps)))

    from concurrent.futures import ThreadPoolExecutor
    jobs = [(ki, sq) for ki, sq in enumerate(sequences) if only is None or ki in only]
    n_jobs = len(jobs)
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as pool:
        done = dict(zip([j[0] for j in jobs], pool.map(one_chunk, jobs)))
    if exchange is not None:
        done = exchange(done)
    chunks = [done[ki][0] for ki in range(len(sequences))]
    kf_info = [done[ki][1] for ki in range(len(sequences))]
    # ---- header (TGTMHeader, TGTMKeyFrameInfo: tilingencoder.pas:30-51, 5338-5370, 5470-5476)
    whole_header = 40 + 28 * len(sequences)
    total = sum(len(c) for c in chunks)
    avg_bps = int(round(total * fps / n_frames))
    max_bps = 0
    last = 0
    for ki, (f0, f1) in enumerate(sequences):
        cnt = f1 - last + 1
        last = f1 + 1
        if ki > 0 or len(sequences) == 1:
            max_bps = max(max_bps, int(round(len(chunks[ki]) * fps / cnt)))
    out = bytearray()
    out += b"GTMv" + struct.pack("<9I", 32, whole_header, ENCODER_VERSION, tw * 8, th * 8, len(sequences), n_frames, avg_bps, max_bps)
    for ki, f0, raw_n, comp_n, ms in kf_info:
        out += b"GTMk" + struct.pack("<6I", 20, ki, f0, raw_n, comp_n, ms)
    for c in chunks:
        out += c
    if path_or_none:
        with open(path_or_none, "wb") as fh:
            fh.write(out)
    return bytes(out)


# ------------------------------------------------------------------ reader / decoder (LoadStream, gtm.player.js)
def parse_header(data):
    if data[:4] != b"GTMv":
        raise ValueError("not a GTM stream")
    riff, whole, ver, pw, ph, kfc, fc, avg, mx = struct.unpack_from("<9I", data, 4)
    kfs = []
    off = 8 + riff
    for _ in range(kfc):
        if data[off:off + 4] != b"GTMk":
            raise ValueError("bad keyframe info")
        rs, ki, fi, raw_n, comp_n, ms = struct.unpack_from("<6I", data, off + 4)
        kfs.append({"kf_index": ki, "frame_index": fi, "raw_size": raw_n, "compressed_size": comp_n, "timecode_ms": ms})
        off += 8 + rs
    return {"whole_header_size": whole, "encoder_version": ver, "width": pw, "height": ph, "kf_count": kfc, "frame_count": fc,
            "avg_bytes_per_sec": avg, "kf_max_bytes_per_sec": mx, "keyframes": kfs, "data_offset": off}


def decode_gtm(data, max_frames=None):
    """-> (frames int32 [n, H, W] packed 0x00BBGGRR, header dict).  Chunks are decompressed back to back (each ends with its
    own end marker, wlzma.wrk.js:50-63) and played through one decoder state."""
    hdr = parse_header(data)
    L = lib()
    dec = L.tmh_gtm_decoder_create()
    try:
        off = hdr["data_offset"]
        n_total = hdr["frame_count"] if max_frames is None else min(max_frames, hdr["frame_count"])
        frames = np.zeros((n_total, hdr["height"], hdr["width"]), dtype=np.int32)
        done = 0
        ki = 0
        while off < len(data) and done < n_total:
            raw_hint = hdr["keyframes"][ki]["raw_size"] if ki < len(hdr["keyframes"]) else None
            raw, used = lzma_decode(data, off, max_out=None if not raw_hint else raw_hint + 16)
            off += used
            ki += 1
            rb = np.frombuffer(raw, dtype=np.uint8)
            room = n_total - done
            got = L.tmh_gtm_decode(dec, rb.ctypes.data, rb.size, frames[done:].ctypes.data, room)
            if got < 0:
                raise ValueError("malformed GTM command stream")
            done += min(got, room)
        return frames[:done], hdr
    finally:
        L.tmh_gtm_decoder_destroy(dec)
