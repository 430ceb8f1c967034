"""CPU tests of the GTM verification I/O (tiler_b200/gtm.py over libtm_gtm.so): the LZMA codec of the stream's chunk format
(lc = 8, end marker; extern.pas:420-440), cross-checked against liblzma where liblzma can express the stream (lc <= 4), the
SaveStream serialiser and the player-equivalent decoder.  The reference's demo streams (docs/demo/*.gtm) are decoded when the
reference checkout is present; a digest of their first frames is committed in tests/golden/gtm_demo_digest.json."""
import hashlib
import json
import lzma
import os

import numpy as np
import pytest

from tiler_b200 import gtm

GOLD = os.path.join(os.path.dirname(__file__), "golden")
DEMO = "/root/reference/docs/demo"


def _samples():
    rng = np.random.default_rng(1)
    return [b"", b"a", b"abcabcabcabcabc" * 100, rng.integers(0, 256, 40000, dtype=np.uint8).tobytes(),
            rng.integers(0, 4, 120000, dtype=np.uint8).tobytes(), bytes(200000),
            (rng.integers(0, 16, 60000, dtype=np.uint8)).tobytes() + b"hello world " * 1000]


def test_lzma_roundtrip_gtm_parameters():
    for data in _samples():
        enc = gtm.lzma_encode(data)
        assert enc[0] == 0x62 and enc[1:5] == (1 << 22).to_bytes(4, "little") and enc[5:13] == b"\xff" * 8   # LZCompress header
        dec, used = gtm.lzma_decode(enc)
        assert dec == data and used == len(enc)


def test_lzma_against_liblzma():
    for data in _samples()[1:]:
        ours = gtm.lzma_encode(data, lc=3, lp=0, pb=2)
        assert lzma.decompress(ours, format=lzma.FORMAT_ALONE) == data          # liblzma reads our encoder's stream
        theirs = lzma.compress(data, format=lzma.FORMAT_ALONE)
        dec, used = gtm.lzma_decode(theirs)
        assert dec == data and used == len(theirs)                              # we read liblzma's stream
    filt = [{"id": lzma.FILTER_LZMA1, "lc": 0, "lp": 2, "pb": 0, "dict_size": 1 << 16}]
    data = _samples()[2]
    assert gtm.lzma_decode(lzma.compress(data, format=lzma.FORMAT_ALONE, filters=filt))[0] == data
    assert lzma.decompress(gtm.lzma_encode(data, lc=0, lp=2, pb=0, dict_size=1 << 16), format=lzma.FORMAT_ALONE) == data


def test_lzma_block_parallel_parse_is_thread_count_independent():
    """Inputs of several 256 KB parse blocks: 1, 3 and 8 parser threads give the same bytes, which decode to the input (our
    decoder; liblzma for lc = 3), with repeats and long matches across block boundaries in the data."""
    rng = np.random.default_rng(5)
    base = rng.integers(0, 16, 300000, dtype=np.uint8).tobytes()
    data = base + bytes(100000) + base[:250000] + rng.integers(0, 256, 200000, dtype=np.uint8).tobytes() + base[1000:180000] * 2
    assert len(data) > 4 * (1 << 18)
    one = gtm.lzma_encode(data, n_threads=1)
    assert gtm.lzma_encode(data, n_threads=3) == one and gtm.lzma_encode(data, n_threads=8) == one
    assert gtm.lzma_decode(one)[0] == data
    assert len(one) < 0.75 * len(data)                       # the repeated 250 KB and 179 KB spans are found across blocks
    l3 = gtm.lzma_encode(data, lc=3, lp=0, pb=2, n_threads=4)
    assert l3 == gtm.lzma_encode(data, lc=3, lp=0, pb=2, n_threads=1)
    assert lzma.decompress(l3, format=lzma.FORMAT_ALONE) == data
    out = np.empty(64, dtype=np.uint8)                        # a buffer that is too small is reported, never overrun
    src = np.frombuffer(data, dtype=np.uint8)
    assert gtm.lib().tmh_lzma_encode(src.ctypes.data, src.size, 8, 0, 2, 1 << 22, out.ctypes.data, out.size) == -1


def test_lzma_back_to_back_streams():
    both = gtm.lzma_encode(b"first" * 10) + gtm.lzma_encode(b"second" * 10)
    a, ua = gtm.lzma_decode(both)
    b, ub = gtm.lzma_decode(both, ua)
    assert a == b"first" * 10 and b == b"second" * 10 and ua + ub == len(both)


def _synthetic_tilemap(seed, n_frames=5, tw=12, th=9, n_tiles=300, n_pal=3, pal_size=16):
    rng = np.random.default_rng(seed)
    nt = tw * th
    tiles = rng.integers(0, pal_size, size=(n_tiles, 64)).astype(np.uint8)
    pal = rng.integers(0, 1 << 24, size=(n_pal, pal_size)).astype(np.int32)
    pal[1, -2:] = gtm.NULL_COLOR
    tm = {"tile_idx": rng.integers(0, n_tiles, size=(n_frames, nt)).astype(np.int32),
          "pal_idx": rng.integers(0, n_pal, size=(n_frames, nt)).astype(np.int32),
          "pred_x": np.zeros((n_frames, nt), np.int32), "pred_y": np.zeros((n_frames, nt), np.int32),
          "is_pred": np.zeros((n_frames, nt), np.uint8), "mirror": rng.integers(0, 4, size=(n_frames, nt)).astype(np.uint8)}
    # frames after the first of each sequence: predicted tiles (inner tiles only, so vectors stay inside the frame)
    for f in (1, 2, 4):
        for t in range(nt):
            y, x = divmod(t, tw)
            if 1 <= y < th - 1 and 1 <= x < tw - 1 and rng.random() < 0.6:
                tm["is_pred"][f, t] = 1
                if rng.random() < 0.5:
                    tm["pred_x"][f, t], tm["pred_y"][f, t] = rng.integers(-8, 8), rng.integers(-8, 8)
        tm["is_pred"][f, 13:22] = 1; tm["pred_x"][f, 13:22] = 0; tm["pred_y"][f, 13:22] = 0   # a SkipBlock run
    return tm, tiles, pal


def _render(tm, tiles, pal, tw, th):
    n_frames = tm["tile_idx"].shape[0]
    out = np.zeros((n_frames, th * 8, tw * 8), np.int32)
    palc = pal.copy()
    palc[palc == gtm.NULL_COLOR] = 0xFFFFFF
    for f in range(n_frames):
        for t in range(tw * th):
            y, x = (t // tw) * 8, (t % tw) * 8
            if tm["is_pred"][f, t]:
                sy, sx = y + tm["pred_y"][f, t], x + tm["pred_x"][f, t]
                out[f, y:y + 8, x:x + 8] = out[f - 1, sy:sy + 8, sx:sx + 8]
            else:
                px = tiles[tm["tile_idx"][f, t]].reshape(8, 8)
                if tm["mirror"][f, t] & 1: px = px[:, ::-1]
                if tm["mirror"][f, t] & 2: px = px[::-1, :]
                out[f, y:y + 8, x:x + 8] = palc[tm["pal_idx"][f, t]][px]
    return out


@pytest.mark.parametrize("skip", [True, False])
def test_gtm_write_decode_roundtrip(skip):
    tw, th = 12, 9
    tm, tiles, pal = _synthetic_tilemap(3)
    final_tiles, use, tmap = gtm.reindex(tiles, tm["tile_idx"])
    assert (np.diff(use) <= 0).all() and use.min() >= 1 and (use == 1).any() and (use > 1).any()
    assert np.array_equal(final_tiles[tmap], tiles[tm["tile_idx"]])              # remapping preserves tile content
    tm2 = dict(tm, tile_idx=tmap)
    data = gtm.write_gtm(None, tm2, final_tiles, use, pal, tw, th, [(0, 2), (3, 4)], fps=24.0, settings_text="x=1", emit_skip_blocks=skip)
    hdr = gtm.parse_header(data)
    assert (hdr["width"], hdr["height"], hdr["kf_count"], hdr["frame_count"], hdr["encoder_version"]) == (96, 72, 2, 5, 4)
    assert [k["frame_index"] for k in hdr["keyframes"]] == [0, 3] and hdr["keyframes"][1]["timecode_ms"] == 125
    assert hdr["data_offset"] + sum(k["compressed_size"] for k in hdr["keyframes"]) == len(data)
    frames, _ = gtm.decode_gtm(data)
    assert np.array_equal(frames, _render(tm, tiles, pal, tw, th))
    raw0, _ = gtm.lzma_decode(data, hdr["data_offset"])
    assert len(raw0) == hdr["keyframes"][0]["raw_size"]
    smaller = gtm.write_gtm(None, tm2, final_tiles, use, pal, tw, th, [(0, 2), (3, 4)], emit_skip_blocks=True)
    bigger = gtm.write_gtm(None, tm2, final_tiles, use, pal, tw, th, [(0, 2), (3, 4)], emit_skip_blocks=False)
    assert gtm.parse_header(smaller)["keyframes"][0]["raw_size"] < gtm.parse_header(bigger)["keyframes"][0]["raw_size"]


def test_gtm_long_forms_and_long_offsets():
    tw, th = 10, 10
    rng = np.random.default_rng(5)
    n_tiles = 70000                                     # forces LongTileIdx forms (tile index > 65535)
    tiles = rng.integers(0, 16, size=(n_tiles, 64)).astype(np.uint8)
    pal = rng.integers(0, 1 << 24, size=(1100, 16)).astype(np.int32)   # palette index >= 1024 forces LongPalIdx
    nt = tw * th
    tm = {"tile_idx": np.tile(np.arange(n_tiles - nt, n_tiles, dtype=np.int32), (2, 1)),
          "pal_idx": np.tile((np.arange(nt) * 11 % 1100).astype(np.int32), (2, 1)),
          "pred_x": np.zeros((2, nt), np.int32), "pred_y": np.zeros((2, nt), np.int32),
          "is_pred": np.zeros((2, nt), np.uint8), "mirror": np.zeros((2, nt), np.uint8)}
    tm["is_pred"][1, 55] = 1; tm["pred_x"][1, 55] = -40; tm["pred_y"][1, 55] = -33     # outside -32..31 -> long offsets
    use = np.full(n_tiles, 2, np.int32); use[-1] = 1
    data = gtm.write_gtm(None, tm, tiles, use, pal, tw, th, [(0, 1)])
    frames, _ = gtm.decode_gtm(data)
    assert np.array_equal(frames, _render(tm, tiles, pal, tw, th))


@pytest.mark.skipif(not os.path.exists(DEMO), reason="reference checkout not present (GPU box)")
def test_reference_demo_streams_decode():
    want = json.load(open(os.path.join(GOLD, "gtm_demo_digest.json")))
    for name, rec in want.items():
        d = open(os.path.join(DEMO, name), "rb").read()
        hdr = gtm.parse_header(d)
        assert [hdr[k] for k in ("width", "height", "kf_count", "frame_count")] == rec["header"]
        off = hdr["data_offset"]
        for kf in hdr["keyframes"]:
            raw, used = gtm.lzma_decode(d, off, max_out=kf["raw_size"] + 16)
            assert len(raw) == kf["raw_size"] and used == kf["compressed_size"]   # sizes the reference's own writer recorded
            off += used
        assert off == len(d)
        frames, _ = gtm.decode_gtm(d, max_frames=rec["frames"])
        assert hashlib.sha256(frames.tobytes()).hexdigest() == rec["sha256"]
