"""Generator of tests/golden/demo_tiles.npz: REAL dithered dictionary tiles and palettes out of the reference's own encoder
outputs docs/demo/city_cif.gtm and football_cif.gtm, for GPU-vs-oracle tests on realistic inputs.

Run in the build container (the reference checkout is not on the GPU box):  python tests/golden/make_demo_fixture.py
Per stream: the TileSet (tiles used more than once, tilingencoder.pas:5292-5316), every LoadPalette (:5270-5290), and for each
TileSet tile the palette of its first tilemap reference (commands :53-86); a strided subset of TILES_PER_STREAM tiles is kept.
"""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tiler_b200 import gtm  # noqa: E402  (host-side LZMA decoder, pinned by the same streams in tests/test_gtm_io.py)

DEMO = "/root/reference/docs/demo"
TILES_PER_STREAM = 20000


def parse_stream(path):
    d = open(path, "rb").read()
    hdr = gtm.parse_header(d)
    off = hdr["data_offset"]
    tiles, palettes, first_pal, pal_size = None, {}, None, None
    for kf in hdr["keyframes"]:
        raw, used = gtm.lzma_decode(d, off, max_out=kf["raw_size"] + 16)
        off += used
        p = 0
        while p < len(raw):
            (w,) = struct.unpack_from("<H", raw, p); p += 2
            op, arg = w & 15, w >> 4
            if op == 15:                                        # ExtendedCommand: uint32 length + bytes
                (n,) = struct.unpack_from("<I", raw, p); p += 4 + n
            elif op == 14:                                      # SetDimensions
                p += 12
            elif op == 13:                                      # TileSet: payload = colours per palette
                pal_size = arg
                s, e = struct.unpack_from("<II", raw, p); p += 8
                cnt = e - s + 1
                t = np.frombuffer(raw, np.uint8, cnt * 64, p).reshape(cnt, 64); p += cnt * 64
                assert s == 0 and tiles is None
                tiles = t.copy()
                first_pal = np.full(cnt, -1, np.int32)
            elif op == 12:                                      # LoadPalette: index + RGBA per colour
                (pi,) = struct.unpack_from("<H", raw, p); p += 2
                rgba = np.frombuffer(raw, np.uint8, pal_size * 4, p).reshape(pal_size, 4); p += pal_size * 4
                palettes[pi] = (rgba[:, 0].astype(np.int32) | (rgba[:, 1].astype(np.int32) << 8) | (rgba[:, 2].astype(np.int32) << 16))
            elif op == 11 or op == 6 or op == 0:                # FrameEnd / SkipBlock / PredictedTileShortOffsets
                pass
            elif op == 1:
                p += 2
            elif op in (2, 3):                                  # Short/LongTileIdx + ShortPalIdx
                (ti,) = struct.unpack_from("<H" if op == 2 else "<I", raw, p); p += 2 if op == 2 else 4
                if ti < len(first_pal) and first_pal[ti] < 0:
                    first_pal[ti] = arg >> 2
            elif op == 4:
                pi, ti = struct.unpack_from("<HI", raw, p); p += 6
                if ti < len(first_pal) and first_pal[ti] < 0:
                    first_pal[ti] = pi
            elif op == 5:                                       # IntraTile
                p += 2 + 64
            else:
                raise ValueError(f"unknown command {op}")
    pal = np.stack([palettes[i] for i in range(len(palettes))])
    used = first_pal >= 0
    return tiles[used], first_pal[used], pal


def main():
    out = {}
    for name in ("city_cif", "football_cif"):
        tiles, tpal, pal = parse_stream(os.path.join(DEMO, name + ".gtm"))
        sel = np.linspace(0, len(tiles) - 1, min(TILES_PER_STREAM, len(tiles))).astype(np.int64)
        out[name + "_tiles"], out[name + "_tile_pal"], out[name + "_palettes"] = tiles[sel], tpal[sel], pal
        print(name, "TileSet tiles with a reference:", len(tiles), "kept:", len(sel), "palettes:", pal.shape, "max index:", int(tiles.max()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "demo_tiles.npz"), **out)


if __name__ == "__main__":
    main()
