"""Generates the committed golden fixtures.  Run here (needs /root/reference for the dlquant part):
    python tests/golden/make_golden.py
dlquant_ref.npz  : outputs of the REFERENCE's own dl1quant/dl3quant (dlquant/quantizer.c compiled into oracle/_ref).
oracle_golden.npz: outputs of the CPU oracle on a seeded input set, frozen so that later edits to the oracle (or a
                   different libm / compiler on the GPU box) cannot silently move the parity target.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import oracle as O  # noqa: E402
from conftest import rand_tiles, rand_palettes  # noqa: E402

rng = np.random.default_rng(20240815)
img = rng.integers(0, 256, size=(48 * 48, 3), dtype=np.uint8)
img[: 48 * 16] //= 4
rc3, dl3 = O.ref_dl3quant(img, 48, 48, 16)
rc1, dl1 = O.ref_dl3quant(img, 48, 48, 16, which="dl1quant")
assert rc3 == 0 and rc1 == 0
np.savez_compressed(os.path.join(HERE, "dlquant_ref.npz"), img=img, dl3_16=dl3, dl1_16=dl1)

tiles = np.concatenate([rand_tiles(96, 1), rand_tiles(32, 2, smooth=False)])
flags = np.array([O.mirror_heuristics(t)[0] | (O.mirror_heuristics(t)[1] << 1) for t in tiles], dtype=np.uint8)
palettes = rand_palettes(4, 16, 3, n_null=2)
tile_pal = (np.arange(len(tiles)) % 4).astype(np.int32)
tk = O.dither(tiles, flags, tile_pal, palettes, use_tk=True)
yl = O.dither(tiles, flags, tile_pal, palettes, use_tk=False)
feat_rgb = O.features_from_rgb(tiles)
feat_pal = O.features_from_pal(tk, tile_pal, palettes)
idx, dist = O.knn_short(feat_pal, feat_rgb, 8)
np.savez_compressed(os.path.join(HERE, "oracle_golden.npz"), tiles=tiles, flags=flags, palettes=palettes, tile_pal=tile_pal,
                    pal_idx_tk=tk, pal_idx_yl=yl, feat_rgb=feat_rgb, feat_pal=feat_pal, knn_idx=idx, knn_dist=dist)
print("golden written")
