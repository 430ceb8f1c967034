"""Stage drivers mirroring TTilingEncoder's call structure for the data-parallel core
(tilingencoder.pas:1843-1962): PreparePalettes -> Dither -> PrepareReconstruct -> Reconstruct, each one a handful of
batched calls into libtm_gpu.so instead of the per-tile DLL calls of the reference.

What is NOT here (kept by the FreePascal host, out of this tier's scope): video loading, keyframe
detection, motion prediction, the PSNR-threshold dictionary selection (TransferTiles), OptimizePalettes, reindexing
and the GTM bitstream writer.  `reduce()` below is a stand-in that picks dictionary tiles so the four stages can run
end to end on a synthetic clip.
"""
import numpy as np

from . import api

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


class TilingEncoder:
    """Names follow the reference's settings (LoadDefaultSettings, tilingencoder.pas:3817-3845)."""

    def __init__(self, palette_size=16, palette_count=16, dithering_mode=api.PVS_WEIGHTED_SPE_DCT,
                 dithering_use_thomas_knoll=True, dithering_yliluoma2_mixed_colors=4,
                 frame_tiling_extended_palette_usage=True, seed=0x42381337, device=None):
        self.palette_size = int(np.clip(palette_size, 2, 256))      # reference clamps to 2..64 (:2965); 256 = stress shape
        self.palette_count = int(np.clip(palette_count, 1, 65536))
        self.dithering_mode = dithering_mode
        self.use_tk = bool(dithering_use_thomas_knoll)
        self.y2_mixed = int(np.clip(dithering_yliluoma2_mixed_colors, 1, 16))
        self.extended = bool(frame_tiling_extended_palette_usage)
        self.seed = seed
        self.device = device  # None: host arrays through the ABI; a torch device: everything stays in HBM
        self.tiles = None       # dictionary tiles, RGB [n,64] (canonical orientation)
        self.tile_flags = None  # their initial mirrors
        self.tile_pal = None    # PalIdx_Initial
        self.palettes = None    # [palette_count, palette_size] int32
        self.tile_idx = None    # dithered palette indices [n,64]
        self.matcher = None

    def _to(self, x):
        if self.device is None or x is None:
            return x
        return x if api._is_dev(x) else torch.as_tensor(np.ascontiguousarray(x)).to(self.device)

    # --- Load (mirror canonicalisation only; tilingencoder.pas:1393-1411)
    def load_tiles(self, frame_tiles):
        """frame_tiles [n_frames, tiles_per_frame, 64] RGB -> canonicalised tiles + per-tile H/V flags."""
        ft = self._to(frame_tiles)
        shape = tuple(ft.shape)
        tiles, flags = api.mirror_canonicalise(ft.reshape(-1, 64))
        return tiles.reshape(shape), flags.reshape(shape[:2])

    # --- Reduce stand-in (the reference's TransferTiles is next-tier work)
    def reduce(self, canon_tiles, canon_flags, tile_count):
        flat = canon_tiles.reshape(-1, 64)
        fl = canon_flags.reshape(-1)
        n = flat.shape[0]
        step = max(1, n // tile_count)
        sel = np.arange(0, step * tile_count, step)[:tile_count]
        if api._is_dev(flat):
            sel_t = torch.as_tensor(sel, device=flat.device)
            self.tiles, self.tile_flags = flat[sel_t].contiguous(), fl[sel_t].contiguous()
        else:
            self.tiles, self.tile_flags = np.ascontiguousarray(flat[sel]), np.ascontiguousarray(fl[sel])
        return self.tiles

    # --- PreparePalettes (tilingencoder.pas:1843-1871)
    def prepare_palettes(self):
        # DoPalettization (:4105-4245): LAB "special weighted DCT" features -> palette label per tile.  The CPU-era
        # BICO coreset + ANN + yakmo chain is replaced by Lloyd on the full tile set (k-means++ seeding).
        feats = api.features_f64(self.tiles, self.dithering_mode, use_lab=True)
        if self.palette_count > 1:
            labels, _, _, _ = api.kmeans_fit(feats, self.palette_count, init=None, seed=self.seed, max_iter=300)
        else:
            labels = np.zeros(feats.shape[0], dtype=np.int32)
        lab = labels.cpu().numpy() if api._is_dev(labels) else np.asarray(labels)
        # palettes re-indexed by descending use count (:4229-4234)
        counts = np.bincount(lab, minlength=self.palette_count)
        order = np.argsort(-counts, kind="stable")
        lut = np.empty(self.palette_count, dtype=np.int32)
        lut[order] = np.arange(self.palette_count, dtype=np.int32)
        self.tile_pal = self._to(lut[lab].astype(np.int32))
        # DoQuantization per palette (:4534-4564), all palettes in one call
        self.palettes, _ = api.palquant_kmeans(self.tiles, self.tile_pal, self.palette_count, self.palette_size, seed=self.seed)
        return self.palettes

    # --- Dither (tilingencoder.pas:1873-1907)
    def dither(self):
        self.tile_idx = api.dither(self.tiles, self.tile_flags, self.tile_pal, self.palettes, use_thomas_knoll=self.use_tk,
                                   y2_mixed_colors=self.y2_mixed)
        return self.tile_idx

    # --- PrepareReconstruct (tilingencoder.pas:4566-4613)
    def prepare_reconstruct(self):
        self.matcher = api.Matcher(self.tile_idx, self.tile_pal, self.palettes, extended=self.extended)
        return self.matcher

    # --- Reconstruct, k-NN branch (tilingencoder.pas:1534-1610); motion prediction is the host's (next tier)
    def reconstruct(self, canon_frame_tiles):
        """canonicalised source tiles [..., 64] -> TTileMapItem fields (TileIdx, PalIdx, err)."""
        return self.matcher.match_rgb(canon_frame_tiles.reshape(-1, 64))

    def finish_reconstruct(self):
        if self.matcher is not None:
            self.matcher.close()
            self.matcher = None

    # --- decode side, for PSNR checks (what gtm.player.js draws for a non-predicted tile, :400-420, 476-499)
    def render_tiles(self, tile_idx, pal_idx, flags):
        """numpy: chosen dictionary tiles recoloured with the chosen palette and un-mirrored -> RGB [n,64]."""
        ti = np.asarray(tile_idx.cpu() if api._is_dev(tile_idx) else tile_idx)
        pi = np.asarray(pal_idx.cpu() if api._is_dev(pal_idx) else pal_idx)
        fl = np.asarray(flags.cpu() if api._is_dev(flags) else flags).reshape(-1)
        didx = np.asarray(self.tile_idx.cpu() if api._is_dev(self.tile_idx) else self.tile_idx)
        pal = np.asarray(self.palettes.cpu() if api._is_dev(self.palettes) else self.palettes)
        px = pal[pi[:, None], didx[ti]].reshape(-1, 8, 8)
        h = (fl & 1).astype(bool)
        v = (fl & 2).astype(bool)
        px[h] = px[h][:, :, ::-1]
        px[v] = px[v][:, ::-1, :]
        return px.reshape(-1, 64)


def psnr_rgb(a, b):
    """RGB PSNR between two packed 0x00BBGGRR arrays."""
    a = np.asarray(a).astype(np.int64)
    b = np.asarray(b).astype(np.int64)
    se = 0.0
    for sh in (0, 8, 16):
        d = ((a >> sh) & 255) - ((b >> sh) & 255)
        se += float((d * d).sum())
    mse = se / (3.0 * a.size)
    return 10.0 * np.log10(255.0 * 255.0 / max(mse, 1e-12))
