"""Stage drivers mirroring TTilingEncoder's call structure for the data-parallel core
(tilingencoder.pas:1843-1962): PreparePalettes -> Dither -> PrepareReconstruct -> Reconstruct, each one a handful of
batched calls into libtm_gpu.so instead of the per-tile DLL calls of the reference.

`encode()` chains every step of TTilingEncoder.Run (Load -> PredictMotion -> Reduce -> PreparePalettes -> Dither ->
Reconstruct -> Reindex -> Save, tilingencoder.pas:5530-5552) on a clip already in memory; the heavy steps are library
calls, the bookkeeping between them (threshold search over per-class PSNRs, dictionary ordering, stream writing) is
host code as it is in the reference.  Not here: video decoding (FFmpeg) and keyframe detection (sequences are an input).  `reduce_sample()` is the old stand-in that samples dictionary
tiles without the motion pass; bench.py's match-stage step still uses it.
"""
import numpy as np

from . import api
from . import gtm as gtm_io

C_PSNR_MAX = np.float64(10.0 * np.log(255.0 * 255.0 / 0.5) / np.log(10.0))   # cPsnrMaxValue, utils.pas:111
C_INV_PHI = 2.0 / (1.0 + np.sqrt(5.0))                                         # cInvPhi, utils.pas:42-43


def _euclidean_to_psnr_1(err):
    r = (np.asarray(err).astype(np.float64) * (1.0 / 192.0)).astype(np.float32)
    m = np.maximum(r.astype(np.float64), 0.5)
    return (10.0 * np.log10(255.0 * 255.0 / m)).astype(np.float32)


def euclidean_to_psnr(err):
    """EuclideanToPSNR (utils.pas:1074-1078), vectorised: Single(d / 192) -> max 0.5 -> 10 log10(255^2 / x) -> Single.
    A whole clip's 3.4 M values are converted in slices on a few threads (numpy releases the GIL inside its loops): the same
    element-wise operations, so the same bits, in a quarter of the time."""
    err = np.asarray(err)
    if err.size < (1 << 18):
        return _euclidean_to_psnr_1(err)
    flat = np.ascontiguousarray(err).reshape(-1)
    out = np.empty(flat.shape, dtype=np.float32)
    n_thr = 8
    bounds = np.linspace(0, flat.size, n_thr + 1).astype(np.int64)

    def work(i):
        out[bounds[i]:bounds[i + 1]] = _euclidean_to_psnr_1(flat[bounds[i]:bounds[i + 1]])
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(n_thr) as pool:
        list(pool.map(work, range(n_thr)))
    return out.reshape(err.shape)


def _reindex_order(tiles_rgb, use_count):
    """Order of ReindexTiles(True) (tilingencoder.pas:4626-4696, CompareTileUseCountRev :582-599): use count descending, then
    CompareDWord on the 64 pixels (unsigned dwords, first difference decides).  The pixel order is a sort on the first two
    dwords packed into one uint64 key; only the rows that tie on it (flat tiles) go through the 256-byte memcmp sort.  A
    stable sort on the use count follows."""
    t = np.ascontiguousarray(np.asarray(tiles_rgb)).view(np.uint32).reshape(len(tiles_rgb), 64)
    key = (t[:, 0].astype(np.uint64) << np.uint64(32)) | t[:, 1].astype(np.uint64)
    o1 = np.argsort(key, kind="stable")
    ks = key[o1]
    tie = np.zeros(len(o1), dtype=bool)
    eq = ks[1:] == ks[:-1]
    tie[1:] |= eq
    tie[:-1] |= eq
    if tie.any():
        # rows of all tie runs, re-sorted by their full big-endian bytes: the key is a prefix of those bytes, so the result is
        # still grouped by key in ascending order and drops back into the same positions
        idx = o1[tie]
        be = np.ascontiguousarray(t[idx].astype(">u4")).view(np.uint8).reshape(len(idx), 256)
        o1[tie] = idx[np.argsort(be.view(np.dtype((np.void, 256))).reshape(-1), kind="stable")]
    o2 = np.argsort(-np.asarray(use_count)[o1].astype(np.int64), kind="stable")
    return o1[o2]


def _quicksort_desc(counts):
    """Permutation QuickSort(..., ComparePaletteUseCount) leaves (extern.pas:370-418, utils.pas:750-753): descending by
    count; equal counts land where the reference's middle-pivot Hoare partition (pivot index tracked through swaps) puts
    them, so the procedure is replayed rather than replaced by a stable sort."""
    key = [int(c) for c in counts]
    idx = list(range(len(key)))
    stack = [(0, len(key) - 1)]
    while stack:
        first, last = stack.pop()
        while last > first:
            i, j, p = first, last, (first + last) >> 1
            while True:
                while key[idx[i]] > key[idx[p]]:
                    i += 1
                while key[idx[j]] < key[idx[p]]:
                    j -= 1
                if i <= j:
                    idx[i], idx[j] = idx[j], idx[i]
                    p = j if p == i else (i if p == j else p)
                    i += 1
                    j -= 1
                if i > j:
                    break
            if first < j:
                stack.append((first, j))   # disjoint sub-ranges: the order in which they are sorted does not matter
            first = i
    return np.asarray(idx, dtype=np.int64)


def golden_ratio_search(func, min_x, max_x, objective_y, eps_x=1e-6, eps_y=0.5):
    """GoldenRatioSearch (utils.pas:1044-1072).  Returns (result x, last evaluated x): the encoder state after the search
    is the one left by the LAST evaluation, not by the returned abscissa."""
    last = None
    while True:
        if abs(min_x - max_x) <= eps_x:
            return min_x, last
        t = (1.0 - C_INV_PHI) if min_x < max_x else C_INV_PHI
        x = min_x + (max_x - min_x) * t
        y = func(x)
        last = x
        if abs(y - objective_y) <= eps_y:
            return x, last
        if y < objective_y:
            min_x = x
        else:
            max_x = x

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


class TilingEncoder:
    """Names follow the reference's settings (LoadDefaultSettings, tilingencoder.pas:3817-3845)."""

    def __init__(self, palette_size=16, palette_count=16, dithering_mode=api.PVS_WEIGHTED_SPE_DCT,
                 dithering_use_thomas_knoll=True, dithering_yliluoma2_mixed_colors=4,
                 frame_tiling_extended_palette_usage=True, seed=0x42381337, device=None, feature_mode="exact",
                 optimize_palettes=True):
        self.palette_size = int(np.clip(palette_size, 2, 256))      # reference clamps to 2..64 (:2965); 256 = stress shape
        self.palette_count = int(np.clip(palette_count, 1, 65536))
        self.dithering_mode = dithering_mode
        self.use_tk = bool(dithering_use_thomas_knoll)
        self.y2_mixed = int(np.clip(dithering_yliluoma2_mixed_colors, 1, 16))
        self.extended = bool(frame_tiling_extended_palette_usage)
        self.seed = seed
        self.device = device  # None: host arrays through the ABI; a torch device: everything stays in HBM
        # "exact": every feature in DCTInner_asm's summation order (bit-exact); "fast": the sliding-window features of the motion
        # searches (DoDCTs) through the separable f64 kernel (<= 1 LSB on < 1e-3 of the coefficients, PSNR within 0.05 dB)
        self.feature_mode = api.FEATURES_FAST if feature_mode == "fast" else api.FEATURES_EXACT
        self.optimize = bool(optimize_palettes)   # PreparePalettes ends with OptimizePalettes (:1868); False leaves the (V,S,H) order
        self.tiles = None       # dictionary tiles, RGB [n,64] (canonical orientation)
        self.tile_flags = None  # their initial mirrors
        self.tile_pal = None    # PalIdx_Initial
        self.palettes = None    # [palette_count, palette_size] int32
        self.tile_idx = None    # dithered palette indices [n,64]
        self.matcher = None
        self.use_count = None
        self.reduce_threshold = None

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

    # --- PredictMotion (tilingencoder.pas:1964-1991, 1154-1282): every tile against the previous SOURCE frame
    def predict_motion(self, frames_packed, canon_tiles, canon_flags, tw, th, radius=32, frame_range=None):
        """frames_packed [n, th*8, tw*8] source pixels; -> (psnr f32 [n, nt], pred_x, pred_y) on the host.
        Frame 0 is predicted from frame 1 (:1982-1984); a one-frame clip is predicted from a black buffer.
        frame_range = (lo, hi) restricts the work to those frames (multi-GPU sharding by frame): rows [lo, hi) are returned."""
        n = int(frames_packed.shape[0])
        nt = tw * th
        lo, hi = frame_range if frame_range is not None else (0, n)
        psnr_all = np.empty((n, nt), np.float32)
        px_all = np.empty((n, nt), np.int32)
        py_all = np.empty((n, nt), np.int32)
        psnr, px, py = psnr_all, px_all, py_all
        fr = self._to(frames_packed)
        pending = []          # device-resident encoder: results stay on the device until the loop is over (one copy, no per-frame sync)
        for f in range(lo, hi):
            if f > 0:
                prev = fr[f - 1]
            elif n > 1:
                prev = fr[1]
            else:
                prev = fr[0] * 0
            x, y, e = api.predict_motion_frame(prev, canon_tiles[f], canon_flags[f], tw, th, radius)
            if api._is_dev(e):
                pending.append((x, y, e))
                continue
            psnr[f] = euclidean_to_psnr(np.asarray(e).view(np.uint32))
            px[f] = x
            py[f] = y
        if pending:
            xs = torch.stack([p[0] for p in pending]).cpu().numpy()
            ys = torch.stack([p[1] for p in pending]).cpu().numpy()
            es = torch.stack([p[2] for p in pending]).cpu().numpy()
            px[lo:hi], py[lo:hi] = xs, ys
            psnr[lo:hi] = euclidean_to_psnr(es.view(np.uint32))
        return psnr_all[lo:hi], px_all[lo:hi], py_all[lo:hi]

    # --- Reduce (tilingencoder.pas:1908-1926, 4014-4103, 4626-4696, 4720-4781)
    def reduce(self, canon_tiles, canon_flags, psnr, seq_start_frames, tile_count):
        """SolveTileCount: golden-ratio search of the PSNR threshold x for which the number of distinct unpredicted tiles
        equals tile_count.  A tile is predicted when its motion PSNR > x (PSNR / 10 > x on the first frame of a keyframe
        sequence, :4028-4031).  The distinct-tile count of MakeTilesUnique(True) comes from exact duplicate classes computed
        once on the GPU; the count for a threshold is then the number of classes whose smallest effective PSNR is <= x.
        -> tilemap TileIdx [n, nt] (-1 = predicted), and fills self.tiles / tile_flags / use_count in ReindexTiles(True)
        order (use count descending, then RGB pixels as unsigned dwords ascending)."""
        shape = tuple(canon_tiles.shape[:2])
        flat = canon_tiles.reshape(-1, 64)
        n_all = int(flat.shape[0])
        cls, n_cls = api.tile_classes(flat)
        # STCGREval (:4028-4031) promotes the Single PSNR to Double, divides by 10.0 in double and compares with the Double x
        eff = np.asarray(psnr, dtype=np.float32).astype(np.float64)
        for f in seq_start_frames:
            eff[f] = eff[f] / 10.0
        eff = eff.reshape(-1)
        target = min(int(tile_count), n_all)
        fl = canon_flags.reshape(-1)
        # per-class bookkeeping over millions of tiles runs in the library (csrc/reduce.cu); only per-class arrays and the
        # <= tile_count chosen representatives come back to the host for the final ReindexTiles ordering
        to_host = lambda a: a.cpu().numpy() if api._is_dev(a) else np.asarray(a)
        eff_d = self._to(eff)
        sorted_min = to_host(api.reduce_class_min(cls, eff_d, n_cls))
        x_res, x_last = golden_ratio_search(lambda x: float(np.searchsorted(sorted_min, x, side="right")), 0.0,
                                            float(C_PSNR_MAX), float(target))
        x = float(x_last if x_last is not None else x_res)
        use_d, rep_d, unpred_d = api.reduce_apply(cls, eff_d, n_cls, x)       # IsPredicted := PSNR > x
        use, rep = to_host(use_d), to_host(rep_d)
        chosen = np.nonzero(use > 0)[0]
        rep_idx = rep[chosen].astype(np.int64)
        if api._is_dev(flat):
            rep_t = torch.from_numpy(rep_idx).to(flat.device)
            rep_tiles_host = flat[rep_t].cpu().numpy()
        else:
            rep_tiles_host = flat[rep_idx]
        order = _reindex_order(rep_tiles_host, use[chosen])                  # ReindexTiles(True): (use count desc, CompareDWord asc)
        chosen, rep_idx = chosen[order], rep_idx[order]
        new_of_cls = np.full(n_cls, -1, dtype=np.int32)
        new_of_cls[chosen] = np.arange(len(chosen), dtype=np.int32)
        tile_idx = to_host(api.reduce_remap(cls, unpred_d, self._to(new_of_cls))).reshape(shape)
        if api._is_dev(flat):
            rep_t = torch.from_numpy(rep_idx).to(flat.device)
            self.tiles, self.tile_flags = flat[rep_t].contiguous(), fl[rep_t].contiguous()
        else:
            self.tiles, self.tile_flags = np.ascontiguousarray(flat[rep_idx]), np.ascontiguousarray(fl[rep_idx])
        self.use_count = use[chosen].astype(np.int32)
        self.reduce_threshold = x
        return tile_idx

    # --- whole pipeline (TTilingEncoder.Run, tilingencoder.pas:5530-5552)
    def encode(self, frames_packed, sequences, tile_count, radius=32, fps=24.0, out_path=None, emit_skip_blocks=True, sharded=False):
        """frames_packed [n, H, W] int32 0x00BBGGRR with H, W multiples of 8; sequences = [(start, end)] inclusive.
        -> dict(gtm bytes, tilemap, recon frames, dictionary, timings).
        sharded=True (inside an initialised torch.distributed group, one process per GPU, every rank holding the clip):
        PredictMotion is sharded by frame and Reconstruct by keyframe sequence (no data-path collective, the
        per-tile PSNRs and the tilemaps are gathered on the host); Reduce, palettes and dithering are deterministic and run
        replicated; rank 0 writes the stream.  The bytes equal the single-GPU encode's."""
        prev_mode = api.set_feature_mode(self.feature_mode)
        try:
            return self._encode(frames_packed, sequences, tile_count, radius, fps, out_path, emit_skip_blocks, sharded)
        finally:
            api.set_feature_mode(prev_mode)

    def _encode(self, frames_packed, sequences, tile_count, radius, fps, out_path, emit_skip_blocks, sharded):
        import time
        from . import dist as tdist
        rank, world = tdist.world_info() if sharded else (0, 1)
        n, H, W = (int(v) for v in frames_packed.shape)
        # the tilemap rounds up to whole tiles and the screen IS the tilemap (:1776, ReframeUI :2631-2638); pixels beyond the
        # image stay 0 (AllocMem'd frame tiles, :1310): configs[0] is 320x180 -> 40x23 tiles, a 320x184 screen
        tw, th = (W - 1) // 8 + 1, (H - 1) // 8 + 1
        nt = tw * th
        t = {}
        t0 = time.perf_counter()
        if (H, W) != (th * 8, tw * 8):
            if api._is_dev(frames_packed):
                padded = torch.zeros((n, th * 8, tw * 8), dtype=frames_packed.dtype, device=frames_packed.device)
            else:
                padded = np.zeros((n, th * 8, tw * 8), dtype=np.int32)
            padded[:, :H, :W] = frames_packed
            frames_packed, H, W = padded, th * 8, tw * 8
        if self.device is not None:   # frames -> tiles is a pure layout change: done on the device (torch = memory plumbing)
            if api._is_dev(frames_packed):
                fr_dev = frames_packed
            elif world > 1 and tdist.device_collectives() and n >= world:
                # every rank holds the clip on the host, but N ranks pushing all of it over the host links at once is the slow way
                # round: each uploads n / N frames and the rest arrives from the peers over NVLink (one all-gather)
                import torch.distributed as tdd
                per = (n + world - 1) // world
                lo_u, hi_u = min(n, rank * per), min(n, (rank + 1) * per)
                buf = torch.zeros((world * per, H, W), dtype=torch.int32, device=self.device)
                if hi_u > lo_u:
                    buf[lo_u:hi_u] = torch.from_numpy(np.ascontiguousarray(frames_packed[lo_u:hi_u])).to(self.device)
                tdd.all_gather_into_tensor(buf, buf[rank * per:(rank + 1) * per].clone())
                fr_dev = buf[:n]
            else:
                fr_dev = torch.from_numpy(np.ascontiguousarray(frames_packed)).to(self.device)
            tiles = fr_dev.view(n, th, 8, tw, 8).permute(0, 1, 3, 2, 4).contiguous().view(n, nt, 64)
            frames_packed = fr_dev
        else:
            tiles = np.ascontiguousarray(np.asarray(frames_packed).reshape(n, th, 8, tw, 8).transpose(0, 1, 3, 2, 4).reshape(n, nt, 64))
        canon, flags = self.load_tiles(tiles)
        t["load"] = time.perf_counter() - t0; t0 = time.perf_counter()
        if world > 1:
            lo, hi = tdist.shard_rows(n, rank, world)
            psnr_loc, _, _ = self.predict_motion(frames_packed, canon, flags, tw, th, radius, frame_range=(lo, hi))
            if self.device is not None and tdist.device_collectives():
                psnr = tdist.allgather_rows_device(psnr_loc, lo, n, self.device).cpu().numpy()
            else:
                psnr = tdist.gather_rows(psnr_loc, n)
        else:
            psnr, _, _ = self.predict_motion(frames_packed, canon, flags, tw, th, radius)
        t["predict_motion"] = time.perf_counter() - t0; t0 = time.perf_counter()
        self.reduce(canon, flags, psnr, [s for s, _ in sequences], tile_count)
        t["reduce"] = time.perf_counter() - t0; t0 = time.perf_counter()
        self.prepare_palettes()
        t["prepare_palettes"] = time.perf_counter() - t0; t0 = time.perf_counter()
        self.dither()
        if self.device is not None:
            api.synchronize()          # the call only enqueues on device tensors: bill the stage its own time
        t["dither"] = time.perf_counter() - t0; t0 = time.perf_counter()
        self.prepare_reconstruct()
        t["prepare_reconstruct"] = time.perf_counter() - t0; t0 = time.perf_counter()
        keys = ("tile_idx", "pal_idx", "pred_x", "pred_y", "is_pred", "err", "psnr")
        mine = tdist.shard_sequences([s1 - s0 + 1 for s0, s1 in sequences], world)
        local, recon_of = {}, {}
        n_dict = int(self.tile_idx.shape[0])
        dev_coll = world > 1 and self.device is not None and tdist.device_collectives()
        full = None
        tile_use_d = None
        for si in mine[rank]:
            s0, s1 = sequences[si]
            r = self.matcher.reconstruct_sequence(canon[s0:s1 + 1], flags[s0:s1 + 1], tw, th, radius=radius)
            ti = r["tile_idx"]
            # references per dictionary tile (Reindex's use count, predicted items included), counted where the tilemap is
            if api._is_dev(ti):
                use_si = torch.bincount(ti.reshape(-1).to(torch.int64) + 1, minlength=n_dict + 1)[1:]
            else:
                use_si = np.bincount(np.asarray(ti).reshape(-1).astype(np.int64) + 1, minlength=n_dict + 1)[1:]
            if dev_coll:
                # multi-GPU, device-resident: this rank's rows go into zero-filled full-size device tensors; one all-reduce per
                # field below is the all-gather (NCCL over NVLink instead of pickled host objects)
                if full is None:
                    full = {k: torch.zeros((n, nt), dtype=r[k].dtype, device=self.device) for k in keys}
                    tile_use_d = torch.zeros(n_dict, dtype=torch.int64, device=self.device)
                for k in keys:
                    full[k][s0:s1 + 1] = r[k]
                tile_use_d += use_si
            else:
                # device results stay on the device until every sequence of this rank is enqueued: a .cpu() here would make the
                # host wait for this sequence before it can launch the next one
                local[si] = {k: r[k] for k in keys}
                local[si]["use"] = use_si
            recon_of[si] = r["recon"]        # stays on the device when the encoder is device-resident
        for si in local:
            local[si] = {k: (v.cpu().numpy() if api._is_dev(v) else v) for k, v in local[si].items()}
        if dev_coll:
            import torch.distributed as tdd
            if full is None:   # a rank without sequences still takes part in the collectives
                ref_dt = {"tile_idx": torch.int32, "pal_idx": torch.int32, "pred_x": torch.int32, "pred_y": torch.int32,
                          "is_pred": torch.uint8, "err": torch.int32, "psnr": torch.float32}
                full = {k: torch.zeros((n, nt), dtype=ref_dt[k], device=self.device) for k in keys}
                tile_use_d = torch.zeros(n_dict, dtype=torch.int64, device=self.device)
            for k in keys:
                tdd.all_reduce(full[k], op=tdd.ReduceOp.SUM)
            tdd.all_reduce(tile_use_d, op=tdd.ReduceOp.SUM)
            tm = {k: full[k].cpu().numpy() for k in keys}
            tile_use = tile_use_d.cpu().numpy()
        else:
            merged = tdist.gather_tilemaps(local, mine) if world > 1 else local
            tm = {k: np.concatenate([merged[si][k] for si in range(len(sequences))]) for k in keys}
            tile_use = sum(merged[si]["use"] for si in range(len(sequences)))
        own = [recon_of[si] for si in sorted(recon_of)]
        recon = (torch.cat(own) if api._is_dev(own[0]) else np.concatenate(own)) if own else None   # this rank's sequences only
        tm["err"] = tm["err"].view(np.uint32)
        tm["mirror"] = flags.cpu().numpy() if api._is_dev(flags) else np.asarray(flags)
        self.finish_reconstruct()
        t["reconstruct"] = time.perf_counter() - t0; t0 = time.perf_counter()
        didx = self.tile_idx.cpu().numpy() if api._is_dev(self.tile_idx) else np.asarray(self.tile_idx)
        pal = self.palettes.cpu().numpy() if api._is_dev(self.palettes) else np.asarray(self.palettes)
        final_tiles, use_count, tile_map = gtm_io.reindex(didx, tm["tile_idx"], tile_use=tile_use)
        tm_out = dict(tm)
        tm_out["tile_idx"] = tile_map
        t["reindex"] = time.perf_counter() - t0; t0 = time.perf_counter()
        # every rank serialises and LZMA-compresses the chunks of its own keyframe sequences; the chunks are all-gathered
        data = gtm_io.write_gtm(out_path if rank == 0 else None, tm_out, final_tiles, use_count, pal, tw, th, sequences, fps=fps,
                                settings_text=f"tiler_b200 PaletteSize={self.palette_size} PaletteCount={self.palette_count}",
                                emit_skip_blocks=emit_skip_blocks, only=(set(mine[rank]) if world > 1 else None),
                                exchange=((lambda d: tdist.gather_tilemaps(d, mine)) if world > 1 else None))
        t["save"] = time.perf_counter() - t0
        return {"gtm": data, "tilemap": tm_out, "recon": recon, "tiles": final_tiles, "use_count": use_count, "palettes": pal,
                "recon_sequences": sorted(recon_of), "timings": t, "mean_tile_psnr": float(tm["psnr"].mean()), "dictionary_before_reindex": int(didx.shape[0])}

    # --- MergeTiles hook (tilingencoder.pas:4783-4840): where a clustering-built dictionary plugs in
    def merge_tiles(self, clusters, best, tile_idx):
        """InitMergeTiles + one MergeTiles(indices, n, BestIdx, nil, nil) per cluster + FinishMergeTiles + ReindexTiles(True):
        every tile of a cluster other than `best[c]` hands its UseCount to the best tile, goes inactive and leaves a MergeIndex;
        tilemap items follow the MergeIndex; the surviving tiles are re-ordered (use count descending, RGB pixels ascending).
        clusters[i] = cluster of dictionary tile i, best[c] = index of the tile cluster c keeps.  -> remapped tilemap TileIdx."""
        clusters = np.asarray(clusters.cpu() if api._is_dev(clusters) else clusters).astype(np.int64)
        best = np.asarray(best).astype(np.int64)
        tmap = np.asarray(tile_idx).astype(np.int64)
        n = len(clusters)
        use = np.asarray(self.use_count).astype(np.int64)
        merge_index = np.full(n, -1, dtype=np.int64)                          # InitMergeTiles (:4817-4823)
        tgt = best[clusters]
        moved = tgt != np.arange(n)
        merge_index[moved] = tgt[moved]                                       # MergeTiles (:4800-4813)
        new_use = np.bincount(tgt, weights=use, minlength=n).astype(np.int64)
        new_use[moved] = 0
        flat = tmap.reshape(-1)
        mi = np.where(flat >= 0, merge_index[np.maximum(flat, 0)], -1)        # FinishMergeTiles (:4825-4840)
        flat = np.where(mi >= 0, mi, flat)
        keep = np.nonzero(new_use > 0)[0]                                     # ReindexTiles(True) (:4626-4696)
        tiles_h = self.tiles.cpu().numpy() if api._is_dev(self.tiles) else np.asarray(self.tiles)
        fl_h = self.tile_flags.cpu().numpy() if api._is_dev(self.tile_flags) else np.asarray(self.tile_flags)
        order = keep[_reindex_order(tiles_h[keep], new_use[keep])]
        new_of = np.full(n, -1, dtype=np.int64)
        new_of[order] = np.arange(len(order))
        out = np.where(flat >= 0, new_of[np.maximum(flat, 0)], -1).astype(np.int32).reshape(tmap.shape)
        self.tiles, self.tile_flags = self._to(np.ascontiguousarray(tiles_h[order])), self._to(np.ascontiguousarray(fl_h[order]))
        self.use_count = new_use[order].astype(np.int32)
        return out

    def cluster_dictionary(self, k, tile_idx, max_iter=300):
        """The north star's stage 3: k-means of the dictionary tiles' 192-d feature vectors (tensor-core assignment,
        tm_kmeans_fit_i16) into k clusters, each keeping its member nearest to the centroid (ties: lowest index), merged
        through the MergeTiles hook.  Initial centroids = the k most used tiles (the dictionary is ordered by use count).
        -> remapped tilemap TileIdx; self.tiles / use_count shrink to <= k tiles."""
        feats = api.features_from_rgb(self.tiles)
        f_h = feats.cpu().numpy() if api._is_dev(feats) else np.asarray(feats)
        n = f_h.shape[0]
        k = int(min(k, n))
        init = f_h[:k].astype(np.float64)
        labels, cent, _, _, _ = api.kmeans_fit_i16(feats, k, self._to(init), max_iter=max_iter)
        lab = labels.cpu().numpy() if api._is_dev(labels) else np.asarray(labels)
        cen = cent.cpu().numpy() if api._is_dev(cent) else np.asarray(cent)
        d = ((f_h.astype(np.float64) - cen[lab]) ** 2).sum(1)
        order = np.lexsort((np.arange(n), d, lab))                            # per cluster: smallest distance, then lowest index
        first = np.r_[True, lab[order][1:] != lab[order][:-1]]
        best = np.full(k, -1, dtype=np.int64)
        best[lab[order][first]] = order[first]
        return self.merge_tiles(lab, best, tile_idx)

    # --- Reduce stand-in used by bench.py's match-stage step: samples dictionary tiles, no motion pass
    def reduce_sample(self, canon_tiles, canon_flags, tile_count):
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
        """DoPalettization (:4105-4245) as the reference chains it, each DLL stage one batched library call:
        LAB "special weighted DCT" features -> coreset of 8 x PaletteCount points, tiles weighted by UseCount (:4149-4173;
        BICO's streaming CF-tree is replaced by the library's weighted-Lloyd summary, tm_coreset_weighted) -> nearest
        coreset point of every tile (ANN, :4183-4188) -> UNWEIGHTED k-means of the coreset points into PaletteCount
        clusters (yakmo, :4198-4207) -> identity when the coreset has <= PaletteCount points (:4214-4219), no k-means when
        PaletteCount = 1 (:4209-4212) -> palettes re-indexed by descending tile count (:4221-4244).  Then DoQuantization for
        every palette in one call (:4534-4564).  OptimizePalettes (:4265-4432) is optimize_palettes()."""
        P = self.palette_count
        feats = api.features_f64(self.tiles, self.dithering_mode, use_lab=True)
        n = int(feats.shape[0])
        use = self.use_count if self.use_count is not None else np.ones(n, dtype=np.int32)
        core, _ = api.coreset_weighted(feats, self._to(np.asarray(use, dtype=np.float64)), P << 3, self.seed)
        core_d = self._to(core)
        ann, _ = api.knn_double(core_d, feats)
        ann = ann.cpu().numpy() if api._is_dev(ann) else np.asarray(ann)
        if len(core) > P:
            if P > 1:
                yk, _, _, _ = api.kmeans_fit(core_d, P, init=None, seed=self.seed, max_iter=300)
                yk = yk.cpu().numpy() if api._is_dev(yk) else np.asarray(yk)
            else:
                yk = np.zeros(len(core), dtype=np.int32)
        else:
            yk = np.arange(len(core), dtype=np.int32)
        lab = yk[ann]
        # palettes re-indexed by descending use count with the repo's own non-stable QuickSort (:4229-4234, extern.pas:370)
        counts = np.bincount(lab, minlength=P)
        order = _quicksort_desc(counts)
        lut = np.empty(P, dtype=np.int32)
        lut[order] = np.arange(P, dtype=np.int32)
        self.tile_pal = self._to(lut[lab].astype(np.int32))
        self.coreset_size = int(len(core))
        # DoQuantization per palette (:4534-4564), all palettes in one call
        self.palettes, _ = api.palquant_kmeans(self.tiles, self.tile_pal, P, self.palette_size, seed=self.seed)
        if self.optimize:
            self.optimize_palettes()
        return self.palettes

    def optimize_palettes(self):
        """OptimizePalettes (:4246-4432): Powell search (powell.pas) over colour permutations inside each palette, repeated over
        all palettes until the mean objective stops improving.  Host code in the reference (no DLL): libtm_gtm.so, CPU threads."""
        pal_h = self.palettes.cpu().numpy() if api._is_dev(self.palettes) else np.asarray(self.palettes)
        new, self.optimize_passes = gtm_io.optimize_palettes(pal_h)
        self.palettes = self._to(new)
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
