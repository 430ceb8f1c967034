"""BASELINE configs[4] -- dither / palette stress: one 3840x2160 frame (129 600 tiles), 32 palettes x 256 colours quantised
from 32 horizontal bands, every tile dithered against every palette (4 147 200 pairs) in Thomas-Knoll and Yliluoma mode.
Device-resident, CUDA events."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tiler_b200 import api, synth

W, H, NP, PS = 3840, 2160, 32, 256
frame = synth.pack_rgb(synth.make_clip(W, H, 1, seed=synth.SEED, n_sprites=40))[0]
tiles_h = synth.frame_to_tiles(frame)
n = tiles_h.shape[0]
tiles = torch.from_numpy(tiles_h).cuda()
canon, flags = api.mirror_canonicalise(tiles)
band = torch.from_numpy(((np.arange(n) // (W // 8)) * NP // (H // 8)).astype(np.int32)).cuda()   # 32 spatial bands
def timed(fn, reps=1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return r, e0.elapsed_time(e1) / reps
(pal, iters), t_pal = timed(lambda: api.palquant_kmeans(canon, band, NP, PS, seed=1))
pair_tile = torch.arange(n, dtype=torch.int32, device="cuda").repeat_interleave(NP)
pair_pal = torch.arange(NP, dtype=torch.int32, device="cuda").repeat(n)
out = {"tiles": n, "pairs": int(pair_tile.numel()), "palquant_ms": round(t_pal, 2), "palquant_iters": iters}
for name, tk in (("thomas_knoll", True), ("yliluoma_mix4", False)):
    api.dither(canon[:1024], flags[:1024], pair_pal[:1024] * 0, pal, use_thomas_knoll=tk)   # warm-up
    idx, ms = timed(lambda: api.dither(canon, flags, pair_pal, pal, use_thomas_knoll=tk, y2_mixed_colors=4, pair_tile=pair_tile))
    pairs = pair_tile.numel()
    out[name] = {"ms": round(ms, 2), "pairs_per_s": round(pairs / ms * 1e3), "algorithmic_GBps": round(pairs * 320 / ms / 1e6, 2),
                 "colour_compares_per_s": (round(pairs * 64 * 64 * PS / ms * 1e3) if tk else None)}
print(json.dumps(out))
