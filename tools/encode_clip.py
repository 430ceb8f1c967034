"""End-to-end encode of a synthetic clip through tiler_b200 (Load -> PredictMotion -> Reduce -> PreparePalettes -> Dither ->
Reconstruct -> Reindex -> Save), decode of the GTM stream and decoded-frame PSNR.  Default = BASELINE configs[1]:
1280x720, 240 frames, 8 keyframe sequences of 30 frames, 65536-tile dictionary, 16 palettes x 16 colours."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tiler_b200 import api, gtm, synth
from tiler_b200.encoder import TilingEncoder, psnr_rgb

ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=1280); ap.add_argument("--height", type=int, default=720)
ap.add_argument("--frames", type=int, default=240); ap.add_argument("--seq", type=int, default=30)
ap.add_argument("--tiles", type=int, default=65536); ap.add_argument("--palettes", type=int, default=16)
ap.add_argument("--palette-size", type=int, default=16); ap.add_argument("--out", default="")
ap.add_argument("--decode", type=int, default=1)
ap.add_argument("--feature-mode", default="fast", choices=["fast", "exact"])
ap.add_argument("--sharded", type=int, default=0, help="run under torchrun: PredictMotion by frame, Reconstruct by sequence")
a = ap.parse_args()
t0 = time.perf_counter()
frames = np.concatenate([synth.pack_rgb(synth.make_clip(a.width, a.height, min(a.seq, a.frames - s), cut_every=0, seed=synth.SEED + s))
                         for s in range(0, a.frames, a.seq)])
seqs = [(s, min(s + a.seq, a.frames) - 1) for s in range(0, a.frames, a.seq)]
t_gen = time.perf_counter() - t0
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if a.sharded and world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
enc = TilingEncoder(palette_size=a.palette_size, palette_count=a.palettes, device=torch.device("cuda", local), feature_mode=a.feature_mode)
TilingEncoder(palette_size=a.palette_size, palette_count=a.palettes, device=torch.device("cuda", local), feature_mode=a.feature_mode).encode(
    np.ascontiguousarray(frames[:2]), [(0, 1)], tile_count=1024, sharded=bool(a.sharded))   # untimed warm-up: kernels, torch ops, memory pool
torch.cuda.synchronize()
l0 = api.kernel_launches()
api.profile_enable(not os.environ.get("TM_CUDA_PROFILER_RANGE"))
if os.environ.get("TM_CUDA_PROFILER_RANGE"):
    torch.cuda.profiler.start()
t0 = time.perf_counter()
if a.sharded and world > 1:
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
res = enc.encode(frames, seqs, tile_count=a.tiles, out_path=a.out or None, sharded=bool(a.sharded))
if a.sharded and world > 1:
    dist.barrier()
torch.cuda.synchronize()
t_enc = time.perf_counter() - t0
if os.environ.get("TM_CUDA_PROFILER_RANGE"):
    torch.cuda.profiler.stop()
if rank != 0:
    if a.sharded and world > 1:
        dist.destroy_process_group()
    sys.exit(0)
import hashlib
out = {"gtm_sha256": hashlib.sha256(res["gtm"]).hexdigest()[:16], "world": world, "clip": f"{a.width}x{a.height}x{a.frames}", "sequences": len(seqs), "generate_s": round(t_gen, 2), "encode_s": round(t_enc, 3),
       "encode_fps": round(a.frames / t_enc, 2), "timings_s": {k: round(v, 3) for k, v in res["timings"].items()},
       "dictionary_tiles_after_reduce": res["dictionary_before_reindex"], "dictionary_tiles_final": int(len(res["tiles"])),
       "gtm_bytes": len(res["gtm"]), "mean_tile_psnr": round(res["mean_tile_psnr"], 3),
       "predicted_fraction": round(float(res["tilemap"]["is_pred"].mean()), 4), "kernel_launches": api.kernel_launches() - l0}
prof = {}
for name in ('features_sliding', 'motion_search', 'knn_topk', 'rerank', 'features_rgb', 'tile_classes'):
    ms, cnt = api.profile_read(name)
    if cnt:
        prof[name] = {'total_ms': round(ms, 2), 'launches': cnt, 'ms_per_launch': round(ms / cnt, 3)}
out['kernel_ms'] = prof
if a.decode:
    t0 = time.perf_counter()
    dec, hdr = gtm.decode_gtm(res["gtm"])
    out["decode_s"] = round(time.perf_counter() - t0, 2)
    fsel = np.concatenate([np.arange(seqs[si][0], seqs[si][1] + 1) for si in res["recon_sequences"]])
    out["decoded_equals_reconstruction"] = bool(np.array_equal(dec[fsel], res["recon"].cpu().numpy()))
    out["psnr_rgb_db"] = round(psnr_rgb(dec, frames), 4)
print(json.dumps(out))
if a.sharded and world > 1:
    dist.destroy_process_group()
