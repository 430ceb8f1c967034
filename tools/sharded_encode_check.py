"""Run under torchrun (one process per GPU): the sharded encode of a small synthetic clip (sharded upload + all-gather of the frames,
PredictMotion by frame, Reconstruct by keyframe sequence, device-tensor exchange of PSNRs / tilemaps / use counts, per-rank LZMA
chunks) must produce the bytes of the single-process encode.  Rank 0 prints one JSON line."""
import hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from tiler_b200 import synth
from tiler_b200.encoder import TilingEncoder

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
w, h, n = 320, 184, 22                                   # 22 frames: not a multiple of the world size; 5 uneven sequences
frames = synth.pack_rgb(synth.make_clip(w, h, n, cut_every=0, seed=99, n_sprites=6))
seqs = [(0, 3), (4, 9), (10, 10), (11, 17), (18, 21)]
out = {}
for mode in ("fast", "exact"):
    sh = TilingEncoder(palette_size=16, palette_count=8, device=dev, seed=5, feature_mode=mode).encode(frames, seqs, tile_count=900, sharded=True)
    dist.barrier()
    if rank == 0:
        one = TilingEncoder(palette_size=16, palette_count=8, device=dev, seed=5, feature_mode=mode).encode(frames, seqs, tile_count=900)
        out[mode] = {"equal": sh["gtm"] == one["gtm"], "bytes": len(one["gtm"]), "sha": hashlib.sha256(one["gtm"]).hexdigest()[:12],
                     "tilemap_equal": all(np.array_equal(sh["tilemap"][k], one["tilemap"][k]) for k in ("tile_idx", "pal_idx", "is_pred", "err"))}
    dist.barrier()
if rank == 0:
    print(json.dumps({"world": world, **out}))
dist.destroy_process_group()
