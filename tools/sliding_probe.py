"""One sliding-features + motion-search pass on a 720p frame (device-resident): shows those kernels to ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tiler_b200 import api, synth
frames = synth.pack_rgb(synth.make_clip(1280, 720, 2, seed=5))
f0 = torch.from_numpy(frames[0]).cuda()
tiles = torch.from_numpy(synth.frame_to_tiles(frames[1])).cuda()
cur = api.features_from_rgb(tiles)
for _ in range(2):
    d = api.sliding_features(f0)
    api.motion_search(cur, 160, 90, d, 32)
torch.cuda.synchronize()
print("ok")
