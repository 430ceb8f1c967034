"""Timing of the feature kernels (device-resident): 432000 RGB tiles and one 720p sliding pass."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tiler_b200 import api, synth
frames = synth.pack_rgb(synth.make_clip(1280, 720, 1, seed=5))
f0 = torch.from_numpy(frames[0]).cuda()
tiles = torch.randint(0, 1 << 24, (432000, 64), dtype=torch.int32, device="cuda")
for _ in range(2):
    api.sliding_features(f0); api.features_from_rgb(tiles)
torch.cuda.synchronize()
api.profile_enable(True)
for _ in range(5):
    api.sliding_features(f0); api.features_from_rgb(tiles)
torch.cuda.synchronize()
a, na = api.profile_read("features_sliding"); b, nb = api.profile_read("features_rgb")
print(json.dumps({"sliding_ms": round(a / na, 3), "rgb_432000_ms": round(b / nb, 3)}))
