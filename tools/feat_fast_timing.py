"""Sliding-window feature timing, bit-exact vs fast mode, one 720p and one 1080p frame (device-resident, CUDA events)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tiler_b200 import api, synth
res = {}
for (w, h) in ((1280, 720), (1920, 1080)):
    f0 = torch.from_numpy(synth.pack_rgb(synth.make_clip(w, h, 1, seed=5))[0]).cuda()
    for mode, name in ((api.FEATURES_EXACT, "exact"), (api.FEATURES_FAST, "fast")):
        api.set_feature_mode(mode)
        for _ in range(3):
            d = api.sliding_features(f0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            d = api.sliding_features(f0)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        nwin = (w - 7) * (h - 7)
        res[f"{w}x{h}_{name}"] = {"ms": ms, "windows_per_s": nwin / ms * 1e3, "out_GBps": nwin * 384 / ms / 1e6}
        if name == "exact":
            ref = d.clone()
        else:
            diff = (d.int() - ref.int()).abs()
            res[f"{w}x{h}_{name}"]["max_lsb"] = int(diff.max()); res[f"{w}x{h}_{name}"]["frac_differing"] = float((diff != 0).float().mean())
    api.set_feature_mode(api.FEATURES_EXACT)
print(json.dumps(res, indent=1))
