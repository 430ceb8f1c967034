"""Fast feature mode, one 720p frame, device-resident: the two-pass path (sliding int16 features -> search with its own limb split)
against the fused path (tm_predict_motion_frame: the sliding kernel writes the search's operands), per-scope device times."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tiler_b200 import api, synth
frames = synth.pack_rgb(synth.make_clip(1280, 720, 2, seed=5))
f0 = torch.from_numpy(frames[0]).cuda()
tiles = torch.from_numpy(synth.frame_to_tiles(frames[1])).cuda()
canon, flags = api.mirror_canonicalise(tiles)
cur = api.features_from_rgb_mirrored(canon, flags)
api.set_feature_mode(api.FEATURES_FAST)
res = {}
for name in ("two_pass", "fused"):
    def run():
        if name == "two_pass":
            d = api.sliding_features(f0)
            return api.motion_search(cur, 160, 90, d, 32)
        return api.predict_motion_frame(f0, canon, flags, 160, 90, 32)
    for _ in range(3):
        out = run()
    torch.cuda.synchronize()
    api.profile_enable(True)
    api.profile_read("features_sliding"); api.profile_read("motion_search")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = run()
    e1.record(); torch.cuda.synchronize()
    fs, n1 = api.profile_read("features_sliding"); ms, n2 = api.profile_read("motion_search")
    api.profile_enable(False)
    res[name] = {"features_sliding_ms": fs / n1, "motion_search_ms": ms / n2, "total_ms": e0.elapsed_time(e1) / 10,
                 "err_sum": int(out[2].to(torch.int64).sum().item())}
print(json.dumps(res))
