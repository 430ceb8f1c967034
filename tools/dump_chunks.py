"""Dumps the raw (uncompressed) GTM chunks of the bench clip's encode to gpurun_out/chunks/ (input for offline LZMA encoder
profiling) and prints per-chunk compression times measured on the GPU box's host cores."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tiler_b200 import gtm, synth
from tiler_b200.encoder import TilingEncoder
frames = np.concatenate([synth.pack_rgb(synth.make_clip(1280, 720, 30, cut_every=0, seed=synth.SEED + s)) for s in range(8)])
seqs = [(s, s + 29) for s in range(0, 240, 30)]
os.makedirs("gpurun_out/chunks", exist_ok=True)
orig = gtm.lzma_encode
times = []
def spy(raw, *a, **k):
    t0 = time.perf_counter(); c = orig(raw, *a, **k); dt = time.perf_counter() - t0
    times.append((len(raw), len(c), dt))
    open(f"gpurun_out/chunks/chunk_{len(raw)}.bin", "wb").write(raw)
    return c
gtm.lzma_encode = spy
enc = TilingEncoder(palette_size=16, palette_count=16, device=torch.device("cuda", 0), feature_mode="fast")
res = enc.encode(frames, seqs, tile_count=65536)
print(json.dumps({"chunks": [{"raw": r, "comp": c, "ms": round(t * 1e3, 1)} for r, c, t in times], "timings": res["timings"], "cpus": os.cpu_count()}))
