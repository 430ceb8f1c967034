"""Run-to-run determinism probe of the config-C k-means leg (bench.py:kmeans_c_leg data, scaled by argv):
same data, same initial centroids -> the sharded Lloyd loop twice; prints checksums of the inputs and of every result."""
import sys, os, json, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tiler_b200 import dist as tdist, synth, api

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 16
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda", 0)
centres = torch.from_numpy(synth.random_features(K, 11, adversarial=True)).to(dev)
g = torch.Generator(device=dev); g.manual_seed(777)
pick0 = torch.randint(0, K, (K,), generator=g, device=dev)
init = (centres[pick0].float() + 25.0 * torch.randn((K, 192), generator=g, device=dev)).round().clamp(-32768, 32767).double()
x = torch.empty((N, 192), dtype=torch.int16, device=dev)
blk = 1 << 19
for a in range(0, N, blk):
    b = min(N, a + blk)
    g.manual_seed(5000 + a // blk)
    pick = torch.randint(0, K, (blk,), generator=g, device=dev)
    pts = (centres[pick].float() + 25.0 * torch.randn((blk, 192), generator=g, device=dev)).round().clamp(-32768, 32767).to(torch.int16)
    x[a:b] = pts[:b - a]
h = lambda t: hashlib.sha1(t.contiguous().cpu().numpy().tobytes()).hexdigest()[:12]
print(json.dumps({"n": N, "k": K, "x": h(x), "init": h(init)}))
for rep in range(3):
    tm = {}
    labels, cent, inertia, it = tdist.kmeans_fit_i16_sharded(x, init, max_iter=iters, check_every=1 << 30, timings=tm)
    print(json.dumps({"rep": rep, "it": it, "inertia": inertia, "labels": h(labels), "cent": h(cent), "cent_sum": float(cent.sum().item())}))
# one assignment from the SAME centroids twice: labels / sums must be bit-identical
lab = torch.full((N,), -1, dtype=torch.int32, device=dev)
shard = api.KmeansI16Shard(x, K)
out = []
for rep in range(3):
    sums = torch.empty((K, 192), dtype=torch.float64, device=dev)
    meta = torch.zeros(K + 2, dtype=torch.int64, device=dev)
    inertia = torch.zeros(1, dtype=torch.float64, device=dev)
    l2 = lab.clone()
    shard.step(init, l2, sums=sums, counts=meta[:K], stats=meta[K:], inertia=inertia)
    torch.cuda.synchronize()
    print(json.dumps({"step_rep": rep, "labels": h(l2), "sums": h(sums), "counts": h(meta[:K]), "stats": meta[K:].tolist(), "inertia": float(inertia.item())}))
shard.close()
# the same assignment with the points split into two shards (what 2 ranks do): labels and summed partials must equal the one-shard step
half = N // 2
tot_s = torch.zeros((K, 192), dtype=torch.float64, device=dev)
tot_m = torch.zeros(K + 2, dtype=torch.int64, device=dev)
tot_i = 0.0
labs = []
for a, b in ((0, half), (half, N)):
    sh = api.KmeansI16Shard(x[a:b], K)
    sums = torch.empty((K, 192), dtype=torch.float64, device=dev)
    meta = torch.zeros(K + 2, dtype=torch.int64, device=dev)
    inertia = torch.zeros(1, dtype=torch.float64, device=dev)
    l2 = torch.full((b - a,), -1, dtype=torch.int32, device=dev)
    sh.step(init, l2, sums=sums, counts=meta[:K], stats=meta[K:], inertia=inertia)
    torch.cuda.synchronize()
    tot_s += sums; tot_m += meta; tot_i += float(inertia.item()); labs.append(l2)
    sh.close()
print(json.dumps({"split": 2, "labels": h(torch.cat(labs)), "sums": h(tot_s), "counts": h(tot_m[:K]), "stats": tot_m[K:].tolist(), "inertia": tot_i}))
