"""Multi-GPU k-means check (run under torchrun, one rank per GPU): points sharded, centroids replicated, NCCL all-reduce of
per-cluster sums/counts each Lloyd iteration.  Rank 0 compares against the single-GPU tm_kmeans_fit on the full set and
prints one JSON line with the timings."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from tiler_b200 import api, dist as tdist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, dim, k = int(os.environ.get("KM_N", 262144)), 192, int(os.environ.get("KM_K", 1024))
rng = np.random.default_rng(11)
centres = rng.normal(0, 300, size=(k, dim))
x = (centres[rng.integers(0, k, size=n)] + rng.normal(0, 25, size=(n, dim)))
init = x[rng.permutation(n)[:k]].copy()
lo, hi = tdist.shard_rows(n, rank, world)
xs = torch.from_numpy(x[lo:hi]).to(dev)
init_t = torch.from_numpy(init).to(dev)
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
labels, cent, inertia, iters = tdist.kmeans_fit_sharded(xs, init_t, max_iter=10)
torch.cuda.synchronize(); dist.barrier()
dt = time.perf_counter() - t0
all_labels = [torch.empty(tdist.shard_rows(n, r, world)[1] - tdist.shard_rows(n, r, world)[0], dtype=torch.int32, device=dev) for r in range(world)]
dist.all_gather(all_labels, labels)
if rank == 0:
    t0 = time.perf_counter()
    l1, c1, in1, it1 = api.kmeans_fit(torch.from_numpy(x).to(dev), k, init=init_t, max_iter=10)
    torch.cuda.synchronize()
    dt1 = time.perf_counter() - t0
    lab = torch.cat(all_labels).cpu().numpy()
    agree = float((lab == l1.cpu().numpy()).mean())
    rel = float((torch.linalg.norm(cent - c1, dim=1) / torch.linalg.norm(c1, dim=1)).max())
    print(json.dumps({"world": world, "n": n, "k": k, "iters": iters, "iters_1gpu": it1, "label_agreement": agree,
                      "max_rel_centroid_err": rel, "inertia_rel_err": abs(inertia - in1) / in1, "sharded_s": dt, "single_s": dt1}))
    assert agree > 0.9999 and rel < 1e-9 and iters == it1
# ---- config-C style run on int16 tile vectors: tensor-core assignment, points sharded, NCCL all-reduce of sums/counts
n2, k2 = int(os.environ.get("KM_N2", 1 << 21)), int(os.environ.get("KM_K2", 1 << 16))
g = torch.Generator(device=dev); g.manual_seed(5)
cen = torch.randint(-3000, 3000, (k2, 192), generator=g, device=dev, dtype=torch.int32).to(torch.float32)
lo2, hi2 = tdist.shard_rows(n2, rank, world)
g2 = torch.Generator(device=dev); g2.manual_seed(100 + rank)
own = torch.randint(0, k2, (hi2 - lo2,), generator=g2, device=dev)
xs2 = (cen[own] + 25.0 * torch.randn(hi2 - lo2, 192, generator=g2, device=dev)).round().clamp(-32768, 32767).to(torch.int16)
init2 = cen.to(torch.float64) + 3.0
tdist.kmeans_fit_i16_sharded(xs2, init2, max_iter=0)           # warm-up
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
lab2, cent2, inertia2, it2 = tdist.kmeans_fit_i16_sharded(xs2, init2, max_iter=3)
torch.cuda.synchronize(); dist.barrier()
dt2 = time.perf_counter() - t0
if rank == 0:
    print(json.dumps({"i16_config_c": {"world": world, "n": n2, "k": k2, "lloyd_updates": it2, "seconds": dt2,
                                       "evals_per_s": n2 * k2 * (it2 + 1) / dt2, "inertia": inertia2}}))
dist.destroy_process_group()
