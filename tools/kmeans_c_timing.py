"""Config-C style isolation of the k-means Lloyd loop on int16 tile vectors (BASELINE.json configs[2], scaled by argv):
n points x 192 -> k centroids; assignment on the tensor cores + exact f64 decision, ordered update."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tiler_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 16
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rng = np.random.default_rng(7)
centres = synth.random_features(k, 3, adversarial=True).astype(np.float32)
x = torch.from_numpy(centres).cuda()[torch.from_numpy(rng.integers(0, k, size=n)).cuda()]
x = (x + 25.0 * torch.randn(n, 192, device="cuda")).round().clamp(-32768, 32767).to(torch.int16)
init = x[torch.randperm(n, device="cuda")[:k]].to(torch.float64)
api.kmeans_fit_i16(x, k, init, max_iter=1)      # warm-up: grows the allocation pool, loads the kernels
torch.cuda.synchronize()
api.profile_enable(True)
t0 = time.perf_counter()
labels, cent, inertia, it, amb = api.kmeans_fit_i16(x, k, init, max_iter=iters)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
knn_ms, knn_n = api.profile_read("knn_k4")
parts = {nm: api.profile_read(nm)[0] for nm in ("km_rerank", "knn_topk", "km_rerank64", "km_amb", "km_update")}
evals = n * k * (it + 1)
print(json.dumps({"n": n, "k": k, "lloyd_updates": it, "assign_passes": it + 1, "seconds": dt, "ambiguous_points": amb,
                  "knn_kernel_ms_per_pass": knn_ms / max(knn_n, 1), "evals_per_s_total": evals / dt,
                  "evals_per_s_knn_kernel": n * k / (knn_ms / max(knn_n, 1) * 1e-3), "inertia": inertia, "ms_total_by_kernel": {"knn_k4": knn_ms, **parts}}))
