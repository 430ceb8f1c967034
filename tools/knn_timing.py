"""k-NN timing probe (device-resident inputs, CUDA events on torch's current stream)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tiler_b200 import api, synth

n_dict = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
n_q = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 256 * 4
adv = len(sys.argv) > 3 and sys.argv[3] == 'adv'   # adversarial set: every norm < 2^29 -> the no-wrap epilogue
d = torch.from_numpy(synth.random_features(n_dict, 1, adversarial=adv)).cuda()
q = torch.from_numpy(synth.random_features(n_q, 2, adversarial=adv)).cuda()
knn = api.KnnShort(d)
res = {}
for k in (1, 64):
    for _ in range(2):
        knn.search(q, k, sorted=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for _ in range(reps):
        knn.search(q, k, sorted=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    api.profile_enable(True)
    for _ in range(reps):
        knn.search(q, k, sorted=False)
    kms, kn = api.profile_read("knn_k1" if k == 1 else "knn_topk")
    api.profile_enable(False)
    evals = n_q * n_dict / (ms * 1e-3)
    res[f"k{k}"] = {"kernel_ms": kms / max(kn, 1), "ms": ms, "evals_per_s": evals, "tflops_algorithmic": evals * 384 / 1e12}
print(json.dumps({"n_dict": n_dict, "n_q": n_q, **res}))
