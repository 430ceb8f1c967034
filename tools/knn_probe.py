"""One k-NN search on random features (device-resident): the smallest program that shows the k-NN kernels to ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tiler_b200 import api, synth

n_dict = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
n_q = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 128 * 4
k = int(sys.argv[3]) if len(sys.argv) > 3 else 64
d = torch.from_numpy(synth.random_features(n_dict, 1)).cuda()
q = torch.from_numpy(synth.random_features(n_q, 2)).cuda()
knn = api.KnnShort(d)
for _ in range(2):
    knn.search(q, k, sorted=False)
torch.cuda.synchronize()
print("ok")
