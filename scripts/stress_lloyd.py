"""Repeats k-means fits that exercise relocation out of multi-count entries and reports whether the results are
stable run to run, per Lloyd path (cluster kernel / cooperative loop).  Debug aid (GPU box)."""
import os, sys, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from neural_network_compression_b200.common import utility as U

def crc(a): return zlib.crc32(np.ascontiguousarray(a).tobytes())
rng = np.random.RandomState(9)
vals = np.sort(((0.5 + 1.5 * rng.rand(20)) * rng.choice([-1.0, 1.0], size=20)).astype(np.float32))
w1 = vals[rng.randint(0, 20, size=100 * 1000)]
g = (np.random.RandomState(3).randn(300000) * 0.02).astype(np.float32)
g[np.abs(g) < 0.02] = 0
rng2 = np.random.RandomState(9)
for _ in range(1):  # the draws of tests/test_gpu_parity.py::test_kmeans_hist_path_multiplicities up to its last tensor
    g0 = (np.random.RandomState(21).randn(200000) * 0.02).astype(np.float32)
lobes = ((1.0 + rng.rand(300 * 1000) * 2.0 ** -12) * rng.choice([-1.0, 1.0, 0.0], size=300 * 1000)).astype(np.float32)
narrow = (1.0 + rng.rand(500 * 1000) * 2.0 ** -11).astype(np.float32)
cases = [("dup20_linear5", w1, 5, "linear"), ("dup20_forgy6", w1, 6, "forgy"), ("pruned_linear8", g, 8, "linear"),
         ("lobes_linear8", lobes, 8, "linear"), ("lobes_linear3", lobes, 3, "linear"), ("narrow_linear4", narrow, 4, "linear")]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 15
for path in ("cluster", "coop"):
    if path == "coop":
        os.environ["NNC_LLOYD_NO_CLUSTER"] = "1"
    for name, w, bits, mode in cases:
        seen = {}
        for r in range(reps):
            np.random.seed(1)
            ris, km = U.get_quantized_weight(w, bits, mode)
            key = (km.n_iter_, km.n_relocations, crc(km.cluster_centers_), crc(km.labels_))
            seen[key] = seen.get(key, 0) + 1
        print(path, name, "distinct results:", len(seen), seen, flush=True)
