"""Generates tests/golden/golden.npz by executing the UNMODIFIED reference helpers
(/root/reference/neural_network_compression/common/utility.py through oracle/ref_shim.py) on the seeded inputs
of tests/_data.py.  Run in the build container (the reference tree is not on the GPU box):

    OMP_NUM_THREADS=1 python -m tests.golden.make_golden                  # everything (about 15 minutes)
    OMP_NUM_THREADS=1 python -m tests.golden.make_golden --only c4        # (re)generate the cases whose name contains
                                                                          # "c4" and merge them into the existing file

Besides the public results the generator records, per k-means case, the CENTRED centroids sklearn iterated on (the
value `_kmeans_single_lloyd` returns before `KMeans.fit` adds X_mean back, sklearn/cluster/_kmeans.py:1546) by wrapping
that sklearn function from the outside; the reference file itself stays unmodified.  labels_ are a function of those
centred centroids, not of cluster_centers_ (the +mean / -mean round trip is not always exact in float32).

One OpenMP thread: scikit-learn's float32 accumulation order is otherwise nondeterministic.
Every k-means case is run twice: as the reference runs it (float32) and with float64 input -- same reference
code, sklearn keeps float64 -- which is the accumulation-error-free statement of the same algorithm.
"""
import os
import sys
import zlib

os.environ.setdefault("OMP_NUM_THREADS", "1")
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from tests import _data as D  # noqa: E402


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def main():
    assert os.environ.get("OMP_NUM_THREADS") == "1", "run with OMP_NUM_THREADS=1"
    ref = ref_shim.load()
    warnings.simplefilter("ignore")
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.npz")
    G = {}
    if only:
        with np.load(out) as old:
            G = {key: old[key] for key in old.files}
    # recorder around sklearn's single-run Lloyd: the centred centroids and whether the stop was strict
    import sklearn.cluster._kmeans as skm

    rec = {}
    inner = skm._kmeans_single_lloyd

    def recording_lloyd(*a, **k):
        res = inner(*a, **k)
        rec["centred"] = np.array(res[2], copy=True).ravel()
        return res

    skm._kmeans_single_lloyd = recording_lloyd
    # ---- pruning (utility.py:134-163)
    for name, w, q in D.prune_cases():
        if only and only not in name:
            continue
        w = w.copy()
        thr = np.std(w) * q
        mask = ref.prune_weigth(w, q, True)
        G["prune/%s/thr" % name] = np.float64(thr)
        G["prune/%s/n_pruned" % name] = np.int64(mask.sum())
        G["prune/%s/mask_crc" % name] = crc(mask)
        G["prune/%s/w_crc" % name] = crc(w)
        if w.size <= 1000:
            G["prune/%s/mask" % name] = mask
    # ---- weight distribution + init + k-means (utility.py:334-392, 172-240; trainer.py:55-69)
    for name, w, bits, mode, seed in D.kmeans_cases():
        if only and only not in name:
            continue
        cdfs = None
        if mode == "density":
            flat = w.flatten()
            (zero_idx,) = np.nonzero(flat == 0)
            nz = np.delete(flat, zero_idx, axis=0)  # trainer.py:55-59
            cdfs = ref.get_weight_distribution(nz)
            G["km/%s/xnew" % name] = cdfs[0]
            G["km/%s/cdf" % name] = cdfs[1]
            G["km/%s/n_nz" % name] = np.int64(nz.size)
        for tag, x in (("f32", w), ("f64", w.astype(np.float64))):
            np.random.seed(seed)
            ris, km = ref.get_quantized_weight(x, bits, mode, cdfs)
            G["km/%s/%s/centers" % (name, tag)] = km.cluster_centers_.ravel()
            G["km/%s/%s/centred" % (name, tag)] = rec["centred"]
            G["km/%s/%s/mean" % (name, tag)] = x.reshape(-1, 1).mean(axis=0)[0]  # as KMeans.fit computes it (_kmeans.py:1486)
            G["km/%s/%s/n_iter" % (name, tag)] = np.int64(km.n_iter_)
            G["km/%s/%s/inertia" % (name, tag)] = np.float64(km.inertia_)
            G["km/%s/%s/labels_crc" % (name, tag)] = crc(km.labels_.astype(np.int32))
            G["km/%s/%s/hist" % (name, tag)] = np.bincount(km.labels_, minlength=km.cluster_centers_.shape[0]).astype(np.int64)
            if tag == "f32":
                G["km/%s/f32/ris_crc" % name] = crc(ris)
                G["km/%s/init" % name] = np.asarray(km.init, dtype=np.float32).ravel()
                if w.size <= 1000:
                    G["km/%s/f32/labels" % name] = km.labels_.astype(np.int32)
        print(name, "k", km.cluster_centers_.shape[0], "iters f32/f64", int(G["km/%s/f32/n_iter" % name]), int(G["km/%s/f64/n_iter" % name]), flush=True)
    import sklearn
    import scipy

    G["meta/versions"] = np.array(["numpy " + np.__version__, "sklearn " + sklearn.__version__, "scipy " + scipy.__version__])
    np.savez_compressed(out, **G)
    print("wrote", out, os.path.getsize(out), "bytes,", len(G), "entries")


if __name__ == "__main__":
    main()
