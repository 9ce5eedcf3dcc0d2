"""GPU against the committed golden vectors of the REFERENCE itself (tests/golden/golden.npz, produced by
tests/golden/make_golden.py from the unmodified utility.py): masks and thresholds bit for bit, centroids within
1e-5 relative of the reference run on float64 input (same reference code, no float32 accumulation drift), labels
bit-exact when the device is given the reference's own float32 centroids."""
import os
import zlib

import numpy as np
import pytest

from . import _data as D
from . import _parity as P

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.npz")


@pytest.fixture(scope="module")
def G():
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def U():
    from neural_network_compression_b200.common import utility

    return utility


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def test_prune_matches_reference_golden(U, G):
    for name, w, q in D.prune_cases():
        w = w.copy()
        mask = U.prune_weigth(w, q)
        assert U.prune_weigth.last_threshold == float(G["prune/%s/thr" % name]), name
        assert int(mask.sum()) == int(G["prune/%s/n_pruned" % name]), name
        assert crc(mask) == G["prune/%s/mask_crc" % name], name
        assert crc(w) == G["prune/%s/w_crc" % name], name


def test_kmeans_matches_reference_golden(U, G):
    """Every golden k-means case against BOTH reference runs (float32 and float64 input), none skipped: the class of
    each case (strict / envelope / multiset) and its asserted bound are documented in tests/_parity.py."""
    report = []
    for name, w, bits, mode, seed in D.kmeans_cases(big=True):
        cdfs = None
        if mode == "density":
            cdfs = U.get_weight_distribution(w, skip_zeros=True)  # fused survivor selection (trainer.py:55-60)
            assert cdfs[0].tobytes() == G["km/%s/xnew" % name].tobytes(), name
            assert cdfs[1].tobytes() == G["km/%s/cdf" % name].tobytes(), name
        np.random.seed(seed)
        ris, km = U.get_quantized_weight(w, bits, mode, cdfs)
        assert km.n_clusters == G["km/%s/init" % name].size, name
        assert np.array_equal(np.bincount(km.labels_, minlength=km.n_clusters), km.code_histogram), name
        report.append(P.check_centroids(name, G, km.cluster_centers_, km.n_iter_, crc(km.labels_), km.code_histogram))
        # the dense result is the codebook gathered by the labels (utility.py:239)
        assert np.array_equal(ris.ravel(), km.cluster_centers_.ravel()[km.labels_]), name
    print("\n".join(report))


def test_labels_exact_given_reference_centroids(U, G):
    """north_star: masks and cluster indices bit-exact given identical centroids.  The device E-step fed with the
    reference's OWN final centroids (the centred values sklearn iterated on and its X_mean, both recorded by
    make_golden.py) reproduces the reference's labels_ and code histogram on every golden case -- no allow-list."""
    for name, w, bits, mode, seed in D.kmeans_cases(big=True):
        c32 = G["km/%s/f32/centers" % name]
        ref_km = U.KMeansResult(c32.reshape(-1, 1), None, 0, 0.0, centred_centers=G["km/%s/f32/centred" % name],
                                mean=np.float32(G["km/%s/f32/mean" % name]), code_bits=U.index_bits(c32.size))
        labels, packed, hist = U.assign_codes(w, ref_km)
        assert crc(labels) == G["km/%s/f32/labels_crc" % name], name
        assert np.array_equal(hist, G["km/%s/f32/hist" % name]), name
        # dequantised tensor = the reference's `ris` (cluster_centers_[labels_]) byte for byte
        ris = U.dequantize(packed, w.size, ref_km.code_bits, c32)
        assert crc(ris.reshape(w.shape)) == G["km/%s/f32/ris_crc" % name], name


@pytest.mark.parametrize("n", [1 << 26])
def test_full_size_properties(U, n):
    """Size-independent properties on a large device-resident layer (the oracle cannot run at this size)."""
    import torch

    g = torch.Generator(device="cuda").manual_seed(99)
    t = torch.empty(n, device="cuda").normal_(0.0, 0.02, generator=g)
    orig = t.clone()
    mask, km = U.compress_weight(t, 1.0, True, 8, "linear")
    thr = np.float32(U.prune_weigth.last_threshold)
    # mask == |w| < thr, pruned in place, threshold = np.std * q recomputed by torch in float64 within 1e-6
    assert bool(torch.equal(mask, orig.abs() < float(thr)))
    assert bool(torch.equal(t, torch.where(mask, torch.zeros_like(orig), orig)))
    assert abs(float(thr) / float(orig.double().std(unbiased=False)) - 1.0) < 1e-6
    assert int(mask.sum()) == U.prune_weigth.last_pruned
    # codes: histogram sums to n, decode(pack) are codebook values, every weight sits at a nearest centroid
    assert int(km.code_histogram.sum()) == n and km.n_nonzero == n - U.prune_weigth.last_pruned
    codes = km.packed_codes.to(torch.int64)
    assert bool(torch.equal(torch.bincount(codes, minlength=256).cpu(), torch.from_numpy(km.code_histogram)))
    cent = torch.from_numpy(km.cluster_centers_.ravel()).cuda()
    deq = U.dequantize(km.packed_codes, n, 8, km.cluster_centers_)
    assert bool(torch.equal(deq, cent[codes]))
    sample = torch.randint(0, n, (1 << 16,), device="cuda")
    d_all = (t[sample, None].double() - cent[None, :].double()).abs()
    d_own = (t[sample].double() - deq[sample].double()).abs()
    # the label rule is float32 arithmetic: at a cell boundary it may pick the other neighbour, ~1e-6 farther
    assert bool((d_own <= d_all.min(dim=1).values + 2e-6).all())
    # idempotence of pruning on the pruned tensor with the same absolute threshold
    t2 = t.clone()
    m2 = U.prune_weigth(t2, float(thr), std_smooth=False)
    assert bool(torch.equal(t2, t)) and bool(torch.equal(m2, t == 0))
