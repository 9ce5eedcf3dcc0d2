"""Pins the CPU oracle (oracle/nnc_oracle.c) to the reference: every golden vector in tests/golden/golden.npz was
produced by the UNMODIFIED reference helpers (tests/golden/make_golden.py).  CPU only."""
import os
import zlib

import numpy as np
import pytest

from oracle import oracle as O
from . import _data as D
from . import _parity as P

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.npz")


@pytest.fixture(scope="module")
def G():
    return np.load(GOLDEN)


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


@pytest.mark.parametrize("n", [1, 5, 8, 100, 128, 129, 1000, 4096, 4097, 100003, 1 << 20])
def test_pairwise_sum_and_std_match_numpy(n):
    w = D.gaussian(n, seed=n, sigma=0.05) + np.float32(0.01)
    assert O.pairwise_sum(w).tobytes() == np.add.reduce(w).tobytes()
    m, v, s = O.std(w)
    assert (m.tobytes(), v.tobytes(), s.tobytes()) == (np.mean(w).tobytes(), np.var(w).tobytes(), np.std(w).tobytes())


def test_linspace_matches_numpy():
    rng = np.random.RandomState(0)
    for _ in range(50):
        a, b = np.float32(rng.randn()), np.float32(rng.randn())
        for num in (2, 16, 32, 256, 300):
            assert O.linspace_f32(a, b, num).tobytes() == np.linspace(a, b, num=num).tobytes()
    assert O.linspace_f32(0.5, 0.5, 8).tobytes() == np.linspace(np.float32(0.5), np.float32(0.5), 8).tobytes()


def test_prune_golden(G):
    for name, w, q in D.prune_cases():
        w = w.copy()
        mask = O.prune_weigth(w, q)
        assert O.prune_weigth.last_threshold == float(G["prune/%s/thr" % name]), name
        assert int(mask.sum()) == int(G["prune/%s/n_pruned" % name]), name
        assert crc(mask) == G["prune/%s/mask_crc" % name], name
        assert crc(w) == G["prune/%s/w_crc" % name], name
        if w.size <= 1000:
            assert np.array_equal(mask, G["prune/%s/mask" % name])


def test_prune_semantics():
    w = np.array([0.5, -0.5, 0.25, -0.0, np.nan, 1.0], dtype=np.float32)
    m = O.prune_weigth(w, 0.5, std_smooth=False)
    assert m.tolist() == [False, False, True, True, False, False]  # strict <, NaN kept
    w = D.gaussian(100, seed=1)
    assert not O.prune_weigth(w, 0).any()  # threshold 0 prunes nothing (le_net_300_100_trainer.py:26)


def test_kmeans_golden(G):
    """REF32 mode == the reference (float32 input, one OpenMP thread) bit for bit on every golden case, the
    benchmark-type case c4_4m_linear8 included."""
    for name, w, bits, mode, seed in D.kmeans_cases(big=True):
        cdfs = None
        if mode == "density":
            nz = O.compact_nonzero(w)
            assert nz.size == int(G["km/%s/n_nz" % name])
            cdfs = O.get_weight_distribution(nz)
            assert cdfs[0].tobytes() == G["km/%s/xnew" % name].tobytes(), name
            assert cdfs[1].tobytes() == G["km/%s/cdf" % name].tobytes(), name
        idx = None
        if mode == "forgy":
            np.random.seed(seed)
            idx = np.random.randint(0, w.size, size=2 ** bits)
        space = O.init_centroids(w, bits, mode, cdfs, idx)
        assert space.tobytes() == G["km/%s/init" % name].tobytes(), name
        km = O.kmeans1d(w, space, mode=O.MODE_REF32)
        # bit-exact with the reference at one OpenMP thread
        assert km.n_iter_ == int(G["km/%s/f32/n_iter" % name]), name
        assert km.cluster_centers_.ravel().tobytes() == G["km/%s/f32/centers" % name].tobytes(), name
        assert km.centred_centers.tobytes() == G["km/%s/f32/centred" % name].tobytes(), name
        assert km.mean.tobytes() == G["km/%s/f32/mean" % name].tobytes(), name
        assert crc(km.labels_) == G["km/%s/f32/labels_crc" % name], name
        ris = km.cluster_centers_[km.labels_].reshape(w.shape)
        assert crc(ris) == G["km/%s/f32/ris_crc" % name], name
        assert np.isclose(km.inertia_, float(G["km/%s/f32/inertia" % name]), rtol=1e-6), name


def test_det_mode_reference_parity_every_case(G):
    """The device semantic (DET: exact per-cluster sums, deterministic far-point order) against BOTH reference runs
    on every golden case -- no case skipped; the class of each case and its bound are in tests/_parity.py."""
    O.set_threads(os.cpu_count() or 1)
    report = []
    for name, w, bits, mode, seed in D.kmeans_cases(big=True):
        det = O.kmeans1d(w, G["km/%s/init" % name], mode=O.MODE_DET)
        hist = np.bincount(det.labels_, minlength=det.cluster_centers_.shape[0])
        report.append(P.check_centroids(name, G, det.cluster_centers_, det.n_iter_, crc(det.labels_), hist))
    print("\n".join(report))
    kinds = [P.classify(c[0]) for c in D.kmeans_cases(big=True)]
    assert kinds.count("strict") == 16 and kinds.count("envelope") == 4 and kinds.count("multiset") == 3


def test_labels_exact_given_reference_centroids(G):
    """north_star: indices bit-exact given identical centroids.  The label rule (sklearn's float32 expression in the
    centred space, lowest index on ties) applied to the reference's OWN final centred centroids reproduces the
    reference's labels_ on every golden case -- no allow-list."""
    for name, w, bits, mode, seed in D.kmeans_cases(big=True):
        lab = O.assign(w, G["km/%s/f32/centred" % name], G["km/%s/f32/mean" % name])
        assert crc(lab) == G["km/%s/f32/labels_crc" % name], name
        assert np.array_equal(np.bincount(lab, minlength=G["km/%s/init" % name].size), G["km/%s/f32/hist" % name]), name


def test_det_equals_ref32_arithmetic_with_same_far_order(G):
    """Closes the MULTISET class: given the SAME far-point order (the deterministic one), the reference's float32
    arithmetic (REF32) and the exact-sum semantic (DET) take the same number of iterations, relocate the same number
    of samples and end with id-aligned centroids within 1e-4 -- what separates the device from the reference in those
    cases is np.argpartition's implementation-defined order / tie choice, nothing else."""
    O.set_threads(os.cpu_count() or 1)
    for name, w, bits, mode, seed in D.kmeans_cases(big=True):
        if P.classify(name) != "multiset":
            continue
        space = G["km/%s/init" % name]
        det = O.kmeans1d(w, space, mode=O.MODE_DET)
        r32 = O.kmeans1d(w, space, mode=O.MODE_REF32, numpy_far_order=False)
        assert det.n_iter_ == r32.n_iter_ and det.n_relocations == r32.n_relocations, name
        scale = np.abs(r32.cluster_centers_).max()
        assert P.rel_err(det.cluster_centers_.ravel(), r32.cluster_centers_.ravel(), scale) <= 1e-4, name
        assert (det.labels_ != r32.labels_).mean() <= 1e-3, name


def test_pack_roundtrip_and_segsum():
    rng = np.random.RandomState(0)
    for bits in (1, 2, 3, 4, 5, 8, 9, 12):
        lab = rng.randint(0, 2 ** bits, size=1001).astype(np.int32)
        p = O.pack_codes(lab, bits)
        assert p.size == (1001 * bits + 7) // 8
        assert np.array_equal(O.unpack_codes(p, 1001, bits), lab)
    g = (rng.randn(5000) * 1e-3).astype(np.float32)
    lab = rng.randint(0, 7, size=5000).astype(np.int32)
    a = O.grad_segsum(g, lab, 7)
    b = O.grad_segsum(g, lab, 7, fixed=True)
    assert np.allclose(a, np.bincount(lab, weights=g.astype(np.float64), minlength=7), rtol=1e-13)
    assert np.allclose(a, b, rtol=1e-6, atol=1e-9)
