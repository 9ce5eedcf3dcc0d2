"""Pins the CPU oracle (oracle/nnc_oracle.c) to the reference: every golden vector in tests/golden/golden.npz was
produced by the UNMODIFIED reference helpers (tests/golden/make_golden.py).  CPU only."""
import os
import zlib

import numpy as np
import pytest

from oracle import oracle as O
from . import _data as D

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


@pytest.mark.parametrize("big", [False])
def test_kmeans_golden(G, big):
    for name, w, bits, mode, seed in D.kmeans_cases(big=False):
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
        assert crc(km.labels_) == G["km/%s/f32/labels_crc" % name], name
        ris = km.cluster_centers_[km.labels_].reshape(w.shape)
        assert crc(ris) == G["km/%s/f32/ris_crc" % name], name
        assert np.isclose(km.inertia_, float(G["km/%s/f32/inertia" % name]), rtol=1e-6), name


def test_det_mode_tracks_float64_reference(G):
    """The device semantic (exact per-cluster sums) against the same reference code run on float64 input."""
    worst = 0.0
    for name, w, bits, mode, seed in D.kmeans_cases(big=False):
        space = G["km/%s/init" % name]
        det = O.kmeans1d(w, space, mode=O.MODE_DET)
        c64 = G["km/%s/f64/centers" % name]
        if det.n_iter_ != int(G["km/%s/f64/n_iter" % name]) or det.n_relocations:
            continue  # relocation order / float32-vs-float64 label ties are compared in the GPU golden test
        err = np.abs(det.cluster_centers_.ravel().astype(np.float64) - c64).max() / np.abs(c64).max()
        worst = max(worst, err)
        assert err <= 1e-5, (name, err)
    assert worst > 0  # at least one case was compared


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
