"""GPU parity tests: the CUDA path (through the C ABI / the utility.py mirror) against the CPU oracle.

Bars (BASELINE.json north_star): masks and cluster indices bit-exact; centroids within 1e-5 relative of the
reference semantics (oracle REF32 mode, itself pinned to the reference by tests/golden) and BIT-EXACT against
the oracle's DET mode (the device's order-independent fixed-point accumulation).
"""
import numpy as np
import pytest

from oracle import oracle as O
from . import _data as D

pytestmark = pytest.mark.gpu

CENTROID_RTOL = 1e-5  # north_star: "Centroids must match within 1e-5 relative"


@pytest.fixture(scope="module")
def U():
    from neural_network_compression_b200.common import utility

    return utility


# ---------------------------------------------------------------------------------------------------------
# np.std restatement
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 7, 8, 9, 127, 128, 129, 1000, 4095, 4096, 4097, 8193, 100003, (1 << 20) + 5, 3 * 1000 * 1000 + 1])
def test_stats_bit_exact_vs_numpy(U, n):
    w = D.gaussian(n, seed=n % 9973, sigma=0.05) + np.float32(0.01)
    m, v, s = U.weight_stats(w)
    assert m.tobytes() == np.mean(w).tobytes()
    assert v.tobytes() == np.var(w).tobytes()
    assert s.tobytes() == np.std(w).tobytes()


def test_stats_16m(U):
    w = D.gaussian(4096 * 4096)
    m, v, s = U.weight_stats(w)
    assert (m, v, s) == (np.mean(w), np.var(w), np.std(w))


# ---------------------------------------------------------------------------------------------------------
# pruning
# ---------------------------------------------------------------------------------------------------------
def _check_prune(U, w, q, std_smooth=True):
    w_dev, w_ora, w_np = w.copy(), w.copy(), w.copy()
    mask_dev = U.prune_weigth(w_dev, q, std_smooth)
    mask_ora = O.prune_weigth(w_ora, q, std_smooth)
    # the reference expression itself (utility.py:159-162), NumPy being the real thing here
    thr = np.std(w_np) * q if std_smooth else q
    mask_np = np.abs(w_np) < thr
    w_np[mask_np] = 0
    assert mask_dev.dtype == np.bool_ and mask_dev.shape == w.shape
    assert np.array_equal(mask_dev, mask_np)
    assert np.array_equal(mask_dev, mask_ora)
    assert w_dev.tobytes() == w_np.tobytes()
    # NEP 50: a Python-float threshold is compared in float32
    assert U.prune_weigth.last_threshold == (float(thr) if isinstance(thr, np.floating) else float(np.float32(thr)))
    assert U.prune_weigth.last_pruned == int(mask_np.sum())


def test_prune_lenet300(U):
    for kind in ("glorot", "gauss"):
        for name, w, b, (qw, qb) in D.lenet300_tensors(kind=kind):
            _check_prune(U, w, qw)
            _check_prune(U, b, qb)  # includes threshold 0 on the output bias: strict < prunes nothing


def test_prune_lenet5(U):
    for name, w, b, (qw, qb) in D.lenet5_tensors():
        _check_prune(U, w, qw)
        _check_prune(U, b, qb)


@pytest.mark.parametrize("n", [1, 3, 5, 31, 1023, 4099, 65537, 1 << 20])
@pytest.mark.parametrize("q", [0.25, 1, 2.5])
def test_prune_sizes(U, n, q):
    _check_prune(U, D.gaussian(n, seed=n), q)


@pytest.mark.parametrize("ratio", [20.0, 300.0, 5000.0])
def test_prune_mean_far_from_zero(U, ratio):
    """|mean| >> std (normalisation scales, biases): the threshold sits in the middle of the data and the estimate behind the
    speculation band must not lose the variance in sum(x^2) / n - mean^2.  Bit-identical mask, no
    speculation failure, also through the fused compress path."""
    rng = np.random.default_rng(int(ratio))
    for n in (1 << 20, 300_001):
        w = (1.0 + rng.standard_normal(n) / ratio).astype(np.float32)
        _check_prune(U, w, float(ratio))          # threshold ~ |mean|: about half of the elements go
        _check_prune(U, w, float(ratio) * 0.999)
        w2, w3 = w.copy(), w.copy()
        mask, km = U.compress_weight(w2, float(ratio), True, 4, "linear")
        assert np.array_equal(mask.reshape(w.shape), O.prune_weigth(w3, float(ratio)))
        assert w2.tobytes() == w3.tobytes()


def test_prune_hard_threshold_and_f64_scalar(U):
    w = D.gaussian(50001, seed=3)
    _check_prune(U, w, 0.01, std_smooth=False)
    _check_prune(U, w, np.float64(0.7))
    _check_prune(U, w, np.float32(0.7))
    _check_prune(U, w, 0)


def test_prune_repeated_on_sparse(U):
    # the trainer re-prunes the already sparse tensor every batch (trainer.py:144-148)
    w = D.gaussian(235200, seed=5, sigma=0.05).reshape(784, 300)
    for _ in range(4):
        _check_prune(U, w, 1)
        m = U.prune_weigth(w, 1)
        assert m.sum() > 0


def test_prune_special_values(U):
    w = D.gaussian(10007, seed=11)
    w[5] = -0.0
    w[6] = 0.0
    w[100:200] = 0.0
    _check_prune(U, w, 1)
    w = np.zeros(300, dtype=np.float32)  # all-zero bias: std 0, nothing is < 0
    _check_prune(U, w, 1)
    w = np.full(1000, 0.25, dtype=np.float32)
    _check_prune(U, w, 1)


def test_prune_torch_device_tensor(U):
    import torch

    w = D.gaussian(1 << 20, seed=21).reshape(1024, 1024)
    ref = w.copy()
    mask_ref = O.prune_weigth(ref, 1)
    t = torch.from_numpy(w.copy()).cuda()
    mask = U.prune_weigth(t, 1)
    assert mask.is_cuda and mask.dtype == torch.bool and mask.shape == t.shape
    assert np.array_equal(mask.cpu().numpy(), mask_ref)
    assert t.cpu().numpy().tobytes() == ref.tobytes()
    # unaligned device views take the scalar path
    t2 = torch.from_numpy(w.copy()).cuda().reshape(-1)[1:100000]
    ref2 = w.reshape(-1)[1:100000].copy()
    m2 = U.prune_weigth(t2, 1)
    mr2 = O.prune_weigth(ref2, 1)
    assert np.array_equal(m2.cpu().numpy(), mr2)
    assert t2.cpu().numpy().tobytes() == ref2.tobytes()


def test_mask_apply(U):
    w = D.gaussian(100003, seed=8)
    mask = np.abs(w) < 0.01
    g = D.gaussian(100003, seed=9)
    expect = g.copy()
    expect[mask] = 0
    U.apply_mask(g, mask)
    assert g.tobytes() == expect.tobytes()


# ---------------------------------------------------------------------------------------------------------
# survivor selection, weight CDF
# ---------------------------------------------------------------------------------------------------------
def test_nonzero_and_cdf(U):
    for n in [40, 1000, 30000, 235200, 1 << 20]:
        w = D.gaussian(n, seed=n + 1, sigma=0.05)
        O.prune_weigth(w, 1)
        nz = U.nonzero_weights(w)
        nz_ref = w[w != 0]
        assert nz.tobytes() == nz_ref.tobytes()
        x_ref, y_ref = O.get_weight_distribution(nz_ref)
        x, y = U.get_weight_distribution(nz_ref)
        assert x.tobytes() == x_ref.tobytes() and y.tobytes() == y_ref.tobytes()
        x2, y2 = U.get_weight_distribution(w, skip_zeros=True)  # fused survivor selection
        assert x2.tobytes() == x_ref.tobytes() and y2.tobytes() == y_ref.tobytes()


# ---------------------------------------------------------------------------------------------------------
# k-means
# ---------------------------------------------------------------------------------------------------------
def _init_for(w, bits, mode, seed=0):
    if mode == "linear":
        return O.init_centroids(w, bits, "linear"), None
    if mode == "density":
        nz = w.ravel()[w.ravel() != 0]
        cdfs = O.get_weight_distribution(nz)
        return O.init_centroids(w, bits, "density", cdfs), cdfs
    if mode == "forgy":
        np.random.seed(seed)
        idx = np.random.randint(0, w.size, size=2 ** bits)
        return O.init_centroids(w, bits, "forgy", forgy_indices=idx), None
    raise ValueError(mode)


def _check_kmeans(U, w, bits, mode, seed=0, check_ref32=True):
    space, cdfs = _init_for(w, bits, mode, seed)
    if mode == "forgy":
        np.random.seed(seed)
    ris, km = U.get_quantized_weight(w, bits, mode, cdfs)
    k = space.size
    assert km.cluster_centers_.shape == (k, 1) and km.cluster_centers_.dtype == np.float32
    assert km.labels_.dtype == np.int32 and km.labels_.shape == (w.size,)
    assert ris.shape == w.shape and ris.dtype == np.float32
    # --- bit-exact against the device semantic (oracle DET mode)
    det = O.kmeans1d(w, space, mode=O.MODE_DET)
    assert km.n_iter_ == det.n_iter_, (km.n_iter_, det.n_iter_)
    assert km.cluster_centers_.tobytes() == det.cluster_centers_.tobytes()
    assert np.array_equal(km.labels_, det.labels_)
    assert km.inertia_ == det.inertia_
    assert km.strict_convergence == det.strict and km.n_relocations == det.n_relocations
    assert km.mean == det.mean and np.float32(km.tol_) == det.tol
    # --- utility.py:239
    assert ris.tobytes() == km.cluster_centers_[km.labels_].reshape(w.shape).tobytes()
    # --- packed codes / histogram / decode round trip
    assert np.array_equal(O.unpack_codes(km.packed_codes, w.size, km.code_bits), km.labels_)
    assert km.packed_codes.tobytes() == O.pack_codes(km.labels_, km.code_bits).tobytes()
    assert np.array_equal(km.code_histogram, np.bincount(km.labels_, minlength=k))
    assert U.dequantize(km.packed_codes, w.size, km.code_bits, km.cluster_centers_).tobytes() == ris.tobytes()
    # --- against the reference semantic (sequential float32 sums, sklearn at one thread)
    if check_ref32:
        ref = O.kmeans1d(w, space, mode=O.MODE_REF32)
        scale = np.abs(ref.cluster_centers_).max()
        # the reference's own sequential float32 sums drift beyond 1e-5 on big clusters;
        # large tensors are compared with the float64 run of the reference in tests/test_gpu_golden.py
        if ref.n_relocations == 0 and det.n_relocations == 0 and w.size <= 65536:
            np.testing.assert_allclose(km.cluster_centers_, ref.cluster_centers_, rtol=CENTROID_RTOL, atol=CENTROID_RTOL * scale)
        # indices: bit-exact given identical centroids (north_star) -- feed the reference's centroids
        ref_km = U.KMeansResult(ref.cluster_centers_, None, 0, 0.0, centred_centers=ref.centred_centers, mean=ref.mean,
                                code_bits=km.code_bits)
        labels, packed, hist = U.assign_codes(w, ref_km)
        if ref.strict and ref.n_relocations:
            pass  # labels of a strict stop belong to the pre-relocation centroids, which the result does not expose
        else:
            assert np.array_equal(labels, ref.labels_)
            assert np.array_equal(hist, np.bincount(ref.labels_, minlength=k))
    return km


def test_kmeans_lenet300_density2(U):
    # config 1: prune with the trainer's thresholds, then 2-bit density k-means (k = 5) on the full tensors
    for name, w, b, (qw, qb) in D.lenet300_tensors():
        O.prune_weigth(w, qw)
        O.prune_weigth(b, qb)
        _check_kmeans(U, w, 2, "density")
        _check_kmeans(U, b, 2, "density")


def test_kmeans_lenet5_linear4(U):
    # config 2
    for name, w, b, (qw, qb) in D.lenet5_tensors():
        O.prune_weigth(w, qw)
        O.prune_weigth(b, qb)
        _check_kmeans(U, w, 4, "linear")
        if b.size >= 17:
            _check_kmeans(U, b, 4, "linear")
        else:
            out, km = U.get_quantized_weight(b, 4, "linear")
            assert out is b and km is None


@pytest.mark.parametrize("bits,mode", [(5, "forgy"), (8, "linear"), (3, "density"), (8, "density"), (1, "linear")])
def test_kmeans_pruned_gaussian(U, bits, mode):
    w = D.gaussian(300 * 1000, seed=77).reshape(300, 1000)
    O.prune_weigth(w, 1)
    _check_kmeans(U, w, bits, mode)


@pytest.mark.parametrize("bits,mode", [(4, "linear"), (5, "forgy"), (2, "density")])
def test_kmeans_dense_gaussian(U, bits, mode):
    w = D.gaussian(200 * 1000, seed=78)
    _check_kmeans(U, w, bits, mode, seed=3)


def test_kmeans_edge_cases(U):
    # all-equal tensor (zero bias): one distinct label, all centroids equal
    w = np.zeros(300, dtype=np.float32)
    _check_kmeans(U, w, 2, "linear")
    w = np.full(1000, 0.5, dtype=np.float32)
    _check_kmeans(U, w, 3, "linear")
    # tiny tensors around the guard n < 2^bits + 1
    w = D.gaussian(17, seed=1)
    _check_kmeans(U, w, 4, "linear")
    out, km = U.get_quantized_weight(D.gaussian(16, seed=1), 4, "linear")
    assert km is None
    # few distinct values
    w = np.repeat(np.array([-1.0, -0.5, 0.0, 0.25, 2.0], dtype=np.float32), 50)
    np.random.RandomState(0).shuffle(w)
    _check_kmeans(U, w, 2, "linear")
    _check_kmeans(U, w, 3, "linear")
    # non-zero mean, wide dynamic range
    w = (D.gaussian(50000, seed=4, sigma=1.0) * 100 + 1000).astype(np.float32)
    _check_kmeans(U, w, 4, "linear")


# ---------------------------------------------------------------------------------------------------------
# key-histogram path (csrc/khist.cu): the sorted survivors as (value, multiplicity) runs.  Chosen by the library
# for large narrow-range layers; forced here (NNC_SORT_PATH=hist) on tensors the oracle finishes in seconds.
# ---------------------------------------------------------------------------------------------------------
@pytest.fixture
def hist_path(monkeypatch):
    monkeypatch.setenv("NNC_SORT_PATH", "hist")


@pytest.mark.parametrize("bits,mode", [(5, "forgy"), (8, "linear"), (3, "density"), (8, "density"), (1, "linear"), (4, "linear")])
def test_kmeans_hist_path_pruned_gaussian(U, hist_path, bits, mode):
    w = D.gaussian(300 * 1000, seed=77).reshape(300, 1000)
    O.prune_weigth(w, 1)
    _check_kmeans(U, w, bits, mode)


def test_kmeans_hist_path_multiplicities(U, hist_path):
    rng = np.random.RandomState(5)
    # a coarse grid: every distinct value occurs hundreds of times (entries with large counts, wide key range)
    w = (np.round(rng.randn(400 * 1000) * 40) / 1024).astype(np.float32)
    _check_kmeans(U, w, 4, "linear")
    _check_kmeans(U, w, 5, "forgy", seed=2)
    # fewer distinct values than clusters: relocation pops samples out of multi-count entries
    vals = np.sort(((0.5 + 1.5 * rng.rand(20)) * rng.choice([-1.0, 1.0], size=20)).astype(np.float32))
    w = vals[rng.randint(0, 20, size=100 * 1000)]
    _check_kmeans(U, w, 5, "linear")
    _check_kmeans(U, w, 6, "forgy", seed=1)
    # narrow positive range: the whole key fits the shared-memory histogram (no partition pass)
    w = (1.0 + rng.rand(500 * 1000) * 2.0 ** -11).astype(np.float32)
    _check_kmeans(U, w, 4, "linear")
    # ... two narrow lobes of opposite sign, zeros in between
    w = ((1.0 + rng.rand(300 * 1000) * 2.0 ** -12) * rng.choice([-1.0, 1.0, 0.0], size=300 * 1000)).astype(np.float32)
    _check_kmeans(U, w, 3, "linear")
    _check_kmeans(U, w, 8, "linear")


def test_kmeans_hist_path_4m(U, hist_path):
    # several work items per bucket region, partition pass over many tiles
    w = D.gaussian(1 << 22, seed=11)
    O.prune_weigth(w, 1)
    _check_kmeans(U, w, 8, "linear", check_ref32=False)
    _check_kmeans(U, w, 4, "linear", check_ref32=False)


def test_kmeans_hist_path_equals_radix_path(U, monkeypatch):
    # the two representations of the sorted survivors give the same fit, bit for bit, on a tensor whose largest
    # bucket spans several histogram work items (a near-constant lobe)
    rng = np.random.RandomState(9)
    g = D.gaussian(1 << 21, seed=3)
    O.prune_weigth(g, 1)  # survivors |x| > 0.02: with the lobe at 0.5 the key range stays within 27 bits
    w = np.concatenate([g, (0.5 + rng.rand(3 << 20) * 2.0 ** -9).astype(np.float32)])
    rng.shuffle(w)
    out = {}
    for path in ("radix", "hist"):
        monkeypatch.setenv("NNC_SORT_PATH", path)
        ris, km = U.get_quantized_weight(w, 6, "linear")
        out[path] = (ris.tobytes(), km.cluster_centers_.tobytes(), km.n_iter_, km.packed_codes.tobytes(), km.code_histogram.tobytes(),
                     km.inertia_, km.n_relocations)
    assert out["radix"] == out["hist"]


@pytest.mark.parametrize("env", ["NNC_LLOYD_MULTI_LAUNCH", "NNC_LLOYD_SPLIT"])
@pytest.mark.parametrize("bits,mode", [(4, "linear"), (8, "linear"), (5, "forgy")])
def test_kmeans_per_phase_launch_paths(U, monkeypatch, env, bits, mode):
    # the Lloyd loop as one launch per phase / with the update split in its three launches (the multi-rank path without a
    # peer mailbox: all-reduces between the launches, no-ops on one rank) gives the same fit as the one-launch loop
    w = D.gaussian(300 * 1000, seed=5)
    w[::9] = 0.0
    O.prune_weigth(w, 1)
    np.random.seed(3)
    ris0, km0 = U.get_quantized_weight(w, bits, mode)
    monkeypatch.setenv(env, "1")
    np.random.seed(3)
    ris1, km1 = U.get_quantized_weight(w, bits, mode)
    assert km1.n_iter_ == km0.n_iter_ and km1.n_relocations == km0.n_relocations
    assert km1.cluster_centers_.tobytes() == km0.cluster_centers_.tobytes()
    assert ris1.tobytes() == ris0.tobytes() and np.array_equal(km1.code_histogram, km0.code_histogram)
    assert km1.inertia_ == km0.inertia_


@pytest.mark.parametrize("bits,mode", [(2, "density"), (4, "linear"), (8, "linear"), (8, "density"), (5, "forgy"), (9, "linear")])
def test_kmeans_cluster_loop_equals_cooperative_loop(U, monkeypatch, bits, mode):
    # lloyd_fast.cu (one thread-block cluster, state in shared memory) against lloyd.cu's cooperative one-launch loop:
    # same fit bit for bit, relocation iterations included (8-bit linear / density on pruned data empty dozens of
    # clusters at once); 9 bits = 512 clusters is the largest codebook of the cluster path
    w = D.gaussian(300 * 1000, seed=6)
    O.prune_weigth(w, 1)
    cdfs = U.get_weight_distribution(w, skip_zeros=True) if mode == "density" else None
    np.random.seed(4)
    ris0, km0 = U.get_quantized_weight(w, bits, mode, cdfs)
    monkeypatch.setenv("NNC_LLOYD_NO_CLUSTER", "1")
    np.random.seed(4)
    ris1, km1 = U.get_quantized_weight(w, bits, mode, cdfs)
    assert km1.n_iter_ == km0.n_iter_ and km1.n_relocations == km0.n_relocations and km1.strict_convergence == km0.strict_convergence
    assert km1.cluster_centers_.tobytes() == km0.cluster_centers_.tobytes()
    assert km1.tol_ == km0.tol_
    assert ris1.tobytes() == ris0.tobytes() and np.array_equal(km1.code_histogram, km0.code_histogram)
    if bits == 8:
        assert km0.n_relocations > 0


@pytest.mark.parametrize("on_device", [False, True])
def test_compress_tensors_batched_equals_per_tensor(U, on_device):
    """all tensors of a model in one call (concurrent streams) == the per-tensor calls, bit for bit,
    and == the oracle: LeNet300-100 with 2-bit density (config 1) and LeNet5 with 4-bit linear (config 2)."""
    import torch

    for tensors, bits, mode in ((D.lenet300_tensors(), 2, "density"), (D.lenet5_tensors(), 4, "linear"), (D.lenet5_tensors(), 3, "forgy")):
        flat, qs = [], []
        for name, w, b, (qw, qb) in tensors:
            flat += [w, b]
            qs += [qw, qb]
        # reference flow per tensor through the single-tensor helpers
        np.random.seed(11)
        single = []
        forgy_idx = [np.random.randint(0, t.size, size=2 ** bits) if t.size >= 2 ** bits + 1 else None for t in flat] if mode == "forgy" else None
        for i, (t, q) in enumerate(zip(flat, qs)):
            t = t.copy()
            m = U.prune_weigth(t, q, True)
            cdfs = U.get_weight_distribution(t, skip_zeros=True) if mode == "density" else None
            if mode == "forgy":
                if forgy_idx[i] is None:
                    single.append((m, t, None))
                    continue
                space = t.ravel()[forgy_idx[i]]
                det = O.kmeans1d(t, space, mode=O.MODE_DET)
                single.append((m, det.cluster_centers_[det.labels_].reshape(t.shape), det))
                continue
            ris, km = U.get_quantized_weight(t, bits, mode, cdfs)
            single.append((m, ris, km))
        np.random.seed(11)
        batch_in = [torch.from_numpy(t.copy()).cuda() if on_device else t.copy() for t in flat]
        out = U.compress_tensors(batch_in, qs, True, bits, mode)
        assert len(out) == len(flat)
        for (m1, r1, k1), (m0, r0, k0), t_in in zip(out, single, batch_in):
            to_np = (lambda x: x.cpu().numpy()) if on_device else (lambda x: x)
            assert np.array_equal(to_np(m1), m0)
            if k0 is None:
                assert k1 is None
                continue
            assert to_np(r1).tobytes() == np.ascontiguousarray(r0).tobytes()
            assert k1.cluster_centers_.tobytes() == k0.cluster_centers_.tobytes()
            assert k1.n_iter_ == k0.n_iter_
            assert np.array_equal(to_np(k1.labels_), k0.labels_)


@pytest.mark.parametrize("on_device", [False, True])
def test_compress_model_native_batch_equals_per_tensor(U, on_device):
    """SURVEY 8f row 4, the native form: nnc_compress_many_f32 (a pool of native worker threads, one context and stream
    each) gives every tensor of LeNet300-100 (2-bit density, config 1) and LeNet5 (4-bit linear, config 2) the same mask,
    codebook, packed codes and histogram as the per-tensor helpers -- and the dequantised tensor equals the oracle's."""
    import torch

    for tensors, bits, mode in ((D.lenet300_tensors(), 2, "density"), (D.lenet5_tensors(), 4, "linear")):
        flat, qs = [], []
        for name, w, b, (qw, qb) in tensors:
            flat += [w, b]
            qs += [qw, qb]
        single = []
        for t, q in zip(flat, qs):
            t = t.copy()
            m = U.prune_weigth(t, q, True)
            if t.size < 2 ** bits + 1:
                single.append((m, t, None, None))
                continue
            cdfs = U.get_weight_distribution(t, skip_zeros=True) if mode == "density" else None
            ris, km = U.get_quantized_weight(t, bits, mode, cdfs)
            space = O.init_centroids(t, bits, mode, cdfs)
            det = O.kmeans1d(t, space, mode=O.MODE_DET)
            assert km.cluster_centers_.tobytes() == det.cluster_centers_.tobytes()
            single.append((m, t, ris, km))
        for rep in range(2):  # the second call reuses the pool and its contexts
            batch_in = [torch.from_numpy(t.copy()).cuda() if on_device else t.copy() for t in flat]
            out = U.compress_model(batch_in, qs, True, bits, mode)
            assert len(out) == len(flat)
            to_np = (lambda x: x.cpu().numpy()) if on_device else (lambda x: x)
            for (m1, k1), (m0, pruned0, ris0, k0), t_in in zip(out, single, batch_in):
                assert np.array_equal(to_np(m1), m0)
                assert to_np(t_in).tobytes() == pruned0.tobytes()  # pruned in place
                if k0 is None:
                    assert k1 is None
                    continue
                assert k1.cluster_centers_.tobytes() == k0.cluster_centers_.tobytes()
                assert k1.n_iter_ == k0.n_iter_
                assert np.array_equal(k1.code_histogram, np.bincount(k0.labels_, minlength=k1.cluster_centers_.size))
                deq = U.dequantize(k1.packed_codes, t_in.size if not on_device else t_in.numel(), k1.code_bits, k1.cluster_centers_)
                assert to_np(deq).reshape(ris0.shape).tobytes() == np.ascontiguousarray(ris0).tobytes()
    # no thresholds: an already pruned model is only quantised
    pruned = [s_[1] for s_ in single]
    out = U.compress_model([p.copy() for p in pruned], None, True, 4, "linear")
    for (m1, k1), (m0, p0, ris0, k0) in zip(out, single):
        assert m1 is None
        if k0 is not None:
            assert k1.cluster_centers_.tobytes() == k0.cluster_centers_.tobytes()


def test_kmeans_errors(U):
    w = D.gaussian(1000, seed=2)
    with pytest.raises(Exception, match="error mode not found"):
        U.get_quantized_weight(w, 2, "nope")
    with pytest.raises(Exception, match="error mode not found"):
        U.get_quantized_weight(w, 2, "density", None)
    bad = w.copy()
    bad[3] = np.nan
    with pytest.raises(ValueError):
        U.get_quantized_weight(bad, 2, "forgy")
    with pytest.raises(TypeError):
        U.get_quantized_weight(w.astype(np.float64), 2, "linear")


def test_kmeans_torch_device(U):
    import torch

    w = D.gaussian(1 << 20, seed=31)
    O.prune_weigth(w, 1)
    space, _ = _init_for(w, 4, "linear")
    det = O.kmeans1d(w, space, mode=O.MODE_DET)
    t = torch.from_numpy(w).cuda().reshape(1024, 1024)
    ris, km = U.get_quantized_weight(t, 4, "linear")
    assert ris.is_cuda and km.labels_.is_cuda and km.packed_codes.is_cuda
    assert km.cluster_centers_.tobytes() == det.cluster_centers_.tobytes()
    assert np.array_equal(km.labels_.cpu().numpy(), det.labels_)
    assert ris.cpu().numpy().tobytes() == det.cluster_centers_[det.labels_].reshape(1024, 1024).tobytes()


# ---------------------------------------------------------------------------------------------------------
# trained-quantization gradient sum
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,k,bits", [(1000, 4, 0), (100003, 16, 4), (1 << 20, 256, 8), (1 << 20, 256, 0), (77777, 33, 6), (50000, 5, 3)])
def test_grad_segsum(U, n, k, bits):
    rng = np.random.RandomState(7)
    grad = (rng.randn(n) * 1e-3).astype(np.float32)
    labels = np.random.RandomState(8).randint(0, k, size=n).astype(np.int32)
    expect = O.grad_segsum(grad, labels, k)
    assert np.allclose(expect, np.bincount(labels, weights=grad.astype(np.float64), minlength=k), rtol=1e-12, atol=0)
    codes = labels if bits == 0 else O.pack_codes(labels, bits)
    got = U.cluster_gradient_sum(grad, codes, k, bits)
    mag = np.bincount(labels, weights=np.abs(grad).astype(np.float64), minlength=k)
    assert np.all(np.abs(got - expect) <= 1e-13 * mag + 1e-300)


# ---------------------------------------------------------------------------------------------------------
# the label rule under adversarial centroid layouts (region table / zones vs brute force over all k)
# ---------------------------------------------------------------------------------------------------------
def _centroid_layouts(rng, lo, hi):
    span = hi - lo
    yield np.linspace(lo, hi, 256).astype(np.float32)
    yield np.sort(rng.uniform(lo, hi, 37)).astype(np.float32)
    c = rng.uniform(lo, hi, 64).astype(np.float32)
    c[10:20] = c[10]  # exact duplicates
    yield c
    c = rng.uniform(lo, hi, 64).astype(np.float32)
    for i in range(0, 60, 3):  # neighbours a few ulps apart
        c[i + 1] = np.nextafter(c[i], np.float32(np.inf))
        c[i + 2] = np.nextafter(c[i + 1], np.float32(np.inf))
    yield c
    c = (lo + span * rng.beta(0.3, 0.3, 500)).astype(np.float32)  # crowded at both ends, unsorted ids
    yield c
    c = (rng.randn(1000) * span * 1e-4 + (lo + hi) / 2).astype(np.float32)  # k = 1000 in a sliver
    yield c
    yield np.array([lo, hi], dtype=np.float32)
    yield np.array([(lo + hi) / 2], dtype=np.float32)
    yield (rng.standard_cauchy(300) * span * 0.01).astype(np.float32).clip(-1e3, 1e3)  # far outside the data too


@pytest.mark.parametrize("offset,sigma", [(0.0, 0.02), (0.37, 0.02), (-1000.0, 3.0), (1e-3, 1e-5)])
def test_label_rule_stress(U, offset, sigma):
    rng = np.random.RandomState(123)
    w = (rng.randn(200003) * sigma + offset).astype(np.float32)
    w[::7] = 0.0
    w[5:5000:11] = w[4]  # repeated values
    mean = np.mean(w)
    lo, hi = float(w.min() - mean), float(w.max() - mean)
    for centred in _centroid_layouts(rng, lo, hi):
        k = centred.size
        expect = O.assign(w, centred, mean)
        km = U.KMeansResult(centred.reshape(-1, 1) + mean, None, 0, 0.0, centred_centers=centred, mean=mean)
        labels, packed, hist = U.assign_codes(w, km)
        assert np.array_equal(labels, expect), (k, int((labels != expect).sum()))
        assert np.array_equal(hist, np.bincount(expect, minlength=k))
        assert np.array_equal(O.unpack_codes(packed, w.size, U.index_bits(k)), expect)


# ---------------------------------------------------------------------------------------------------------
# fused prune + quantize (compress_weight) against the two separate reference steps
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode,bits", [("linear", 8), ("linear", 4), ("density", 2), ("forgy", 5)])
def test_compress_weight_matches_separate_steps(U, mode, bits):
    w = D.gaussian(400 * 1000, seed=91).reshape(400, 1000)
    a, b = w.copy(), w.copy()
    np.random.seed(5)
    mask, km = U.compress_weight(a, 1, True, bits, mode)
    mask_ref = U.prune_weigth(b, 1, True)
    assert np.array_equal(mask, mask_ref) and a.tobytes() == b.tobytes()
    cdfs = U.get_weight_distribution(b, skip_zeros=True) if mode == "density" else None
    np.random.seed(5)
    ris, km_ref = U.get_quantized_weight(b, bits, mode, cdfs)
    assert km.cluster_centers_.tobytes() == km_ref.cluster_centers_.tobytes()
    assert km.n_iter_ == km_ref.n_iter_ and km.code_bits == km_ref.code_bits
    assert np.array_equal(km.packed_codes, km_ref.packed_codes)
    assert np.array_equal(km.code_histogram, km_ref.code_histogram)
    assert U.dequantize(km.packed_codes, w.size, km.code_bits, km.cluster_centers_).reshape(w.shape).tobytes() == ris.tobytes()
    # host array, no write-back: same compressed output, the argument keeps its original values
    c = w.copy()
    np.random.seed(5)
    mask2, km2 = U.compress_weight(c, 1, True, bits, mode, update_weights=False)
    if mode == "linear":
        assert c.tobytes() == w.tobytes()
    assert np.array_equal(mask2, mask_ref) and np.array_equal(km2.packed_codes, km_ref.packed_codes)


@pytest.mark.parametrize("n,q", [(1 << 22, 1.0), ((1 << 22) + 12345, 0.5), (3 * 1000 * 1000 + 1, 2.0), (70001, 1.0), (1 << 20, 0.0)])
def test_compress_weight_fused_prologue_large(U, n, q, monkeypatch):
    # the k-means prologue rides on the pruning pass (reduce_np.cu: VisitApplyQuant): enough elements for the speculation
    # band to be populated, ragged tile sizes, a threshold of zero (nothing pruned); against the unfused library path
    w = D.gaussian(n, seed=n % 1013)
    a, b = w.copy(), w.copy()
    mask, km = U.compress_weight(a, q, True, 6, "linear")
    monkeypatch.setenv("NNC_NO_FUSE", "1")
    mask_ref, km_ref = U.compress_weight(b, q, True, 6, "linear")
    w_np = w.copy()
    mask_np = D.prune_np(w_np, q)
    assert np.array_equal(mask, mask_np) and a.tobytes() == w_np.tobytes()
    assert np.array_equal(mask, mask_ref) and a.tobytes() == b.tobytes()
    assert km.mean == km_ref.mean == np.float32(np.mean(w_np))
    assert km.cluster_centers_.tobytes() == km_ref.cluster_centers_.tobytes()
    assert km.n_iter_ == km_ref.n_iter_ and km.n_nonzero == km_ref.n_nonzero == int(np.count_nonzero(w_np))
    assert np.array_equal(km.packed_codes, km_ref.packed_codes)
    assert np.array_equal(km.code_histogram, km_ref.code_histogram)


# ---------------------------------------------------------------------------------------------------------
# multi-GPU: sharded run == single-rank run, bit for bit (needs >= 2 GPUs; the driver's 1-GPU box skips it)
# ---------------------------------------------------------------------------------------------------------
def test_sharded_equals_single_rank():
    import os
    import subprocess
    import sys

    import torch

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if ngpu < 4 else 4
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multi_gpu_worker.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", "29517", worker], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("OK") == 5 * world and "MISMATCH" not in r.stdout


# ---------------------------------------------------------------------------------------------------------
# Trainer drivers on a device-resident torch model (trainer.py:177-206, :42-72) against the NumPy flow on the oracle
# ---------------------------------------------------------------------------------------------------------
def test_compressor_lenet300_matches_reference_flow(U):
    import torch

    from neural_network_compression_b200.common.trainer import Compressor
    from neural_network_compression_b200.neural_networks import LeNet300100

    torch.manual_seed(0)
    net = LeNet300100().cuda()
    host = {name: [p.detach().cpu().numpy().copy() for p in (layer.weight, layer.bias)] for name, layer in net.get_config().items()}
    comp = Compressor(net, net.layers_to_prune_with_threshold())
    # pruned_train step: prune, (optimizer step), re-apply
    comp._prune_parameters(True)
    thresholds = {"dense1": (1, 0.1), "dense2": (1, 0.1), "out": (0.5, 0)}
    for name, layer in net.get_config().items():
        w, b = host[name]
        mw = O.prune_weigth(w, thresholds[name][0])
        mb = O.prune_weigth(b, thresholds[name][1])
        zw, zb = comp.pruned_indexes_by_layer[layer]
        assert np.array_equal(zw.cpu().numpy(), mw) and np.array_equal(zb.cpu().numpy(), mb)
        assert layer.weight.detach().cpu().numpy().tobytes() == w.tobytes()
    with torch.no_grad():
        for layer in net.get_config().values():
            layer.weight.add_(0.001)  # an "optimizer step" that revives pruned weights
    comp._reset_pruned_parameters()
    for name, layer in net.get_config().items():
        zw, _ = comp.pruned_indexes_by_layer[layer]
        assert bool((layer.weight.detach()[zw] == 0).all())
        host[name][0] = layer.weight.detach().cpu().numpy().copy()
    # quantize: 2-bit density with the CDF of the non-zero weights (main.py:96-105)
    fitted = comp.quantize(True, 2, "density")
    for name, layer in net.get_config().items():
        for p, h, km in zip((layer.weight, layer.bias), host[name], fitted[layer]):
            nz = h.ravel()[h.ravel() != 0]
            cdfs = O.get_weight_distribution(nz)
            space = O.init_centroids(h, 2, "density", cdfs)
            det = O.kmeans1d(h, space, mode=O.MODE_DET)
            expect = det.cluster_centers_[det.labels_].reshape(h.shape)
            assert km.cluster_centers_.tobytes() == det.cluster_centers_.tobytes()
            assert p.detach().cpu().numpy().tobytes() == expect.tobytes()


def test_trained_quantization_step_matches_numpy(U):
    # Deep Compression's codebook fine-tuning (report.tex:149-153): centroid gradient = per-cluster sum of the weight
    # gradients; SGD step on the codebook; layer re-materialised from the packed indices
    import torch

    from neural_network_compression_b200.common.trainer import Compressor
    from neural_network_compression_b200.neural_networks import LeNet300100

    torch.manual_seed(1)
    net = LeNet300100().cuda()
    comp = Compressor(net, net.layers_to_prune_with_threshold())
    comp._prune_parameters(True)
    fitted = comp.quantize(False, 4, "linear")
    before = {}
    for layer, models in fitted.items():
        for p, km in zip((layer.weight, layer.bias), models):
            p.grad = torch.randn_like(p) * 1e-2
            if km is not None:
                before[p] = (km.cluster_centers_.copy(), km.labels_.cpu().numpy())
    lr = 0.05
    comp.trained_quantization_step(fitted, lr)
    checked = 0
    for layer, models in fitted.items():
        for p, km in zip((layer.weight, layer.bias), models):
            if km is None:
                continue
            centers0, labels = before[p]
            g = np.bincount(labels, weights=p.grad.cpu().numpy().ravel().astype(np.float64), minlength=km.n_clusters)
            expect = (centers0.astype(np.float64).ravel() - lr * g).astype(np.float32)
            np.testing.assert_allclose(km.cluster_centers_.ravel(), expect, rtol=1e-6, atol=1e-9)
            assert p.detach().cpu().numpy().ravel().tobytes() == km.cluster_centers_.ravel()[labels].tobytes()
            checked += 1
    assert checked >= 5

