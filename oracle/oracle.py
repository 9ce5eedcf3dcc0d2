"""ctypes front-end of oracle/nnc_oracle.c.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs as the checker / CPU baseline.  The product package
(neural_network_compression_b200) never imports this module.

Mirrors the reference helper signatures (utility.py:134-240, 334-392) on top of the C
restatement so that parity tests read like calls into the reference.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libnnc_oracle.so")

MODE_REF32 = 0
MODE_DET = 1


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "nnc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


class _Info(C.Structure):
    _fields_ = [
        ("n_iter", C.c_int),
        ("strict", C.c_int),
        ("n_relocations", C.c_int),
        ("fixed_exp", C.c_int),
        ("mean", C.c_float),
        ("tol", C.c_float),
        ("inertia", C.c_double),
    ]


_FAR_CB = C.CFUNCTYPE(None, C.POINTER(C.c_float), C.c_int64, C.c_int, C.POINTER(C.c_int64))
_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i64, f32p = C.c_void_p, C.c_int64, C.POINTER(C.c_float)
        L.nnco_set_threads.argtypes = [C.c_int]
        L.nnco_pairwise_sum_f32.restype = C.c_float
        L.nnco_pairwise_sum_f32.argtypes = [vp, i64]
        L.nnco_std_f32.argtypes = [vp, i64, f32p, f32p, f32p]
        L.nnco_prune_f32.restype = i64
        L.nnco_prune_f32.argtypes = [vp, i64, C.c_double, C.c_int, C.c_int, vp, C.POINTER(C.c_double)]
        L.nnco_mask_apply_f32.argtypes = [vp, vp, i64]
        L.nnco_compact_nonzero_f32.restype = i64
        L.nnco_compact_nonzero_f32.argtypes = [vp, i64, vp]
        L.nnco_linspace_f32.argtypes = [C.c_float, C.c_float, C.c_int, vp]
        L.nnco_weight_cdf_f32.restype = C.c_int
        L.nnco_weight_cdf_f32.argtypes = [vp, i64, vp, vp, vp]
        L.nnco_init_linear_f32.argtypes = [vp, i64, C.c_int, vp]
        L.nnco_init_density_f32.restype = C.c_int
        L.nnco_init_density_f32.argtypes = [vp, vp, C.c_int, vp]
        L.nnco_kmeans1d_f32.restype = C.c_int
        L.nnco_kmeans1d_f32.argtypes = [vp, i64, vp, C.c_int, C.c_int, C.c_double, C.c_int, _FAR_CB, vp, vp, vp,
                                        C.POINTER(_Info)]
        L.nnco_assign_f32.argtypes = [vp, i64, vp, C.c_int, C.c_float, vp]
        L.nnco_gather_f32.argtypes = [vp, vp, i64, vp]
        L.nnco_pack_codes.argtypes = [vp, i64, C.c_int, vp]
        L.nnco_unpack_codes.argtypes = [vp, i64, C.c_int, vp]
        L.nnco_code_histogram.argtypes = [vp, i64, C.c_int, vp]
        L.nnco_grad_segsum_f64.argtypes = [vp, vp, i64, C.c_int, vp]
        L.nnco_grad_segsum_fixed.restype = C.c_int
        L.nnco_grad_segsum_fixed.argtypes = [vp, vp, i64, C.c_int, vp]
        _lib = L
    return _lib


def set_threads(t: int):
    """Threads for the DET-mode loops (result independent of the count); REF32 is always sequential."""
    lib().nnco_set_threads(int(t))


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def _p(a):
    return a.ctypes.data


# ---------------------------------------------------------------------------------------------
def pairwise_sum(a) -> np.float32:
    a = _f32(a).ravel()
    return np.float32(lib().nnco_pairwise_sum_f32(_p(a), a.size))


def std(w):
    """(mean, var, std) exactly as np.mean/np.var/np.std compute them for float32."""
    w = _f32(w).ravel()
    m, v, s = C.c_float(), C.c_float(), C.c_float()
    lib().nnco_std_f32(_p(w), w.size, C.byref(m), C.byref(v), C.byref(s))
    return np.float32(m.value), np.float32(v.value), np.float32(s.value)


def prune_weigth(original_weigth, threshold=0.25, std_smooth=True):
    """Restatement of utility.py:134-163.  Mutates its argument, returns the bool mask."""
    w = original_weigth
    assert isinstance(w, np.ndarray) and w.dtype == np.float32 and w.flags.c_contiguous
    mode = 1 if isinstance(threshold, np.float64) else 0
    mask = np.empty(w.shape, dtype=np.uint8)
    thr = C.c_double()
    lib().nnco_prune_f32(_p(w), w.size, float(threshold), int(bool(std_smooth)), mode, _p(mask), C.byref(thr))
    prune_weigth.last_threshold = thr.value
    return mask.view(np.bool_)


def mask_apply(w, mask):
    assert w.dtype == np.float32 and w.flags.c_contiguous
    m = np.ascontiguousarray(mask).view(np.uint8)
    lib().nnco_mask_apply_f32(_p(w), _p(m), w.size)


def compact_nonzero(w):
    w = _f32(w).ravel()
    out = np.empty(w.size, dtype=np.float32)
    c = lib().nnco_compact_nonzero_f32(_p(w), w.size, _p(out))
    return out[:c].copy()


def linspace_f32(start, stop, num):
    out = np.empty(num, dtype=np.float32)
    lib().nnco_linspace_f32(np.float32(start), np.float32(stop), num, _p(out))
    return out


def get_weight_distribution(weight_matrix, return_counts=False):
    """Restatement of utility.py:334-392."""
    w = _f32(weight_matrix).ravel()
    xnew = np.empty(300, dtype=np.float32)
    cdf = np.empty(300, dtype=np.float64)
    cnt = np.empty(31, dtype=np.int64)
    rc = lib().nnco_weight_cdf_f32(_p(w), w.size, _p(xnew), _p(cdf), _p(cnt))
    if rc != 0:
        raise ValueError("empty input")
    return (xnew, cdf, cnt) if return_counts else (xnew, cdf)


def init_centroids(layer_weight, bits, mode, cdfs=None, forgy_indices=None):
    """Restatement of utility.py:206-226.  forgy needs the indices np.random.randint drew."""
    w = _f32(layer_weight).ravel()
    if mode == "linear":
        space = np.empty(2 ** bits, dtype=np.float32)
        lib().nnco_init_linear_f32(_p(w), w.size, bits, _p(space))
        return space
    if mode == "density" and cdfs is not None:
        xnew = _f32(cdfs[0])
        cdf = np.ascontiguousarray(cdfs[1], dtype=np.float64)
        space = np.empty(2 ** bits + 1, dtype=np.float32)
        lib().nnco_init_density_f32(_p(xnew), _p(cdf), bits, _p(space))
        return space
    if mode == "forgy":
        return w[np.asarray(forgy_indices)]
    raise Exception(" error mode not found")


@dataclass
class KMeansOracleResult:
    cluster_centers_: np.ndarray  # (k, 1) float32
    labels_: np.ndarray  # (n,) int32
    centred_centers: np.ndarray  # (k,) float32, c' = centre - mean (internal space of sklearn)
    n_iter_: int
    inertia_: float
    strict: bool
    n_relocations: int
    fixed_exp: int
    mean: np.float32
    tol: np.float32


def _numpy_far_cb(dist_p, n, n_empty, out_p):
    # skl: _k_means_common.pyx:187 -- the order of np.argpartition's tail is implementation defined,
    # so REF32 pinning asks NumPy itself.
    dist = np.ctypeslib.as_array(dist_p, shape=(n,))
    far = np.argpartition(dist, -n_empty)[:-n_empty - 1:-1]
    out = np.ctypeslib.as_array(out_p, shape=(n_empty,))
    out[:] = far


_NUMPY_FAR = _FAR_CB(_numpy_far_cb)
_NULL_FAR = C.cast(None, _FAR_CB)


def kmeans1d(w, init, max_iter=300, tol=1e-4, mode=MODE_REF32, numpy_far_order=None) -> KMeansOracleResult:
    """Restatement of KMeans(n_clusters=k, init=init, n_init=1, algorithm='lloyd').fit(w.reshape(-1,1))."""
    w = _f32(w).ravel()
    init = _f32(init).ravel()
    k = init.size
    if numpy_far_order is None:
        numpy_far_order = mode == MODE_REF32
    centers = np.empty(k, dtype=np.float32)
    centred = np.empty(k, dtype=np.float32)
    labels = np.empty(w.size, dtype=np.int32)
    info = _Info()
    rc = lib().nnco_kmeans1d_f32(_p(w), w.size, _p(init), k, max_iter, tol, mode,
                                 _NUMPY_FAR if numpy_far_order else _NULL_FAR,
                                 _p(centers), _p(centred), _p(labels), C.byref(info))
    if rc != 0:
        raise ValueError("n_samples=%d should be >= n_clusters=%d" % (w.size, k))
    return KMeansOracleResult(centers.reshape(-1, 1), labels, centred, info.n_iter, info.inertia, bool(info.strict),
                              info.n_relocations, info.fixed_exp, np.float32(info.mean), np.float32(info.tol))


def assign(w, centred_centers, mean):
    w = _f32(w).ravel()
    c = _f32(centred_centers).ravel()
    labels = np.empty(w.size, dtype=np.int32)
    lib().nnco_assign_f32(_p(w), w.size, _p(c), c.size, np.float32(mean), _p(labels))
    return labels


def get_quantized_weight(layer_weight, bits=4, mode="linear", cdfs=None, kmeans_mode=MODE_REF32, forgy_indices=None):
    """Restatement of utility.py:172-240 (modes linear / density / forgy)."""
    if np.prod(layer_weight.shape) < (2 ** bits) + 1:
        return layer_weight, None
    if mode == "forgy" and forgy_indices is None:
        forgy_indices = np.random.randint(0, layer_weight.size, size=2 ** bits)
    space = init_centroids(layer_weight, bits, mode, cdfs, forgy_indices)
    km = kmeans1d(layer_weight, space, mode=kmeans_mode)
    ris = km.cluster_centers_[km.labels_].reshape(layer_weight.shape)
    return ris, km


def pack_codes(labels, bits):
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    out = np.empty((labels.size * bits + 7) // 8, dtype=np.uint8)
    lib().nnco_pack_codes(_p(labels), labels.size, bits, _p(out))
    return out


def unpack_codes(packed, n, bits):
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    labels = np.empty(n, dtype=np.int32)
    lib().nnco_unpack_codes(_p(packed), n, bits, _p(labels))
    return labels


def code_histogram(labels, k):
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    h = np.empty(k, dtype=np.int64)
    lib().nnco_code_histogram(_p(labels), labels.size, k, _p(h))
    return h


def grad_segsum(grad, labels, k, fixed=False):
    grad = _f32(grad).ravel()
    labels = np.ascontiguousarray(labels, dtype=np.int32).ravel()
    out = np.empty(k, dtype=np.float64)
    if fixed:
        lib().nnco_grad_segsum_fixed(_p(grad), _p(labels), grad.size, k, _p(out))
    else:
        lib().nnco_grad_segsum_f64(_p(grad), _p(labels), grad.size, k, _p(out))
    return out
