"""ctypes binding of the C ABI in include/nnc.h (libnnc_b200.so, built in-tree by build.py).

There is no CPU fallback: if the library is missing or no CUDA device is visible, every compute call
raises.  PyTorch / NumPy are used only to own buffers; the library sees raw pointers and sizes.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnnc_b200.so")

NNC_OK = 0
NNC_ERR_BAD_ARG = 1
NNC_ERR_CUDA = 2
NNC_ERR_NOT_ENOUGH = 3
NNC_ERR_NONFINITE = 4
NNC_ERR_UNSUPPORTED = 5
NNC_ERR_INTERNAL = 6
NNC_ERR_COMM = 7
NNC_KMAX = 1024
NNC_KM_INERTIA = 1
NNC_KM_INIT_LINEAR = 2
NNC_KM_MASK_BITS = 4

# every symbol include/nnc.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "nnc_version", "nnc_last_error", "nnc_ctx_create", "nnc_ctx_destroy", "nnc_ctx_set_stream", "nnc_ctx_reserve",
    "nnc_timer_start", "nnc_timer_stop", "nnc_last_profile", "nnc_stats_f32", "nnc_prune_f32", "nnc_mask_apply_f32",
    "nnc_compact_nonzero_f32", "nnc_minmax_f32", "nnc_hist_edges_f32", "nnc_weight_cdf_f32", "nnc_gather_f32",
    "nnc_kmeans1d_f32", "nnc_assign_f32", "nnc_unpack_gather_f32", "nnc_grad_segsum_f32", "nnc_ctx_set_comm",
    "nnc_ctx_set_kernel_timing", "nnc_last_kernel_times", "nnc_ctx_total_launches", "nnc_compress_many_f32",
    "nnc_compress_f32", "nnc_shard_range", "nnc_comm_unique_id", "nnc_ctx_init_nccl",
    "nnc_peer_mailbox_create", "nnc_peer_mailbox_connect", "nnc_pack_bits_u8", "nnc_ctx_hint_global_n",
]


class NncError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("libnnc_b200 error %d: %s" % (code, msg))
        self.code = code
        self.msg = msg


class KMeansInfo(C.Structure):
    _fields_ = [
        ("n_iter", C.c_int),
        ("strict", C.c_int),
        ("n_relocations", C.c_int),
        ("fixed_exp", C.c_int),
        ("mean", C.c_float),
        ("tol", C.c_float),
        ("inertia", C.c_double),
        ("n_nonzero", C.c_int64),
    ]


class TensorJob(C.Structure):
    """nnc_tensor_job (include/nnc.h): one tensor of nnc_compress_many_f32."""
    _fields_ = [
        ("w", C.c_void_p),
        ("n", C.c_int64),
        ("threshold", C.c_double),
        ("prune", C.c_int),
        ("pad_", C.c_int),
        ("mask", C.c_void_p),
        ("centers", C.c_void_p),
        ("centred", C.c_void_p),
        ("packed", C.c_void_p),
        ("hist", C.c_void_p),
        ("info", KMeansInfo),
        ("thr", C.c_double),
        ("n_pruned", C.c_int64),
        ("k", C.c_int),
        ("code_bits", C.c_int),
        ("status", C.c_int),
        ("error", C.c_char * 196),
    ]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p)


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(lib().nnc_comm_unique_id(buf))
    return buf.raw


def shard_range(n: int, rank: int, world: int):
    """[begin, end) of the flattened n-element tensor that `rank` of `world` owns (reduction-tile aligned)."""
    b, e = C.c_int64(), C.c_int64()
    check(lib().nnc_shard_range(int(n), int(rank), int(world), C.byref(b), C.byref(e)))
    return b.value, e.value


class _DevView:
    """A raw device pointer dressed as __cuda_array_interface__ so that torch can wrap it without a copy."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def torch_allreduce(group=None):
    """An all-reduce callback for Context.set_comm built on torch.distributed (NCCL on GPU boxes)."""
    import torch
    import torch.distributed as dist

    ops = {0: dist.ReduceOp.SUM, 1: dist.ReduceOp.MIN, 2: dist.ReduceOp.MAX}

    def allreduce(ptr, count, op, stream):
        with torch.cuda.stream(torch.cuda.ExternalStream(stream)):
            t = torch.as_tensor(_DevView(ptr, count), device="cuda")
            dist.all_reduce(t, op=ops[op], group=group)

    return allreduce

_lib = None
_lib_lock = threading.Lock()


def lib():
    """Loads libnnc_b200.so.  Fails loudly when it has not been built (python -m neural_network_compression_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libnnc_b200.so is missing at %s: build it with `python -m neural_network_compression_b200.build` "
                "(needs nvcc; there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, i64, i32, f32, f64 = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_double
        P = C.POINTER
        L.nnc_version.restype = i32
        L.nnc_last_error.restype = C.c_char_p
        L.nnc_ctx_create.argtypes = [i32, P(vp)]
        L.nnc_ctx_destroy.argtypes = [vp]
        L.nnc_ctx_destroy.restype = None
        L.nnc_ctx_set_stream.argtypes = [vp, vp]
        L.nnc_ctx_reserve.argtypes = [vp, C.c_size_t]
        L.nnc_timer_start.argtypes = [vp]
        L.nnc_timer_stop.argtypes = [vp, P(f32)]
        L.nnc_last_profile.argtypes = [vp, P(f32), i32, P(i32), P(C.c_char_p), P(i64)]
        L.nnc_stats_f32.argtypes = [vp, vp, i64, P(f32), P(f32), P(f32)]
        L.nnc_prune_f32.argtypes = [vp, vp, i64, f64, i32, i32, vp, P(f64), P(i64)]
        L.nnc_mask_apply_f32.argtypes = [vp, vp, vp, i64]
        L.nnc_compact_nonzero_f32.argtypes = [vp, vp, i64, vp, P(i64)]
        L.nnc_minmax_f32.argtypes = [vp, vp, i64, i32, P(f32), P(f32), P(i64)]
        L.nnc_hist_edges_f32.argtypes = [vp, vp, i64, vp, i32, i32, vp]
        L.nnc_weight_cdf_f32.argtypes = [vp, vp, i64, i32, vp, vp]
        L.nnc_gather_f32.argtypes = [vp, vp, i64, vp, i32, vp]
        L.nnc_kmeans1d_f32.argtypes = [vp, vp, i64, vp, i32, i32, f64, i32, vp, vp, vp, vp, vp, i32, vp, P(KMeansInfo)]
        L.nnc_compress_f32.argtypes = [vp, vp, i64, f64, i32, i32, i32, vp, P(f64), P(i64), vp, i32, i32, f64, i32, vp, vp, vp, i32,
                                       vp, P(KMeansInfo)]
        L.nnc_assign_f32.argtypes = [vp, vp, i64, vp, i32, f32, vp, vp, vp, vp, i32, vp, P(f64)]
        L.nnc_unpack_gather_f32.argtypes = [vp, vp, i64, i32, vp, i32, vp]
        L.nnc_grad_segsum_f32.argtypes = [vp, vp, vp, i64, i32, i32, vp]
        L.nnc_pack_bits_u8.argtypes = [vp, vp, i64, vp]
        L.nnc_compress_many_f32.argtypes = [vp, P(TensorJob), i32, i32, i32, i32, i32, i32]
        L.nnc_ctx_hint_global_n.argtypes = [vp, i64]
        L.nnc_ctx_set_kernel_timing.argtypes = [vp, i32, C.c_char_p]
        L.nnc_last_kernel_times.argtypes = [vp, P(C.c_char_p)]
        L.nnc_ctx_total_launches.argtypes = [vp, P(i64)]
        L.nnc_shard_range.argtypes = [i64, i32, i32, P(i64), P(i64)]
        L.nnc_peer_mailbox_create.argtypes = [vp, i32, C.c_char_p]
        L.nnc_peer_mailbox_connect.argtypes = [vp, C.c_char_p, i32, i32]
        L.nnc_comm_unique_id.argtypes = [C.c_char_p]
        L.nnc_ctx_init_nccl.argtypes = [vp, C.c_char_p, i32, i32]
        L.nnc_ctx_set_comm.argtypes = [vp, i32, i32, ALLREDUCE_FN, vp]
        for name in EXPORTS:
            fn = getattr(L, name)
            if name not in ("nnc_last_error", "nnc_ctx_destroy"):
                fn.restype = i32
        _lib = L
    return _lib


def check(rc: int):
    if rc != NNC_OK:
        raise NncError(rc, lib().nnc_last_error().decode("utf-8", "replace"))


class Context:
    """One nnc_ctx: a device, a stream and a workspace.  Not thread-safe; one per host thread / rank."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        check(lib().nnc_ctx_create(int(device), C.byref(self._h)))
        self.device = int(device)
        self._comm_cb = None
        self.peer_exchange = False  # True once the in-kernel peer-memory exchange is connected
        self.rank, self.world = 0, 1  # multi-GPU: set by init_nccl / set_comm
        self.dist_group = None        # the torch.distributed group the ranks were taken from (utility.init_distributed)

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            lib().nnc_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int | None):
        check(lib().nnc_ctx_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def reserve(self, nbytes: int):
        check(lib().nnc_ctx_reserve(self._h, int(nbytes)))

    def timer_start(self):
        check(lib().nnc_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        check(lib().nnc_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def last_profile(self):
        """({phase: ms}, kernel launches) of the last prune / k-means call."""
        buf = (C.c_float * 64)()
        n = C.c_int()
        names = C.c_char_p()
        launches = C.c_int64()
        check(lib().nnc_last_profile(self._h, buf, 64, C.byref(n), C.byref(names), C.byref(launches)))
        keys = names.value.decode().split(";") if names.value else []
        out = {}
        for i, kname in enumerate(keys[: n.value]):
            out[kname] = out.get(kname, 0.0) + buf[i]
        return out, launches.value

    def set_kernel_timing(self, on: bool, name_filter: str | None = None):
        check(lib().nnc_ctx_set_kernel_timing(self._h, int(bool(on)), name_filter.encode() if name_filter else None))

    def total_launches(self) -> int:
        v = C.c_int64()
        check(lib().nnc_ctx_total_launches(self._h, C.byref(v)))
        return v.value

    def last_kernel_times(self):
        """{kernel name: (launches, total ms)} accumulated since set_kernel_timing(True)."""
        s = C.c_char_p()
        check(lib().nnc_last_kernel_times(self._h, C.byref(s)))
        out = {}
        for item in (s.value.decode() if s.value else "").split(";"):
            if item:
                name, cnt, ms = item.rsplit(":", 2)
                out[name] = (int(cnt), float(ms))
        return out

    def init_nccl(self, unique_id: bytes, rank: int, world: int):
        """The library's own NCCL communicator (all-reduces enqueued natively on the context's stream)."""
        assert len(unique_id) == 128
        check(lib().nnc_ctx_init_nccl(self._h, unique_id, int(rank), int(world)))
        self.rank, self.world = int(rank), int(world)

    def hint_global_size(self, n_global: int):
        """Sharded calls: the element count of the whole tensor (same on every rank; 0 = ask the ranks in every call)."""
        check(lib().nnc_ctx_hint_global_n(self._h, int(n_global)))

    def peer_mailbox_create(self, world: int) -> bytes:
        buf = C.create_string_buffer(64)
        check(lib().nnc_peer_mailbox_create(self._h, int(world), buf))
        return buf.raw

    def peer_mailbox_connect(self, handles: bytes | None, rank: int, world: int):
        """handles: the world x 64 bytes of all ranks' mailboxes in rank order, or None to disconnect."""
        assert handles is None or len(handles) == 64 * world
        check(lib().nnc_peer_mailbox_connect(self._h, handles, int(rank), int(world)))
        self.peer_exchange = handles is not None

    def set_comm(self, rank: int, world: int, allreduce):
        """allreduce(dev_ptr: int, count: int, op: int, stream: int) -> None sums/mins/maxes int64 in place."""
        if world > 1:
            def _cb(_user, buf, count, op, stream):
                try:
                    allreduce(buf, count, op, stream)
                    return 0
                except Exception:  # pragma: no cover - surfaced as NNC_ERR_COMM
                    import traceback

                    traceback.print_exc()
                    return 1

            self._comm_cb = ALLREDUCE_FN(_cb)
        else:
            self._comm_cb = C.cast(None, ALLREDUCE_FN)
        check(lib().nnc_ctx_set_comm(self._h, rank, world, self._comm_cb, None))
        self.rank, self.world = int(rank), int(world)


_tls = threading.local()


def default_context(device: int | None = None) -> Context:
    """A per-thread, per-device context, created on first use."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if "NNC_DEVICE" not in os.environ else int(os.environ["NNC_DEVICE"])
    cache = getattr(_tls, "ctx", None)
    if cache is None:
        cache = _tls.ctx = {}
    if device not in cache:
        cache[device] = Context(device)
        with _all_lock:
            _all_contexts.append(cache[device])
    return cache[device]


_all_contexts: list = []  # every default context of the process (the batched calls run on worker threads with their own)
_all_lock = threading.Lock()


def total_launches_all() -> int:
    """Kernel launches of all default contexts of this process (main thread and batch workers)."""
    with _all_lock:
        ctxs = list(_all_contexts)
    return sum(c.total_launches() for c in ctxs)


# ---- buffer plumbing -------------------------------------------------------------------------------------
def is_torch(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def ptr(x) -> int:
    """Raw address of a NumPy array or torch tensor (host or device)."""
    if x is None:
        return 0
    if is_torch(x):
        return x.data_ptr()
    return x.ctypes.data


def device_of(x) -> int | None:
    """CUDA device index of a torch CUDA tensor, else None."""
    if is_torch(x) and x.is_cuda:
        return x.device.index
    return None


def np_buffer(shape, dtype):
    return np.empty(shape, dtype=dtype)
