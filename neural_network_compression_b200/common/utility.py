"""B200-native drop-in for the compression helpers of the reference's
neural_network_compression/common/utility.py: same names, arguments, return values and error behaviour,
computed by the sm_100a kernels behind the C ABI of include/nnc.h.

    prune_weigth(original_weigth, threshold=0.25, std_smooth=True)        utility.py:134-163
    get_weight_distribution(weight_matrix)                                 utility.py:334-392
    get_quantized_weight(layer_weight, bits=4, mode="linear", cdfs=None)   utility.py:172-240

Inputs may be float32 NumPy arrays (what the reference's callers pass, trainer.py:185-191, :52-69) or float32
torch tensors on the host or on a CUDA device; outputs come back as the same kind (device tensors stay on the
device).  There is no CPU implementation in this package: without the CUDA library the calls raise.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any, Optional

import numpy as np

from .. import _native as N

__all__ = [
    "prune_weigth", "apply_mask", "get_weight_distribution", "get_quantized_weight", "KMeansResult",
    "compress_weight", "init_distributed", "shard_range", "nonzero_weights", "weight_stats", "assign_codes", "dequantize", "cluster_gradient_sum", "index_bits",
    "pack_mask_bits", "compress_tensors", "sharded_weight_distribution", "sharded_forgy_init",
]


# ---------------------------------------------------------------------------------------------------------
# buffer plumbing
# ---------------------------------------------------------------------------------------------------------
class _Buf:
    """A flat, contiguous float32 view of a caller array plus what is needed to hand results back."""

    def __init__(self, x, name: str, writable: bool = False):
        self.orig = x
        self.copied_back = None
        if N.is_torch(x):
            import torch

            if x.dtype != torch.float32:
                raise TypeError("%s must be float32 (got %s)" % (name, x.dtype))
            self.kind = "torch"
            self.shape = tuple(x.shape)
            self.device = N.device_of(x)
            self.arr = x if x.is_contiguous() else x.contiguous()
            if writable and self.arr is not x:
                self.copied_back = lambda: x.copy_(self.arr)
            self.n = x.numel()
        else:
            if not isinstance(x, np.ndarray):
                raise TypeError("%s must be a numpy.ndarray or a torch.Tensor (got %s)" % (name, type(x).__name__))
            if x.dtype != np.float32:
                raise TypeError("%s must be float32 (got %s); the reference's callers pass Keras float32 weights"
                                % (name, x.dtype))
            self.kind = "numpy"
            self.shape = x.shape
            self.device = None
            if x.flags.c_contiguous and (x.flags.writeable or not writable):
                self.arr = x
            else:
                if writable and not x.flags.writeable:
                    raise ValueError("assignment destination is read-only")
                self.arr = np.ascontiguousarray(x)
                if writable:
                    self.copied_back = lambda: np.copyto(x, self.arr)
            self.n = x.size

    @property
    def ptr(self) -> int:
        return N.ptr(self.arr)

    def finish(self):
        if self.copied_back is not None:
            self.copied_back()

    def empty(self, n, dtype):
        """A new flat output buffer living where the input lives."""
        if self.kind == "torch":
            import torch

            tdt = {np.uint8: torch.uint8, np.int32: torch.int32, np.float32: torch.float32, np.bool_: torch.bool}[dtype]
            return torch.empty(int(n), dtype=tdt, device=self.arr.device)
        return np.empty(int(n), dtype=dtype)


def init_distributed(group=None, device=None, peer_exchange=True):
    """Multi-GPU: one process per GPU.  Installs a torch.distributed all-reduce on this process's context; from then
    on prune_weigth / get_quantized_weight / compress_weight take THIS RANK'S slice (see `shard_range`) of the
    flattened tensor and return per-slice masks / codes with global thresholds, centroids and histograms."""
    import torch.distributed as dist

    ctx = N.default_context(device)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ctx.dist_group = group
    if world == 1:
        return ctx
    # preferred: the library's own NCCL communicator, bootstrapped through the existing process group
    try:
        box = [N.nccl_unique_id() if rank == 0 else None]
    except N.NncError:
        box = [None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    if box[0] is not None:
        ctx.init_nccl(box[0], rank, world)
    else:  # no loadable libnccl: all-reduce through torch.distributed (slower: one host callback per exchange)
        ctx.set_comm(rank, world, N.torch_allreduce(group))
    # one box: peer mailboxes, so that the per-iteration exchanges of the Lloyd loop run inside its update kernel
    if peer_exchange and world <= 16:
        try:
            mine = ctx.peer_mailbox_create(world)
        except N.NncError:
            mine = None
        handles = [None] * world
        dist.all_gather_object(handles, mine, group=group)
        ok = False
        if all(h is not None for h in handles):
            try:
                ctx.peer_mailbox_connect(b"".join(handles), rank, world)
                ok = True
            except N.NncError:
                ok = False
        oks = [None] * world
        dist.all_gather_object(oks, ok, group=group)
        if not all(oks):  # all or nothing: a half-connected box would dead-lock the in-kernel exchange
            ctx.peer_mailbox_connect(None, rank, world)
    return ctx


def shard_range(n: int, rank: int, world: int):
    """[begin, end) of the flattened n-element tensor owned by `rank` (aligned to NumPy's summation tree tiles)."""
    return N.shard_range(n, rank, world)


def _ctx_for(buf: _Buf) -> N.Context:
    ctx = N.default_context(buf.device)
    if buf.kind == "torch" and buf.device is not None:
        import torch

        ctx.set_stream(torch.cuda.current_stream(buf.device).cuda_stream)
    else:
        ctx.set_stream(None)
    return ctx


# ---------------------------------------------------------------------------------------------------------
# pruning
# ---------------------------------------------------------------------------------------------------------
def prune_weigth(original_weigth, threshold=0.25, std_smooth=True):
    """Std-scaled magnitude pruning, restating utility.py:134-163.

    thr = np.std(w) * threshold when std_smooth (population std, NumPy float32 pairwise arithmetic, bit exact);
    mask = |w| < thr (strict); w[mask] = 0 IN PLACE; returns the boolean mask, same shape as the input.
    """
    buf = _Buf(original_weigth, "original_weigth", writable=True)
    ctx = _ctx_for(buf)
    # NEP 50: a Python float/int or np.float32 threshold keeps the product and the comparison in float32;
    # a np.float64 scalar promotes both to float64.
    thr_mode = 1 if isinstance(threshold, np.float64) else 0
    mask = buf.empty(buf.n, np.uint8)
    thr_out = C.c_double()
    n_pruned = C.c_int64()
    N.check(N.lib().nnc_prune_f32(ctx.handle, buf.ptr, buf.n, float(threshold), int(bool(std_smooth)), thr_mode,
                                  N.ptr(mask), C.byref(thr_out), C.byref(n_pruned)))
    buf.finish()
    prune_weigth.last_threshold = thr_out.value
    prune_weigth.last_pruned = n_pruned.value
    if buf.kind == "torch":
        import torch

        return mask.view(torch.bool).reshape(buf.shape)
    return mask.view(np.bool_).reshape(buf.shape)


prune_weigth.last_threshold = None
prune_weigth.last_pruned = None


def apply_mask(weights, mask):
    """weights[mask] = 0 in place: Trainer._reset_pruned_parameters (trainer.py:195-206); also the masked
    gradient apply."""
    buf = _Buf(weights, "weights", writable=True)
    ctx = _ctx_for(buf)
    if N.is_torch(mask):
        import torch

        m = mask.contiguous().view(torch.uint8) if mask.dtype == torch.bool else mask.contiguous()
        if m.numel() != buf.n:
            raise IndexError("boolean index did not match indexed array")
    else:
        m = np.ascontiguousarray(mask)
        if m.size != buf.n:
            raise IndexError("boolean index did not match indexed array")
        m = m.view(np.uint8) if m.dtype == np.bool_ else m.astype(np.uint8)
    N.check(N.lib().nnc_mask_apply_f32(ctx.handle, buf.ptr, N.ptr(m), buf.n))
    buf.finish()
    return weights


def pack_mask_bits(mask):
    """The pruning mask as bits (bit i of byte i // 8 = mask.flat[i]): the 1-bit mask of the compressed-layer format
    (common/storage.py).  Host bool / uint8 arrays and device tensors; the result lives where the input lives."""
    if N.is_torch(mask):
        import torch

        m = mask.contiguous().view(torch.uint8) if mask.dtype == torch.bool else mask.contiguous()
        n = m.numel()
        out = torch.empty((n + 7) // 8, dtype=torch.uint8, device=m.device)
        ctx = N.default_context(N.device_of(m))
        if m.is_cuda:
            ctx.set_stream(torch.cuda.current_stream(m.device).cuda_stream)
    else:
        m = np.ascontiguousarray(mask)
        m = m.view(np.uint8) if m.dtype == np.bool_ else m.astype(np.uint8)
        n = m.size
        out = np.empty((n + 7) // 8, dtype=np.uint8)
        ctx = N.default_context(None)
        ctx.set_stream(None)
    N.check(N.lib().nnc_pack_bits_u8(ctx.handle, N.ptr(m), n, N.ptr(out)))
    return out


def weight_stats(w):
    """(mean, var, std) as np.mean / np.var / np.std give them for a float32 array."""
    buf = _Buf(w, "w")
    ctx = _ctx_for(buf)
    m, v, s = C.c_float(), C.c_float(), C.c_float()
    N.check(N.lib().nnc_stats_f32(ctx.handle, buf.ptr, buf.n, C.byref(m), C.byref(v), C.byref(s)))
    return np.float32(m.value), np.float32(v.value), np.float32(s.value)


# ---------------------------------------------------------------------------------------------------------
# weight distribution
# ---------------------------------------------------------------------------------------------------------
def nonzero_weights(params):
    """flat[flat != 0], order preserving: the survivor selection of Trainer.quantize (trainer.py:55-59)."""
    buf = _Buf(params, "params")
    ctx = _ctx_for(buf)
    out = buf.empty(buf.n, np.float32)
    cnt = C.c_int64()
    N.check(N.lib().nnc_compact_nonzero_f32(ctx.handle, buf.ptr, buf.n, N.ptr(out), C.byref(cnt)))
    return out[: cnt.value]


# ---- the same on a tensor sharded over the ranks (min / max all-reduce, 31-bin int64 all-reduce) ----------
def sharded_weight_distribution(local_minmax, local_hist, allreduce):
    """get_weight_distribution of a tensor whose slices live on several ranks; every rank gets the same result, equal to
    the single-rank one.  The exchange logic is separated from the device calls so that it runs under any backend:
        local_minmax() -> (min, max, count) of this rank's selected elements (count 0: no element)
        local_hist(edges float32[32]) -> int64[31] counts of this rank's elements in the half-open bins
        allreduce(np.ndarray, op) -> np.ndarray, op in "min" / "max" / "sum", elementwise over the ranks.
    Restates utility.py:359-390 on the global counts (NumPy / SciPy do the 31-value CDF and its interpolation exactly
    as in the reference)."""
    from scipy.interpolate import interp1d

    mn, mx, cnt = local_minmax()
    big = np.float32(np.inf)
    lo = allreduce(np.array([mn if cnt else big], dtype=np.float32), "min")[0]
    hi = allreduce(np.array([mx if cnt else -big], dtype=np.float32), "max")[0]
    total = int(allreduce(np.array([cnt], dtype=np.int64), "sum")[0])
    if total == 0:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    steps = np.linspace(np.float32(lo), np.float32(hi), num=32)  # utility.py:365 (float32 under NumPy 2)
    counts = allreduce(np.asarray(local_hist(steps), dtype=np.int64), "sum")
    with np.errstate(invalid="ignore", divide="ignore"):
        p = counts / counts.sum()          # :375
        cdf = np.cumsum(p)                 # :377-385, a running float64 sum
        cdf = cdf / cdf[-1]
    x = steps[:-1]
    xnew = np.linspace(x[0], x[30], num=300)  # :387
    return xnew, interp1d(x, cdf)(xnew)       # :388-390


def _torch_allreduce(ctx, device):
    """allreduce(np array, op) over the ranks of the context's process group (NCCL needs device tensors)."""
    import torch
    import torch.distributed as dist

    ops = {"min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX, "sum": dist.ReduceOp.SUM}
    on_gpu = dist.get_backend(ctx.dist_group) == "nccl"

    def allreduce(a, op):
        t = torch.from_numpy(np.ascontiguousarray(a).copy())
        if on_gpu:
            t = t.cuda(device if device is not None else ctx.device)
        dist.all_reduce(t, op=ops[op], group=ctx.dist_group)
        return t.cpu().numpy()

    return allreduce


def sharded_forgy_init(idx_global, begin, end, local_gather, allreduce):
    """np.random.choice(flat, size=k) on a sharded tensor (utility.py:224-226): every rank draws the same GLOBAL indices
    from the global legacy RNG, the owner of an index supplies the value, the bit patterns are summed over the ranks
    (one non-zero contribution each)."""
    idx_global = np.asarray(idx_global, dtype=np.int64)
    mine = (idx_global >= begin) & (idx_global < end)
    bits = np.zeros(idx_global.size, dtype=np.int64)
    if mine.any():
        vals = np.asarray(local_gather(idx_global[mine] - begin), dtype=np.float32)
        bits[mine] = vals.view(np.uint32).astype(np.int64)
    return allreduce(bits, "sum").astype(np.uint32).view(np.float32)


def get_weight_distribution(weight_matrix, skip_zeros: bool = False):
    """Restates utility.py:334-392: 31 half-open bins over linspace(min, max, 32), normalised cumulative sum,
    linear interpolation onto 300 points.  Returns (xnew float32[300], cdf float64[300]).

    skip_zeros=True fuses the caller's survivor selection (trainer.py:55-60): the distribution is taken over
    the non-zero entries without materialising them.
    """
    buf = _Buf(weight_matrix, "weight_matrix")
    if buf.n == 0:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    ctx = _ctx_for(buf)
    if ctx.world > 1:  # this rank's slice of a sharded tensor: global min / max and global 31-bin counts
        def local_minmax():
            mn, mx, cnt = C.c_float(), C.c_float(), C.c_int64()
            N.check(N.lib().nnc_minmax_f32(ctx.handle, buf.ptr, buf.n, int(bool(skip_zeros)), C.byref(mn), C.byref(mx), C.byref(cnt)))
            return np.float32(mn.value), np.float32(mx.value), cnt.value

        def local_hist(edges):
            edges = np.ascontiguousarray(edges, dtype=np.float32)
            counts = np.empty(edges.size - 1, dtype=np.int64)
            N.check(N.lib().nnc_hist_edges_f32(ctx.handle, buf.ptr, buf.n, N.ptr(edges), edges.size, int(bool(skip_zeros)), N.ptr(counts)))
            return counts

        return sharded_weight_distribution(local_minmax, local_hist, _torch_allreduce(ctx, buf.device))
    xnew = np.empty(300, dtype=np.float32)
    cdf = np.empty(300, dtype=np.float64)
    try:
        N.check(N.lib().nnc_weight_cdf_f32(ctx.handle, buf.ptr, buf.n, int(bool(skip_zeros)), N.ptr(xnew), N.ptr(cdf)))
    except N.NncError as e:
        if e.code == N.NNC_ERR_NOT_ENOUGH:
            raise ValueError("zero-size array to reduction operation minimum which has no identity") from None
        raise
    return xnew, cdf


# ---------------------------------------------------------------------------------------------------------
# quantization
# ---------------------------------------------------------------------------------------------------------
def index_bits(n_clusters: int) -> int:
    """Bits per packed cluster index.  Density init yields 2^bits + 1 centroids, hence bits + 1 here."""
    return max(1, int(n_clusters - 1).bit_length())


@dataclass
class KMeansResult:
    """What the reference's callers read from the fitted sklearn model (utility.py:239), plus the compressed
    representation the reference never materialises (packed n-bit indices, code histogram)."""

    cluster_centers_: Any  # (k, 1) float32 ndarray
    labels_: Any  # (n,) int32, array kind of the input
    n_iter_: int
    inertia_: float
    packed_codes: Any = None  # uint8 little-endian bit stream, `code_bits` per weight
    code_bits: int = 0
    code_histogram: Optional[np.ndarray] = None  # (k,) int64
    centred_centers: Optional[np.ndarray] = None  # (k,) float32: centres in sklearn's mean-centred space
    mean: np.float32 = np.float32(0)
    strict_convergence: bool = False
    n_relocations: int = 0
    n_nonzero: int = 0
    tol_: float = 0.0
    profile: dict = field(default_factory=dict)

    @property
    def n_clusters(self) -> int:
        return int(self.cluster_centers_.shape[0])


def _init_density(bits: int, cdfs):
    # utility.py:210-223, verbatim semantics: for each of the 2^bits + 1 targets take the FIRST cdf value
    # closest to it (Python min), then the x of the first cdf entry equal to that value (np.argmax).
    tmp = np.linspace(0, 1, num=(2 ** bits) + 1)
    xval, yval = cdfs[0], np.asarray(cdfs[1])
    space = []
    for t in tmp:
        j = int(np.argmin(np.abs(yval - t)))  # first minimum, like min(key=abs(x - t))
        idx_val = int(np.argmax(yval == yval[j]))
        space.append(xval[idx_val])
    return np.array(space, dtype=np.float32)


def _kmeans_device(buf: _Buf, ctx: N.Context, space, want_labels=True, want_ris=True, want_packed=True,
                   max_iter: int = 300, tol: float = 1e-4, want_inertia=True, linear_k: int = 0):
    """space: initial centroids, or None with linear_k = k for np.linspace(min, max, k) computed by the library."""
    flags = (N.NNC_KM_INERTIA if want_inertia else 0) | (N.NNC_KM_INIT_LINEAR if space is None else 0)
    k = int(linear_k if space is None else space.size)
    bits = index_bits(k)
    centers = np.empty(k, dtype=np.float32)
    centred = np.empty(k, dtype=np.float32)
    hist = np.empty(k, dtype=np.int64)
    labels = buf.empty(buf.n, np.int32) if want_labels else None
    ris = buf.empty(buf.n, np.float32) if want_ris else None
    packed = buf.empty((buf.n * bits + 7) // 8, np.uint8) if want_packed else None
    info = N.KMeansInfo()
    if space is not None:
        space = np.ascontiguousarray(space, dtype=np.float32)
    N.check(N.lib().nnc_kmeans1d_f32(ctx.handle, buf.ptr, buf.n, N.ptr(space), k, int(max_iter), float(tol), flags,
                                     N.ptr(centers), N.ptr(centred), N.ptr(labels), N.ptr(ris), N.ptr(packed), bits,
                                     N.ptr(hist), C.byref(info)))
    prof, launches = ctx.last_profile()
    prof["launches"] = launches
    res = KMeansResult(
        cluster_centers_=centers.reshape(-1, 1), labels_=labels, n_iter_=info.n_iter, inertia_=info.inertia,
        packed_codes=packed, code_bits=bits, code_histogram=hist, centred_centers=centred, mean=np.float32(info.mean),
        strict_convergence=bool(info.strict), n_relocations=info.n_relocations, n_nonzero=info.n_nonzero,
        tol_=float(info.tol), profile=prof)
    return ris, res


def get_quantized_weight(layer_weight, bits=4, mode="linear", cdfs=None):
    """1-D k-means weight sharing, restating utility.py:172-240.

    Returns (ris, kmeans): `ris` has the input's shape with every weight replaced by its centroid
    (cluster_centers_[labels_]); `kmeans` carries cluster_centers_ (k, 1), labels_ (n,), n_iter_, inertia_ like
    the sklearn model the reference returns, plus packed n-bit codes and their histogram.  Fewer elements than
    2^bits + 1: prints the reference's message and returns (layer_weight, None).  Unknown mode, or "density"
    without cdfs: Exception(" error mode not found").
    """
    n_elem = int(np.prod(tuple(layer_weight.shape)))
    if n_elem < (2 ** bits) + 1:
        print("not enough bits:", n_elem, " vs ", 2 ** bits)
        return layer_weight, None

    if mode == "linear":
        buf = _Buf(layer_weight, "layer_weight")
        ctx = _ctx_for(buf)
        space = None  # np.linspace(min, max, 2**bits): the library derives it from its own min/max sweep
    elif mode == "density" and cdfs is not None:
        buf = _Buf(layer_weight, "layer_weight")
        ctx = _ctx_for(buf)
        space = _init_density(bits, cdfs)
    elif mode == "forgy":
        buf = _Buf(layer_weight, "layer_weight")
        ctx = _ctx_for(buf)
        space = _forgy_space(buf, ctx, bits)
    elif mode == "kmeans++":
        # Outside the hot path: the reference's unseeded sklearn default.  The
        # seeding runs on the host through scikit-learn; the Lloyd iterations run on the device like every
        # other mode.
        from sklearn.cluster import kmeans_plusplus

        buf = _Buf(layer_weight, "layer_weight")
        ctx = _ctx_for(buf)
        host = buf.arr.detach().cpu().numpy() if buf.kind == "torch" else buf.arr
        centers, _ = kmeans_plusplus(host.reshape(-1, 1), n_clusters=2 ** bits)
        space = centers.astype(np.float32).ravel()
    else:
        raise Exception(" error mode not found")

    try:
        ris, res = _kmeans_device(buf, ctx, None if space is None else np.asarray(space, dtype=np.float32), linear_k=2 ** bits)
    except N.NncError as e:
        if e.code == N.NNC_ERR_NONFINITE:
            raise ValueError("Input X contains NaN or infinity.") from None
        if e.code == N.NNC_ERR_NOT_ENOUGH:
            raise ValueError(e.msg) from None
        raise
    return ris.reshape(buf.shape), res


def _forgy_space(buf: _Buf, ctx: N.Context, bits: int):
    """np.random.choice(flat, size=2**bits) (utility.py:224-226): draws randint(0, n, k) from the global legacy RNG and
    indexes with it; on a sharded tensor n is the GLOBAL element count and the owners supply the values."""
    def gather(idx):
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        out = np.empty(idx.size, dtype=np.float32)
        N.check(N.lib().nnc_gather_f32(ctx.handle, buf.ptr, buf.n, N.ptr(idx), idx.size, N.ptr(out)))
        return out

    if ctx.world <= 1:
        return gather(np.random.randint(0, buf.n, size=2 ** bits))
    allreduce = _torch_allreduce(ctx, buf.device)
    n_global = int(allreduce(np.array([buf.n], dtype=np.int64), "sum")[0])
    begin, end = N.shard_range(n_global, ctx.rank, ctx.world)
    if end - begin != buf.n:
        raise ValueError("rank %d holds %d elements, shard_range assigns [%d, %d)" % (ctx.rank, buf.n, begin, end))
    return sharded_forgy_init(np.random.randint(0, n_global, size=2 ** bits), begin, end, gather, allreduce)


def _init_space(buf: _Buf, ctx: N.Context, bits: int, mode: str, cdfs):
    """Initial centroids for the deterministic / seeded modes (utility.py:206-226)."""
    if mode == "linear":
        mn, mx, cnt = C.c_float(), C.c_float(), C.c_int64()
        N.check(N.lib().nnc_minmax_f32(ctx.handle, buf.ptr, buf.n, 0, C.byref(mn), C.byref(mx), C.byref(cnt)))
        if cnt.value != buf.n:
            raise ValueError("Input X contains NaN.")
        lo, hi = np.float32(mn.value), np.float32(mx.value)
        if ctx.world > 1:
            allreduce = _torch_allreduce(ctx, buf.device)
            lo, hi = allreduce(np.array([lo], np.float32), "min")[0], allreduce(np.array([hi], np.float32), "max")[0]
        return np.linspace(lo, hi, num=2 ** bits)
    if mode == "density" and cdfs is not None:
        return _init_density(bits, cdfs)
    if mode == "forgy":
        return _forgy_space(buf, ctx, bits)
    raise Exception(" error mode not found")


def compress_weight(original_weigth, threshold=0.25, std_smooth=True, bits=4, mode="linear", with_cdf=True,
                    update_weights=True, out_mask=None, out_packed=None, mask_bits=False):
    """The whole compression of one tensor as Trainer._prune_parameters + Trainer.quantize apply it
    (trainer.py:177-193, :42-72): std-threshold prune, then k-means weight sharing of the pruned tensor, keeping
    only the compressed representation -- the boolean mask, the codebook, the packed n-bit cluster indices and
    their histogram -- instead of the dense de-quantised tensor and int32 labels.

    update_weights=True keeps prune_weigth's contract (the argument is pruned IN PLACE).  With a host array and
    update_weights=False the pruned weights are not copied back (the caller is about to replace them with the
    codebook values anyway): the tensor crosses the bus once.  out_mask / out_packed: optional preallocated
    (e.g. pinned) uint8 output buffers for host arrays.

    mask_bits=True (linear init only): the mask comes back bit-packed -- ceil(n / 8) uint8, bit i of byte i // 8, the
    layout of common/storage.py -- instead of one bool per weight: a compressed layer then leaves the device as
    n * (bits + 1) / 8 bytes.

    Returns (mask, KMeansResult); `KMeansResult.labels_` is None (decode with `dequantize`)."""
    buf = _Buf(original_weigth, "original_weigth", writable=True)
    if buf.n < (2 ** bits) + 1:
        mask = prune_weigth(original_weigth, threshold, std_smooth)
        print("not enough bits:", buf.n, " vs ", 2 ** bits)
        return mask, None
    ctx = _ctx_for(buf)
    if mode != "linear":
        if mask_bits:
            raise ValueError("mask_bits=True needs mode='linear' (the fused call); pack other masks with pack_mask_bits")
        # density / forgy initialise from the PRUNED tensor: prune first, then the reference's init, then k-means
        mask = prune_weigth(original_weigth, threshold, std_smooth)
        prune_prof, _ = ctx.last_profile()
        cdfs = get_weight_distribution(original_weigth, skip_zeros=True) if (mode == "density" and with_cdf) else None
        buf = _Buf(original_weigth, "original_weigth")
        space = np.asarray(_init_space(buf, ctx, bits, mode, cdfs), dtype=np.float32)
        try:
            _, res = _kmeans_device(buf, ctx, space, want_labels=False, want_ris=False, want_packed=True, want_inertia=False)
        except N.NncError as e:
            if e.code == N.NNC_ERR_NONFINITE:
                raise ValueError("Input X contains NaN or infinity.") from None
            raise
        for name, ms in prune_prof.items():
            res.profile["prune." + name] = ms
        return mask, res
    # linear init needs only min / max of the pruned tensor, which the k-means prologue measures: one fused call
    k = 2 ** bits
    cbits = index_bits(k)
    thr_mode = 1 if isinstance(threshold, np.float64) else 0
    mask = out_mask if out_mask is not None else buf.empty((buf.n + 7) // 8 if mask_bits else buf.n, np.uint8)
    packed = out_packed if out_packed is not None else buf.empty((buf.n * cbits + 7) // 8, np.uint8)
    centers = np.empty(k, dtype=np.float32)
    centred = np.empty(k, dtype=np.float32)
    hist = np.empty(k, dtype=np.int64)
    info = N.KMeansInfo()
    thr_out, n_pruned = C.c_double(), C.c_int64()
    try:
        N.check(N.lib().nnc_compress_f32(ctx.handle, buf.ptr, buf.n, float(threshold), int(bool(std_smooth)), thr_mode,
                                         int(bool(update_weights)), N.ptr(mask), C.byref(thr_out), C.byref(n_pruned), None, k,
                                         300, 1e-4, N.NNC_KM_INIT_LINEAR | (N.NNC_KM_MASK_BITS if mask_bits else 0), N.ptr(centers), N.ptr(centred), N.ptr(packed), cbits,
                                         N.ptr(hist), C.byref(info)))
    except N.NncError as e:
        if e.code == N.NNC_ERR_NONFINITE:
            raise ValueError("Input X contains NaN or infinity.") from None
        raise
    if update_weights:
        buf.finish()
    prune_weigth.last_threshold = thr_out.value
    prune_weigth.last_pruned = n_pruned.value
    prof, launches = ctx.last_profile()
    prof["launches"] = launches
    res = KMeansResult(
        cluster_centers_=centers.reshape(-1, 1), labels_=None, n_iter_=info.n_iter, inertia_=info.inertia,
        packed_codes=packed, code_bits=cbits, code_histogram=hist, centred_centers=centred, mean=np.float32(info.mean),
        strict_convergence=bool(info.strict), n_relocations=info.n_relocations, n_nonzero=info.n_nonzero,
        tol_=float(info.tol), profile=prof)
    if mask_bits:
        return mask, res
    if buf.kind == "torch":
        import torch

        mask = mask.view(torch.bool) if mask.dtype == torch.uint8 else mask
        return mask.reshape(buf.shape), res
    return mask.view(np.bool_).reshape(buf.shape), res


# ---------------------------------------------------------------------------------------------------------
# batched many-small-tensor mode
# ---------------------------------------------------------------------------------------------------------
_pool = None
_pool_streams = {}


def _worker_stream(device):
    """One CUDA stream per (worker thread, device), created on first use."""
    import threading

    import torch

    key = (threading.get_ident(), device)
    s = _pool_streams.get(key)
    if s is None:
        s = _pool_streams[key] = torch.cuda.Stream(device=device)
    return s


def compress_tensors(tensors, thresholds=None, std_smooth=True, bits=4, mode="linear", with_cdf=True, workers=8):
    """All tensors of a model in ONE call: what the reference's loops do tensor by tensor (prune every kernel / bias,
    trainer.py:177-193; then quantize every array of every layer, trainer.py:50-70), with the tensors processed
    CONCURRENTLY -- every tensor's kernels run on a stream of their own (own context, own workspace), driven by a small
    pool of host threads, so that launch latencies and the few host round trips of the per-tensor pipeline overlap
    instead of adding up.  A LeNet has 6-8 tensors of 10 .. 627 200 weights: the call costs about as much as its largest
    tensor.  Results are bit-identical to the per-tensor calls (each tensor runs exactly the same kernels).

    tensors: float32 ndarrays / torch tensors (device tensors are pruned and read in place, as by prune_weigth).
    thresholds: one pruning quality parameter per tensor, or None to skip pruning (already pruned model).
    Returns [(mask or None, ris, kmeans or None)] in input order; a tensor with fewer than 2**bits + 1 elements comes back
    unquantised with kmeans None, like get_quantized_weight (utility.py:202-204).
    """
    global _pool
    from concurrent.futures import ThreadPoolExecutor

    tensors = list(tensors)
    if thresholds is not None and len(thresholds) != len(tensors):
        raise ValueError("one threshold per tensor")
    if mode == "forgy":
        # np.random.choice draws from the global legacy RNG in layer order (utility.py:224-226): draw here, in order
        forgy_idx = [np.random.randint(0, int(np.prod(tuple(t.shape))), size=2 ** bits).astype(np.int64)
                     if int(np.prod(tuple(t.shape))) >= (2 ** bits) + 1 else None for t in tensors]
    else:
        forgy_idx = [None] * len(tensors)
    # one single-thread executor per worker slot; tensor i always goes to the same slot (largest tensors first, round
    # robin), so that every slot's context keeps seeing the same sizes: its workspace arena reaches its high-water mark
    # in the first call and is never regrown afterwards (a regrow is a cudaFree + cudaMalloc: a device-wide stall)
    n_slots = max(1, min(int(workers), len(tensors)))
    if _pool is None or len(_pool) < n_slots:
        _pool = (_pool or []) + [ThreadPoolExecutor(max_workers=1, thread_name_prefix="nnc-batch") for _ in range(n_slots - len(_pool or []))]
    order = sorted(range(len(tensors)), key=lambda i: -int(np.prod(tuple(tensors[i].shape))))
    slot_of = {i: pos % n_slots for pos, i in enumerate(order)}
    producer = {}
    if any(N.is_torch(t) and t.is_cuda for t in tensors):
        import torch

        for t in tensors:
            if N.is_torch(t) and t.is_cuda and t.device.index not in producer:
                producer[t.device.index] = torch.cuda.current_stream(t.device)

    def one(i):
        t = tensors[i]
        on_dev = N.is_torch(t) and t.is_cuda

        def work():
            mask = None
            if thresholds is not None:
                mask = prune_weigth(t, thresholds[i], std_smooth)
            n_elem = int(np.prod(tuple(t.shape)))
            if n_elem < (2 ** bits) + 1:
                print("not enough bits:", n_elem, " vs ", 2 ** bits)
                return mask, t, None
            if mode == "forgy":
                buf = _Buf(t, "layer_weight")
                ctx = _ctx_for(buf)
                space = np.empty(forgy_idx[i].size, dtype=np.float32)
                N.check(N.lib().nnc_gather_f32(ctx.handle, buf.ptr, buf.n, N.ptr(forgy_idx[i]), forgy_idx[i].size, N.ptr(space)))
                ris, res = _kmeans_device(buf, ctx, space)
                return mask, ris.reshape(buf.shape), res
            cdfs = None
            if mode == "density" and with_cdf:
                try:
                    cdfs = get_weight_distribution(t, skip_zeros=True)
                except ValueError:  # an all-zero tensor has no survivors: the reference fails there as well
                    raise
            ris, res = get_quantized_weight(t, bits, mode, cdfs)
            return mask, ris, res

        if on_dev:
            import torch

            s = _worker_stream(t.device.index)
            s.wait_stream(producer[t.device.index])
            with torch.cuda.stream(s):
                out = work()
            return out, s
        return work(), None

    futures = {i: _pool[slot_of[i]].submit(one, i) for i in order}
    results = []
    for f in (futures[i] for i in range(len(tensors))):
        out, s = f.result()
        if s is not None:
            prod = producer[s.device.index]
            prod.wait_stream(s)
            mask, ris, res = out
            for x in (mask, ris, getattr(res, "labels_", None), getattr(res, "packed_codes", None)):
                if N.is_torch(x) and x.is_cuda:
                    x.record_stream(prod)  # allocated on the worker's stream, used on the caller's from here on
        results.append(out)
    return results


def assign_codes(weights, kmeans: KMeansResult, want_labels=True, want_packed=True):
    """E-step only: labels / packed codes of `weights` against a fitted codebook (the sklearn label rule)."""
    buf = _Buf(weights, "weights")
    ctx = _ctx_for(buf)
    k = kmeans.n_clusters
    bits = kmeans.code_bits or index_bits(k)
    labels = buf.empty(buf.n, np.int32) if want_labels else None
    packed = buf.empty((buf.n * bits + 7) // 8, np.uint8) if want_packed else None
    hist = np.empty(k, dtype=np.int64)
    centred = np.ascontiguousarray(kmeans.centred_centers, dtype=np.float32)
    N.check(N.lib().nnc_assign_f32(ctx.handle, buf.ptr, buf.n, N.ptr(centred), k, float(kmeans.mean), None,
                                   N.ptr(labels), None, N.ptr(packed), bits, N.ptr(hist), None))
    return labels, packed, hist


def compress_model(tensors, thresholds=None, std_smooth=True, bits=4, mode="linear", workers=8):
    """All tensors of a model compressed in ONE native call (nnc_compress_many_f32): what the trainer does layer by layer
    and tensor by tensor (prune every kernel / bias, trainer.py:177-193; quantise every array, trainer.py:50-70;
    le_net_5.py:17-34), keeping the compressed form of every tensor -- mask, codebook, packed indices, histogram -- like
    compress_weight.  The tensors run concurrently on a pool of native worker threads with a context and stream each; no
    interpreter between a tensor's ~25 launches.  Bit-identical to compress_weight per tensor.

    tensors: float32 ndarrays / torch tensors (pruned in place).  thresholds: one quality parameter per tensor, or None for
    an already pruned model.  mode: "linear" or "density" (2**bits + 1 centroids, as the reference has it).
    Returns [(mask or None, KMeansResult or None)] in input order; a tensor with fewer than 2**bits + 1 elements comes back
    pruned and unquantised (None), like get_quantized_weight (utility.py:202-204)."""
    tensors = list(tensors)
    if thresholds is not None and len(thresholds) != len(tensors):
        raise ValueError("one threshold per tensor")
    if mode not in ("linear", "density"):
        raise Exception(" error mode not found")
    if not tensors:
        return []
    bufs = [_Buf(t, "tensors[%d]" % i, writable=thresholds is not None) for i, t in enumerate(tensors)]
    devices = {b.device for b in bufs if b.device is not None}
    if len(devices) > 1:
        raise ValueError("compress_model: all device tensors must live on one GPU")
    # NEP-50: a np.float64 threshold promotes the comparison to float64, a Python float does not; one mode per call
    modes = {isinstance(q, np.float64) for q in thresholds} if thresholds is not None else {False}
    if len(modes) > 1:
        raise ValueError("compress_model: thresholds must be all Python floats or all numpy.float64")
    thr_mode = int(modes.pop())
    ctx = _ctx_for(next((b for b in bufs if b.device is not None), bufs[0]))
    k_lin = 2 ** bits
    k_max = k_lin + (1 if mode == "density" else 0)
    cbits = index_bits(k_max)
    jobs = (N.TensorJob * len(bufs))()
    keep = []
    for i, b in enumerate(bufs):
        q = thresholds[i] if thresholds is not None else 0.0
        mask = b.empty(b.n, np.uint8) if thresholds is not None else None
        packed = b.empty((b.n * cbits + 7) // 8, np.uint8) if b.n >= k_lin + 1 else None
        centers, centred, hist = np.empty(k_max, np.float32), np.empty(k_max, np.float32), np.empty(k_max, np.int64)
        keep.append((mask, packed, centers, centred, hist))
        j = jobs[i]
        j.w, j.n, j.threshold, j.prune = b.ptr, b.n, float(q), int(thresholds is not None)
        j.mask, j.packed = N.ptr(mask), N.ptr(packed)
        j.centers, j.centred, j.hist = N.ptr(centers), N.ptr(centred), N.ptr(hist)
    rc = N.lib().nnc_compress_many_f32(ctx.handle, jobs, len(bufs), int(bool(std_smooth)), thr_mode, int(bits),
                                       0 if mode == "linear" else 1, int(workers))
    if rc == N.NNC_ERR_NONFINITE:
        raise ValueError("Input X contains NaN or infinity.")
    N.check(rc)
    out = []
    for i, b in enumerate(bufs):
        b.finish()
        mask, packed, centers, centred, hist = keep[i]
        j = jobs[i]
        m = None
        if mask is not None:
            m = (mask.view(np.bool_) if isinstance(mask, np.ndarray) else mask.bool()).reshape(tuple(b.shape))
        if j.k == 0:
            print("not enough bits:", b.n, " vs ", k_lin)
            out.append((m, None))
            continue
        k = int(j.k)
        info = j.info
        out.append((m, KMeansResult(
            cluster_centers_=centers[:k].reshape(-1, 1), labels_=None, n_iter_=info.n_iter, inertia_=info.inertia,
            packed_codes=packed, code_bits=int(j.code_bits), code_histogram=hist[:k], centred_centers=centred[:k],
            mean=np.float32(info.mean), strict_convergence=bool(info.strict), n_relocations=info.n_relocations,
            n_nonzero=info.n_nonzero, tol_=float(info.tol), profile={})))
    return out


def dequantize(packed_codes, n: int, bits: int, cluster_centers, like=None):
    """cluster_centers_[codes] from packed n-bit codes (the decode side of utility.py:239)."""
    values = np.ascontiguousarray(np.asarray(cluster_centers, dtype=np.float32).ravel())
    if N.is_torch(packed_codes):
        import torch

        out = torch.empty(int(n), dtype=torch.float32, device=packed_codes.device)
        dev = N.device_of(packed_codes)
    else:
        packed_codes = np.ascontiguousarray(packed_codes, dtype=np.uint8)
        out = np.empty(int(n), dtype=np.float32)
        dev = None
    ctx = N.default_context(dev)
    if dev is not None:  # run on the stream the codes were produced on (the context keeps the stream of its previous call)
        import torch

        ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    else:
        ctx.set_stream(None)
    N.check(N.lib().nnc_unpack_gather_f32(ctx.handle, N.ptr(packed_codes), int(n), int(bits), N.ptr(values), values.size,
                                          N.ptr(out)))
    return out


def cluster_gradient_sum(grad, codes, n_clusters: int, bits: int = 0):
    """Trained-quantization centroid gradient (papers/lat/report.tex:152): out[k] = sum_i grad_i [code_i == k].
    `codes`: int32 labels (bits=0) or a packed n-bit stream.  Returns float64[n_clusters]."""
    buf = _Buf(grad, "grad")
    ctx = _ctx_for(buf)
    if not N.is_torch(codes):
        codes = np.ascontiguousarray(codes, dtype=np.int32 if bits == 0 else np.uint8)
    elif not codes.is_contiguous():
        codes = codes.contiguous()
    out = np.empty(int(n_clusters), dtype=np.float64)
    N.check(N.lib().nnc_grad_segsum_f32(ctx.handle, buf.ptr, N.ptr(codes), buf.n, int(bits), int(n_clusters), N.ptr(out)))
    return out
