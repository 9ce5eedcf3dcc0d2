#!/usr/bin/env python
"""bench.py -- weights/sec of prune + k-means weight sharing on B200 (BASELINE.json metric).

One "step" = the whole compression of one synthetic layer:
    std-threshold magnitude prune (q = 1.0, in place, bool mask out)            utility.py:134-163
    + 8-bit linear-init 1-D k-means of the pruned tensor (zeros included)        utility.py:172-240
    emitting the codebook, packed 8-bit cluster indices and their histogram.
Workload (configs[3] of BASELINE.json): N(0, 0.02^2) float32, 2^30 weights in total ("bimodal post-prune":
about 68 % zeros, survivors in two lobes |w| > sigma), contiguous shards over the ranks.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 the CUDA path (one JSON line)
    python bench.py --impl reference ...                                the CPU restatement, all host threads

Timing: CUDA events on the library's stream, barrier + synchronize on both sides, max over ranks.  Every step
reads a fresh 4 GiB tensor (larger than the 126 MB L2), so no L2 flush is needed between steps.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line

import numpy as np  # noqa: E402

METRIC = "weights/sec prune+k-means"
UNIT = "weights/s"
SIGMA = 0.02
SEED = 2024
QUALITY = 1.0
BITS = 8
MODE = "linear"


def env_int(name, default):
    return int(os.environ.get(name, default))


# ---------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  NVML in a background thread (two light
    queries every 10 ms); an `nvidia-smi -lms` child process was measured to stall kernel launches for milliseconds
    per query on some boxes, which showed up as idle gaps between the kernels of a step.  Falls back to nvidia-smi
    when NVML is not importable."""
    QUERY = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None
        self.thread = None
        self.samples = []  # (wall time, sm MHz, reasons bit mask)
        self.max_mhz = None
        self._stop = False

    def _nvml_loop(self, nv, handle):
        while not self._stop:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(handle)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                self.samples.append((time.time(), float(mhz), int(reasons)))
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if os.environ.get("NNC_BENCH_NO_CLOCKS"):
            return
        try:
            import threading

            import pynvml as nv

            nv.nvmlInit()
            # NVML enumerates all GPUs of the box; CUDA_VISIBLE_DEVICES may remap the CUDA ordinal
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    idx = self.gpu
            handle = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
            self._nv = nv
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t_begin=None, t_end=None):
        """Median SM clock over the samples taken between the two wall-clock times (the timed region)."""
        import datetime

        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            self._stop = True
            self.thread.join(timeout=2)
            nv = self._nv
            names = (("hw_slowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                     ("hw_thermal_slowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                     ("sw_thermal_slowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                     ("sw_power_cap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)))
            sel = [s for s in self.samples if t_begin is None or (t_begin - 0.02 <= s[0] <= t_end + 0.02)]
            reasons = set()
            for _, _, mask in sel:
                for nm, bit in names:
                    if mask & bit:
                        reasons.add(nm)
            if sel:
                out["sm_mhz"] = statistics.median(s[1] for s in sel)
                out["samples"] = len(sel)
            out["sm_max_mhz"] = self.max_mhz
            out["reasons"] = sorted(reasons)
            out["source"] = "nvml"
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons = [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    if t_begin is not None:
                        ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                        if ts < t_begin - 0.05 or ts > t_end + 0.05:
                            continue
                    sm.append(float(f[1]))
                    out["sm_max_mhz"] = float(f[2])
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        out["source"] = "nvidia-smi"
        return out


# ---------------------------------------------------------------------------------------------------------
# CPU arms.  Preferred: the library calls the reference's helpers make (NumPy for prune_weigth, utility.py:159-162;
# scikit-learn's KMeans with the reference's arguments, utility.py:237-239) -- "reference".  Without an importable
# scikit-learn: the C restatement under oracle/ (test infrastructure; used here only as the timed baseline) -- "port".
# ---------------------------------------------------------------------------------------------------------
CPU_THREADS_CAP = 16  # the same thread count in every run of a round (BENCH and SCALE boxes differ in core count)


def cpu_threads():
    return max(1, min(CPU_THREADS_CAP, os.cpu_count() or 1))


def cpu_tensor(n_sample):
    return (np.random.RandomState(SEED).randn(n_sample) * SIGMA).astype(np.float32)


def cpu_pipeline_sklearn(w, threads, bits=BITS, quality=QUALITY):
    """The reference's own sequence of library calls on `w` (pruned in place).  Returns (seconds, n_iter)."""
    import warnings

    from sklearn.cluster import KMeans
    from threadpoolctl import threadpool_limits

    with threadpool_limits(limits=threads), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter()
        thr = np.std(w) * quality                      # utility.py:159
        mask = np.abs(w) < thr                          # utility.py:161
        w[mask] = 0                                     # utility.py:162
        space = np.linspace(np.min(w), np.max(w), num=2 ** bits)  # utility.py:206-209
        km = KMeans(n_clusters=len(space), init=space.reshape(-1, 1), n_init=1, algorithm="lloyd").fit(w.reshape(-1, 1))  # :237-238
        ris = km.cluster_centers_[km.labels_].reshape(w.shape)  # utility.py:239
        dt = time.perf_counter() - t0
    del ris
    return dt, int(km.n_iter_)


def cpu_pipeline_port(w, threads):
    from oracle import oracle as O

    O.set_threads(threads)
    t0 = time.perf_counter()
    O.prune_weigth(w, QUALITY)
    space = O.init_centroids(w, BITS, MODE)
    km = O.kmeans1d(w, space, mode=O.MODE_DET if threads > 1 else O.MODE_REF32)
    O.pack_codes(km.labels_, BITS)
    return time.perf_counter() - t0, km.n_iter_


def cpu_pipeline(n_sample, threads):
    """prune + 8-bit linear k-means to convergence of an n_sample-weight tensor of the workload's distribution on the
    host.  Returns (seconds, n_iter, kind, engine)."""
    w = cpu_tensor(n_sample)
    try:
        import sklearn

        dt, it = cpu_pipeline_sklearn(w, threads)
        return dt, it, "reference", "numpy %s + scikit-learn %s KMeans(init=linspace, n_init=1, algorithm='lloyd'): the calls of utility.py:159-162, 206-209, 237-239" % (np.__version__, sklearn.__version__)
    except ImportError:
        dt, it = cpu_pipeline_port(w, threads)
        return dt, it, "port", "oracle/nnc_oracle.c (C restatement, pthreads)"


def cpu_sample_text(n_sample, n_iter, seconds):
    return ("a 2^%d-weight tensor of the workload's distribution (NOT the 2^30 layer: the reference needs ~2 min per Lloyd iteration "
            "there), same pipeline to convergence: %d Lloyd iterations, %.1f s" % (int(np.log2(n_sample)), n_iter, seconds))


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    threads = cpu_threads()
    n_sample = args.cpu_sample
    times = []
    n_iter, kind, engine = 0, "port", ""
    for i in range(args.warmup + args.steps):
        dt, n_iter, kind, engine = cpu_pipeline(n_sample, threads)
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = n_sample * len(times) / total
    cfg = workload_config(args, args.gpus)
    # this arm times a SAMPLE of the workload: say so where the driver compares configs
    cfg["n_weights_timed"] = n_sample
    cfg["sampled"] = True
    cfg["workload"] += " -- CPU arm: timed on a 2^%d-weight sample of the same distribution" % int(np.log2(n_sample))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "engine": engine,
                         "sample": cpu_sample_text(n_sample, n_iter, total / len(times))},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, world):
    return {
        "workload": "synthetic 2^%d-weight fp32 layer N(0,0.02^2), std-threshold prune q=1.0 (bimodal post-prune), "
                    "8-bit linear-init k-means to convergence, packed 8-bit codes + histogram out" % int(np.log2(args.n)),
        "n_weights": args.n, "bits": BITS, "init": MODE, "quality": QUALITY, "shards": world,
        "exchange": "none (one rank)" if world == 1 else "per-iteration (count, sum) all-reduce inside the Lloyd update kernel over "
                    "NVLink peer memory; NCCL int64 all-reduces for the reduction-tree partials and scalars",
        "l2": "every step reads a fresh %.1f GiB tensor (> 126 MB L2); no flush needed" % (args.n * 4 / world / 2 ** 30),
    }


# ---------------------------------------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------------------------------------
def run_b200(args):
    # libraries (NCCL's version banner, ...) may write to stdout: rank 0's JSON line goes to the real stdout, the
    # rest of the run sees stderr there
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch

    from neural_network_compression_b200 import _native as N
    from neural_network_compression_b200.common import utility as U

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = N.default_context(local)
    if world > 1:
        U.init_distributed(device=local)
    if world > 1:
        ctx.hint_global_size(args.n)  # every rank knows the layer's size: no all-reduce + host round trip per call to agree on it
    lo, hi = U.shard_range(args.n, rank, world)  # contiguous slice of the flattened layer owned by this rank
    n_local = hi - lo
    K, W = args.steps, args.warmup

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- inputs: one fresh tensor per step (pruning is in place), generated on the device before timing
    pool = min(K + W, args.pool)
    gen = torch.Generator(device=dev)
    bufs = [torch.empty(n_local, dtype=torch.float32, device=dev) for _ in range(pool)]
    # The GLOBAL tensor of a step is the same for every rank count: it is generated in fixed chunks of 2^22 weights, each
    # from its own Philox seed (step, chunk); a rank fills the part of its slice that every chunk covers.  Results
    # (threshold, centroids, n_iter, code histogram, code checksum -> `result_hash`) must then agree across N = 1/2/4/8.
    CH = 1 << 22
    chunk_tmp = torch.empty(CH, dtype=torch.float32, device=dev)
    fills = [0]

    def fill(buf, tensor_index):
        for c in range(lo // CH, (hi - 1) // CH + 1):
            gen.manual_seed((SEED * 1000003 + tensor_index) * 4099 + c)
            chunk_tmp.normal_(0.0, SIGMA, generator=gen)
            g0, g1 = max(lo, c * CH), min(hi, (c + 1) * CH)
            buf[g0 - lo:g1 - lo].copy_(chunk_tmp[g0 - c * CH:g1 - c * CH])

    def refill(count):
        for b in bufs[:count]:
            fill(b, fills[0])
            fills[0] += 1

    def step(t):
        mask, km = U.compress_weight(t, QUALITY, True, BITS, MODE)
        return mask, km

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()  # running well before the timed region: nvidia-smi needs a moment to deliver its first sample
    # ---- warm-up (the last warm-up step runs with every kernel bracketed by events: it names the dominant kernel
    # and gives the per-kernel breakdown; the timed region then brackets only that kernel, so that the event
    # overhead -- two records per launch -- stays out of `value`)
    done = 0
    refill(min(pool, W))
    warm_ktimes = {}
    for i in range(W):
        if i and i % pool == 0:
            refill(min(pool, W - i))
        if i == W - 1:
            ctx.set_kernel_timing(True)
        # same binding pattern as the timed loop: the previous step's outputs stay alive during the next call, so the
        # torch caching allocator holds two sets of output buffers before the timing starts (a first-time cudaMalloc of
        # the second set used to land in the second timed step: +10..30 ms)
        mask, km = step(bufs[i % pool])
    warm_ktimes = {name: [c, ms] for name, (c, ms) in ctx.last_kernel_times().items()}
    hbm = {name: v for name, v in warm_ktimes.items() if kernel_bytes(name, 1.0, 0.3) > 0}  # streaming kernels only
    dom_name = max(hbm.items(), key=lambda kv: kv[1][1])[0] if hbm else None
    dom_filter = dom_name.strip("()").split("<")[0] if dom_name else None
    # ---- timed region: K steps in chunks of `pool` fresh tensors
    ctx.set_kernel_timing(True, dom_filter)
    launches0 = ctx.total_launches()
    total_ms = 0.0
    n_iters = []
    step_ms = []
    launches = 0
    ktimes = {}
    phases = {}
    t_region0 = time.time()
    while done < K:
        cnt = min(pool, K - done)
        refill(cnt)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(torch.cuda.current_stream())
        for i in range(cnt):
            t_s = time.perf_counter()
            mask, km = step(bufs[i])
            # accounting only (host side, after the call returned; every call ends with a stream synchronize)
            step_ms.append(1e3 * (time.perf_counter() - t_s))
            n_iters.append(km.n_iter_)
            for kname, ms in km.profile.items():
                if kname != "launches":
                    phases[kname] = phases.get(kname, 0.0) + ms
        ev1.record(torch.cuda.current_stream())
        barrier()
        total_ms += ev0.elapsed_time(ev1)
        done += cnt
    t_region1 = time.time()
    clk = clocks.stop(t_region0, t_region1) if rank == 0 else None
    ktimes = {name: [c, ms] for name, (c, ms) in ctx.last_kernel_times().items()}
    launches = ctx.total_launches() - launches0
    ctx.set_kernel_timing(False)
    if dist is not None:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = args.n * K / (total_ms * 1e-3)

    # ---- result hash of the LAST timed step (shard independent: the driver compares it across N)
    result = result_hash(torch, dist, dev, lo, hi, mask, km, U.prune_weigth.last_threshold, U.prune_weigth.last_pruned, n_iters)

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, torch, U, n_local, rank, world, dist, dev)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0
    # ---- roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    survivors = km.n_nonzero / world  # n_nonzero is the global count
    s_frac = km.n_nonzero / float(args.n)
    dom_cnt, dom_ms = ktimes.get(dom_name, (1, 0.0))
    kbytes = kernel_bytes(dom_name, n_local, survivors)
    avg_ms = dom_ms / max(dom_cnt, 1)
    achieved = kbytes / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    step_bytes = 13.0 + 4.0 + BITS / 8.0 + 52.0 * s_frac  # prune 13 + quantize 4 + b/8 + 52 s
    moved_bytes = 4.0 + 9.0 + 4.0 * s_frac + 16.0 * s_frac + 0.5 + 4.0 + BITS / 8.0  # stats, fused apply, histogram path, emission
    floor_bytes = 13.0 + 8.0 + BITS / 8.0
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            # static value from the committed ncu capture of this kernel on this workload, NOT measured in this run
            "traffic": NCU_TRAFFIC.get(dom_name, [None])[0] if world == 1 and args.n == (1 << 30) else None,
            "traffic_source": NCU_TRAFFIC.get(dom_name, [None, None])[1],
            "kernel": dom_name, "kernel_launches_per_step": dom_cnt / K, "kernel_ms_per_launch": avg_ms,
            "kernel_algorithmic_bytes_per_launch": kbytes, "peak_source": peak_src,
            "step_algorithmic_bytes_per_weight": step_bytes,
            "step_achieved_gbs": step_bytes * args.n / world / (total_ms / K * 1e-3) / 1e9,
            "step_frac": step_bytes * args.n / world / (total_ms / K * 1e-3) / 1e9 / peak,
            # the same step against the bytes the pipeline actually moves (the sort of SURVEY 8d is
            # replaced by a key histogram) and against the algorithm-independent floor (prune 13 + quantize 8 + b/8)
            "step_frac_by_bytes": {
                "prescribed_34.5": step_bytes * args.n / world / (total_ms / K * 1e-3) / 1e9 / peak,
                "moved_%.1f" % moved_bytes: moved_bytes * args.n / world / (total_ms / K * 1e-3) / 1e9 / peak,
                "floor_%.1f" % floor_bytes: floor_bytes * args.n / world / (total_ms / K * 1e-3) / 1e9 / peak,
            },
        },
        "result_hash": result,
        "gpu_launches": int(launches),
        "clocks": clk,
        "n_iter": n_iters,
        "ms_each_step_host_clock": [round(v, 3) for v in step_ms],
        "survivor_fraction": s_frac,
        "phase_ms_per_step": {kname: ms / K for kname, ms in sorted(phases.items())},
        "kernel_ms_per_step": {kname: v[1] for kname, v in sorted(warm_ktimes.items(), key=lambda kv: -kv[1][1])},
        "kernel_ms_note": "per-kernel CUDA-event times of the last warm-up step; the roofline kernel is timed inside the timed region",
    }
    if e2e is not None:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu:
        threads = cpu_threads()
        t_cpu, it_cpu, kind, engine = cpu_pipeline(args.cpu_sample, threads)
        line["cpu_baseline"] = {"value": args.cpu_sample / t_cpu, "unit": UNIT, "cores": threads, "kind": kind, "engine": engine,
                                "sample": cpu_sample_text(args.cpu_sample, it_cpu, t_cpu)}
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()
    return 0


def result_hash(torch, dist, dev, lo, hi, mask, km, thr, n_pruned, n_iters):
    """CRC over everything a caller gets back, in a form that does not depend on how the tensor was sharded: the global
    scalars and codebook as they are, the per-element outputs (mask, packed codes) through position-weighted checksums
    summed over the ranks (int64 arithmetic modulo 2^64)."""
    import struct
    import zlib

    codes = km.packed_codes
    sums = torch.zeros(4, dtype=torch.int64, device=dev)
    step = 1 << 26
    for b0 in range(0, hi - lo, step):
        b1 = min(hi - lo, b0 + step)
        wgt = (torch.arange(lo + b0, lo + b1, device=dev, dtype=torch.int64) % 65521) + 1
        c = codes[b0:b1].to(torch.int64)
        m = mask.reshape(-1)[b0:b1].to(torch.int64)
        sums[0] += (c * wgt).sum()
        sums[1] += c.sum()
        sums[2] += (m * wgt).sum()
        sums[3] += m.sum()
    if dist is not None:
        dist.all_reduce(sums)
    sums = [int(v) for v in sums.cpu().tolist()]
    blob = struct.pack("<dq", float(thr), int(n_pruned)) + np.ascontiguousarray(km.cluster_centers_).tobytes() + \
        np.ascontiguousarray(km.code_histogram).tobytes() + struct.pack("<%dq" % len(n_iters), *n_iters) + struct.pack("<4q", *sums)
    return {"crc32": "%08x" % (zlib.crc32(blob) & 0xffffffff), "threshold": float(thr), "n_pruned": int(n_pruned),
            "centroids_crc32": "%08x" % (zlib.crc32(np.ascontiguousarray(km.cluster_centers_).tobytes()) & 0xffffffff),
            "histogram_crc32": "%08x" % (zlib.crc32(np.ascontiguousarray(km.code_histogram).tobytes()) & 0xffffffff),
            "codes_weighted_sum": sums[0], "codes_sum": sums[1], "mask_weighted_sum": sums[2], "mask_sum": sums[3]}


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the round's `ncu --set full` captures on this workload
# (2^30 weights, 1 GPU; profiles/r2_summary.md)
NCU_TRAFFIC = {
    "np_tree_kernel<VisitApplyQuant>": [11.01e9, "profiles/r2_summary.md section 3 (ncu --set full capture prof_r2_final.ncu-rep: dram__bytes_read.sum 4.334 GB + dram__bytes_write.sum 6.674 GB)"],
    "kh_scatter_kernel": [2.90e9, "profiles/r2_summary.md section 3 (prof_r2_final.ncu-rep: 1.819 GB read + 1.083 GB written)"],
    "(rs_scatter_kernel<A, B>)": [3.135e9, "profiles/r1_summary.md (gpurun_out/prof_r1_scatter.ncu-rep)"],
}


def kernel_bytes(name, n, n_nz):
    """Algorithmic bytes one launch of `name` moves."""
    table = {
        "np_tree_kernel<VisitStats>": 4.0 * n,                          # mean + std estimate: one read
        "np_tree_kernel<VisitCenSqApply>": 9.0 * n,                     # read, pruned write, mask
        "np_tree_kernel<VisitApplyQuant>": 9.0 * n + 4.0 * n_nz,        # ... + the compacted survivors
        "np_tree_kernel<VisitQuant>": 4.0 * n + 4.0 * n_nz,             # unfused k-means prologue
        "kh_scatter_kernel": 8.0 * n_nz,                                # read + write every key
        "kh_count_kernel": 4.0 * n_nz,
        "kh_hist_kernel<false>": 4.0 * n_nz,
        "(rs_scatter_kernel<A, B>)": 8.0 * n_nz,
        "rs_count_kernel<true>": 4.0 * n_nz,
        "rs_count_kernel<false>": 4.0 * n_nz,
        "(emit_kernel<VEC, INERTIA, BITS>)": 4.0 * n + BITS / 8.0 * n,
    }
    return table.get(name, 0.0)


def run_e2e(args, torch, U, n_local, rank, world, dist, dev):
    steps = max(1, min(args.steps, args.e2e_steps))
    host = torch.empty(n_local, dtype=torch.float32).pin_memory()
    out_mask = torch.empty((n_local + 7) // 8, dtype=torch.uint8).pin_memory()  # the 1-bit mask of the compressed-layer format
    out_packed = torch.empty(n_local * BITS // 8, dtype=torch.uint8).pin_memory()
    src = (torch.randn(n_local, generator=torch.Generator().manual_seed(SEED + 100 + rank)) * SIGMA) if n_local <= (1 << 26) else None
    total = 0.0
    h2d = d2h = 0
    WARM = 2  # untimed calls: the first grows the workspace arena, the second runs with the regrown block
    for i in range(steps + WARM):
        if src is not None:
            host.copy_(src)
        else:  # big tensors: fill from the device generator (untimed)
            tmp = torch.empty(n_local, dtype=torch.float32, device=dev).normal_(0.0, SIGMA)
            host.copy_(tmp)
            del tmp
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        arr = host.numpy()
        mask, km = U.compress_weight(arr, QUALITY, True, BITS, MODE, update_weights=False, out_mask=out_mask.numpy(),
                                     out_packed=out_packed.numpy(), mask_bits=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if os.environ.get("NNC_BENCH_VERBOSE"):
            sys.stderr.write("[e2e] call %d: %.1f ms %s\n" % (i, dt * 1e3, {k: round(v, 2) for k, v in km.profile.items()}))
        if i >= WARM:
            total += dt
        # one fused call: w in; mask + packed codes (+ k centroids, histogram) out
        h2d = 4 * n_local
        d2h = (n_local + 7) // 8 + km.packed_codes.nbytes + 4 * km.n_clusters + 8 * km.n_clusters
    if dist is not None:
        t = torch.tensor([total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total = float(t.item())
    return {"value": args.n * steps / total, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "steps": steps, "ms_per_step": 1e3 * total / steps, "api": "utility.compress_weight(pinned host ndarray, update_weights=False, pinned out_mask/out_packed, mask_bits=True) -> nnc_compress_f32: "
                   "float32 weights in; the compressed layer out (1-bit mask, packed 8-bit codes, codebook, histogram)"}


# ---------------------------------------------------------------------------------------------------------
# the other configs of BASELINE.json (--config c1 | c2 | c3 | c5); the default (c4) is run_b200 above
# ---------------------------------------------------------------------------------------------------------
def run_config(args):
    """C1 LeNet300-100 prune + 2-bit density k-means; C2 LeNet5 prune + 4-bit linear; C3 4096 x 4096 prune + 5-bit forgy
    (seeded); C5 trained-quantization gradient sum (2^28 gradients, 256 clusters, shards over the ranks).  Same JSON
    contract as the default config.  Inputs smaller than the L2 are followed by an L2 flush between the timed steps."""
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch

    from neural_network_compression_b200 import _native as N
    from neural_network_compression_b200.common import utility as U
    from tests import _data as D

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        if args.config != "c5":
            raise SystemExit("--config %s is a single-GPU configuration" % args.config)
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
        U.init_distributed(device=local)
    ctx = N.default_context(local)
    K, W = args.steps, args.warmup
    cfgname = args.config
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    metric, unit = METRIC, UNIT
    cpu = None
    if cfgname in ("c1", "c2"):
        model = D.lenet300_tensors() if cfgname == "c1" else D.lenet5_tensors()
        bits, mode = (2, "density") if cfgname == "c1" else (4, "linear")
        host, qs = [], []
        for name, w, b, (qw, qb) in model:
            host += [w, b]
            qs += [qw, qb]
        units = sum(t.size for t in host)
        master = [torch.from_numpy(t).to(dev) for t in host]
        work = {}

        def prepare():
            work["t"] = [m.clone() for m in master]

        def step():
            return U.compress_model(work["t"], qs, True, bits, mode)

        def e2e_step():
            return U.compress_model([t.copy() for t in host], qs, True, bits, mode)

        h2d = 4 * units
        cb = bits if mode == "linear" else bits + 1  # (density init has 2^bits + 1 centroids)
        d2h = units * 4 + units + (units * cb + 7) // 8  # pruned weights back in place, mask, packed codes
        workload = ("%s: every kernel and bias pruned with the trainer's thresholds, then %d-bit %s-init k-means, all "
                    "tensors in one native batched call (utility.compress_model -> nnc_compress_many_f32; mask, codebook and "
                    "packed codes out)" %
                    ("LeNet300-100 (784-300-100-10), 6 tensors" if cfgname == "c1" else "LeNet5 conv+dense layers, 8 tensors", bits, mode))
        sfrac = 0.35
        step_bytes = 13.0 + 4.0 + cb / 8.0 + 52.0 * sfrac  # prune + quantize, packed codes out

        def cpu():
            from oracle import oracle as O
            import warnings

            from sklearn.cluster import KMeans
            from threadpoolctl import threadpool_limits

            threads = cpu_threads()
            ts = [t.copy() for t in host]
            with threadpool_limits(limits=threads), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                t0 = time.perf_counter()
                for t, q in zip(ts, qs):
                    thr = np.std(t) * q
                    m = np.abs(t) < thr
                    t[m] = 0
                    if t.size < 2 ** bits + 1:
                        continue
                    if mode == "density":
                        flat = t.flatten()
                        nz = np.delete(flat, np.nonzero(flat == 0)[0], axis=0)  # trainer.py:55-59
                        space = O.init_centroids(t, bits, "density", O.get_weight_distribution(nz))  # utility.py:334-392, 210-223 (C restatement)
                    else:
                        space = np.linspace(np.min(t), np.max(t), num=2 ** bits)
                    km = KMeans(n_clusters=len(space), init=space.reshape(-1, 1), n_init=1, algorithm="lloyd").fit(t.reshape(-1, 1))
                    _ = km.cluster_centers_[km.labels_].reshape(t.shape)
                dt = time.perf_counter() - t0
            return dt, units, threads, "the whole model, to convergence (NumPy prune + scikit-learn KMeans per tensor; CDF / density init through the C restatement)"
    elif cfgname == "c3":
        w3 = D.gaussian(4096 * 4096, seed=1234).reshape(4096, 4096)
        units = w3.size
        master = torch.from_numpy(w3).to(dev)
        work = {}

        def prepare():
            work["t"] = master.clone()

        def step():
            np.random.seed(0)
            return U.compress_weight(work["t"], 1.0, True, 5, "forgy")

        def e2e_step():
            np.random.seed(0)
            return U.compress_weight(w3.copy(), 1.0, True, 5, "forgy")

        h2d, d2h = 4 * units, units + units * 5 // 8
        workload = "synthetic 4096x4096 fp32 dense layer N(0,0.02^2), std-threshold prune q=1 + 5-bit forgy (np.random.seed(0)) k-means, mask + packed 5-bit codes out"
        step_bytes = 13.0 + 4.0 + 5 / 8.0 + 52.0 * 0.317

        def cpu():
            import warnings

            from sklearn.cluster import KMeans
            from threadpoolctl import threadpool_limits

            threads = cpu_threads()
            t = w3.copy()
            with threadpool_limits(limits=threads), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                t0 = time.perf_counter()
                thr = np.std(t) * 1.0
                t[np.abs(t) < thr] = 0
                np.random.seed(0)
                space = np.random.choice(t.flatten(), size=2 ** 5)  # utility.py:224-226
                km = KMeans(n_clusters=len(space), init=space.reshape(-1, 1), n_init=1, algorithm="lloyd", max_iter=args.cpu_max_iter).fit(t.reshape(-1, 1))
                dt = time.perf_counter() - t0
            return dt, units, threads, "the whole tensor, k-means capped at %d of its ~227 Lloyd iterations (%d run): %.1f s" % (args.cpu_max_iter, km.n_iter_, dt)
    else:  # c5
        n_total = 1 << 28
        n_local = n_total // world
        units = n_total
        metric, unit = "gradient elements/sec per-cluster segmented sum (256 clusters)", "elements/s"
        g = torch.empty(n_local, device=dev).normal_(0, 1e-3, generator=torch.Generator(device=dev).manual_seed(7 + rank))
        gen = torch.Generator(device=dev).manual_seed(8 + rank)
        codes = torch.randint(0, 256, (n_local,), device=dev, generator=gen, dtype=torch.int32)
        if args.c5_codes == "skewed":  # the code histogram of a pruned layer: 68 % of the weights in one cluster
            codes = torch.where(torch.rand(n_local, device=dev, generator=gen) < 0.683, torch.full_like(codes, 126), codes)
        packed = codes.to(torch.uint8)
        ref = torch.zeros(256, dtype=torch.float64, device=dev).index_add_(0, codes.long(), g.double())
        if dist is not None:
            dist.all_reduce(ref)
        del codes
        g_host, p_host = None, None

        def prepare():
            pass

        def step():
            return U.cluster_gradient_sum(g, packed, 256, 8)

        def e2e_step():
            return U.cluster_gradient_sum(g_host, p_host, 256, 8)

        h2d, d2h = 5 * n_local, 8 * 256
        workload = "trained-quantization step: 2^28 fp32 gradients, 8-bit codes (%s histogram), per-cluster segmented sum, 256 clusters, contiguous shards" % args.c5_codes
        step_bytes = 5.0

        def cpu():
            ns = 1 << 25
            gs = (np.random.RandomState(7).randn(ns) * 1e-3).astype(np.float32)
            cs = np.random.RandomState(8).randint(0, 256, size=ns)
            t0 = time.perf_counter()
            np.bincount(cs, weights=gs, minlength=256)
            dt = time.perf_counter() - t0
            return dt, ns, 1, "np.bincount(codes, weights=grad, minlength=256) on a 2^25-element sample (single threaded NumPy)"

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    out = None
    for i in range(W):
        prepare()
        if i == W - 1:
            ctx.set_kernel_timing(True)
        out = step()
    torch.cuda.synchronize()
    ktimes = {}
    # worker threads have their own contexts in the batched mode: kernel times come from this thread's context only
    ktimes = {name: ms / max(c, 1) * c for name, (c, ms) in ctx.last_kernel_times().items()}
    ctx.set_kernel_timing(False)
    launches0 = N.total_launches_all()  # (the batched configs run on worker threads with contexts of their own)
    total_ms = 0.0
    t_region0 = time.time()
    for i in range(K):
        prepare()
        if cfgname != "c5":
            flush_buf.fill_(i & 0xff)  # L2 flush: the inputs are smaller than the 126 MB L2
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(torch.cuda.current_stream())
        out = step()
        ev1.record(torch.cuda.current_stream())
        barrier()
        total_ms += ev0.elapsed_time(ev1)
    t_region1 = time.time()
    clk = clocks.stop(t_region0, t_region1) if rank == 0 else None
    launches = N.total_launches_all() - launches0
    if dist is not None:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = units * K / (total_ms * 1e-3)
    # end to end: host arrays in, host results out
    e2e = None
    if not args.no_e2e:
        if cfgname == "c5":
            g_host, p_host = g.cpu().numpy(), packed.cpu().numpy()
        for _ in range(2):  # untimed: the first call grows the workspace arenas, the second runs with the regrown blocks
            e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_e2e = max(1, min(K, args.e2e_steps))
        for _ in range(n_e2e):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": units * n_e2e / dt, "unit": unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "steps": n_e2e, "ms_per_step": 1e3 * dt / n_e2e, "api": "the same public call on host (NumPy) arrays"}
    check = None
    if cfgname == "c5":
        got = out
        check = float(np.abs(got - ref.cpu().numpy()).max() / np.abs(ref.cpu().numpy()).max())
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    dom = max(ktimes.items(), key=lambda kv: kv[1])[0] if ktimes else None
    per_rank_units = units / world
    achieved = step_bytes * per_rank_units / (total_ms / K * 1e-3) / 1e9
    line = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": total_ms / K,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "name": cfgname, "units_per_step": units,
                   "l2": "inputs > L2, no flush" if cfgname == "c5" else "256 MB written between timed steps (L2 flush)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "kernel": "whole step" if cfgname != "c5" else "segsum_warp_kernel (+ fold)",
                     "algorithmic_bytes_per_unit": step_bytes, "dominant_kernel_by_time": dom,
                     "note": "latency bound at this size: tens of microsecond-scale launches per tensor" if cfgname in ("c1", "c2") else None},
        "gpu_launches": int(launches), "clocks": clk,
        "kernel_ms_last_warmup_step": {k2: v for k2, v in sorted(ktimes.items(), key=lambda kv: -kv[1])[:12]},
    }
    if check is not None:
        line["max_rel_err_vs_float64_index_add"] = check
    if e2e is not None:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu and cpu is not None:
        dt, n_cpu, threads, sample = cpu()
        line["cpu_baseline"] = {"value": n_cpu / dt, "unit": unit, "cores": threads, "kind": "reference", "sample": sample}
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=1 << 30, help="total weights over all ranks")
    ap.add_argument("--pool", type=int, default=8, help="distinct input tensors kept resident")
    ap.add_argument("--cpu-sample", type=int, default=1 << 22, help="weights of the sample the CPU arms time")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--config", default="c4", choices=["c1", "c2", "c3", "c4", "c5"], help="BASELINE.json configs; c4 (1B-weight layer) is the metric's")
    ap.add_argument("--c5-codes", default="uniform", choices=["uniform", "skewed"])
    ap.add_argument("--cpu-max-iter", type=int, default=30, help="c3: Lloyd iterations the CPU baseline is capped at")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if args.config != "c4":
        return run_config(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
