"""Wall-clock diagnosis of the batched many-small-tensor mode (GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neural_network_compression_b200.common import utility as U
from tests import _data as D

dev = torch.device("cuda", 0)
for label, model, bits, mode in (("c1", D.lenet300_tensors(), 2, "density"), ("c2", D.lenet5_tensors(), 4, "linear")):
    host, qs = [], []
    for name, w, b, (qw, qb) in model:
        host += [w, b]; qs += [qw, qb]
    master = [torch.from_numpy(t).to(dev) for t in host]
    def seq(ts):
        out = []
        for t, q in zip(ts, qs):
            t0 = time.perf_counter()
            m = U.prune_weigth(t, q, True)
            cdfs = U.get_weight_distribution(t, skip_zeros=True) if mode == "density" else None
            ris, km = U.get_quantized_weight(t, bits, mode, cdfs)
            torch.cuda.synchronize()
            out.append((tuple(t.shape), round(1e3 * (time.perf_counter() - t0), 2), None if km is None else km.n_iter_))
        return out
    for rep in range(3):
        ts = [m.clone() for m in master]; torch.cuda.synchronize(); t0 = time.perf_counter(); r = seq(ts); torch.cuda.synchronize()
        print(label, "sequential device ms", round(1e3 * (time.perf_counter() - t0), 2), r if rep == 2 else "", flush=True)
    for workers in (1, 4, 8):
        U._pool = None
        for rep in range(3):
            ts = [m.clone() for m in master]; torch.cuda.synchronize(); t0 = time.perf_counter()
            U.compress_tensors(ts, qs, True, bits, mode, workers=workers); torch.cuda.synchronize()
            print(label, "batched device workers", workers, "ms", round(1e3 * (time.perf_counter() - t0), 2), flush=True)
    for rep in range(2):
        ts = [t.copy() for t in host]; t0 = time.perf_counter(); U.compress_tensors(ts, qs, True, bits, mode, workers=8)
        print(label, "batched host ms", round(1e3 * (time.perf_counter() - t0), 2), flush=True)
