"""Kernel launches per call of the small-tensor path (GPU box): which kernels, how many times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neural_network_compression_b200 import _native as N
from neural_network_compression_b200.common import utility as U
from tests import _data as D

dev = torch.device("cuda", 0)
ctx = N.default_context(0)
for label, model, bits, mode in (("c1", D.lenet300_tensors(), 2, "density"), ("c2", D.lenet5_tensors(), 4, "linear")):
    name, w, b, (qw, qb) = model[0] if label == "c1" else model[2]
    t = torch.from_numpy(w).to(dev)
    for rep in range(2):
        tt = t.clone()
        ctx.set_kernel_timing(True)
        l0 = ctx.total_launches()
        m = U.prune_weigth(tt, qw, True)
        l1 = ctx.total_launches()
        cdfs = U.get_weight_distribution(tt, skip_zeros=True) if mode == "density" else None
        l2 = ctx.total_launches()
        ris, km = U.get_quantized_weight(tt, bits, mode, cdfs)
        l3 = ctx.total_launches()
        torch.cuda.synchronize()
        kt = ctx.last_kernel_times()
        ctx.set_kernel_timing(False)
    print(label, tuple(w.shape), "launches: prune", l1 - l0, "distribution", l2 - l1, "quantize", l3 - l2, "iterations", km.n_iter_)
    for k, (c, ms) in sorted(kt.items(), key=lambda kv: -kv[1][0]):
        print("    %-60s x%-3d %.3f ms" % (k, c, ms))
