"""Configs C1-C3 of BASELINE.json / SURVEY.md 8d on one GPU, device-resident tensors, through the public helpers:
   C1 LeNet300-100: prune with the trainer's thresholds + 2-bit density k-means (CDF of the non-zeros), six tensors
   C2 LeNet5: prune (q = 1 / 0.1) + 4-bit linear k-means
   C3 4096 x 4096 N(0, 0.02^2): std-threshold prune q = 1 + 5-bit forgy (seeded) k-means on the pruned tensor
Prints one JSON line per config: milliseconds (best of 5, host clock around the calls, outputs on the device)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neural_network_compression_b200.common import utility as U
from tests import _data as D


def timed(fn, reps=5):
    best = 1e30
    out = None
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, 1e3 * (time.perf_counter() - t0))
    return best, out


def layers_job(tensors, bits, mode, with_cdf):
    def run():
        iters = []
        for w, b, (qw, qb) in tensors:
            for t, q in ((w.clone(), qw), (b.clone(), qb)):
                U.prune_weigth(t, q, True)
                cdfs = U.get_weight_distribution(t, skip_zeros=True) if with_cdf and bool((t != 0).any()) else None
                if mode == "density" and cdfs is None:
                    continue
                ris, km = U.get_quantized_weight(t, bits, mode, cdfs)
                if km is not None:
                    iters.append(int(km.n_iter_))
        return iters
    return run


def main():
    dev = torch.device("cuda", 0)
    c1 = [(torch.from_numpy(w).to(dev), torch.from_numpy(b).to(dev), q) for _, w, b, q in D.lenet300_tensors()]
    ms, iters = timed(layers_job(c1, 2, "density", True))
    n1 = sum(w.numel() + b.numel() for w, b, _ in c1)
    print(json.dumps({"config": "C1 LeNet300-100 prune + 2-bit density k-means (6 tensors)", "weights": n1, "ms": ms, "n_iter": iters,
                      "weights_per_s": n1 / ms * 1e3}))
    c2 = [(torch.from_numpy(w).to(dev), torch.from_numpy(b).to(dev), q) for _, w, b, q in D.lenet5_tensors()]
    ms, iters = timed(layers_job(c2, 4, "linear", False))
    n2 = sum(w.numel() + b.numel() for w, b, _ in c2)
    print(json.dumps({"config": "C2 LeNet5 prune + 4-bit linear k-means (8 tensors)", "weights": n2, "ms": ms, "n_iter": iters,
                      "weights_per_s": n2 / ms * 1e3}))
    w3 = torch.from_numpy(D.gaussian(4096 * 4096, seed=1234)).to(dev).reshape(4096, 4096)

    def c3():
        t = w3.clone()
        np.random.seed(0)
        mask, km = U.compress_weight(t, 1.0, True, 5, "forgy")
        return [int(km.n_iter_)]
    ms, iters = timed(c3, reps=3)
    print(json.dumps({"config": "C3 4096x4096 std-prune q=1 + 5-bit forgy (seed 0) k-means, packed codes out", "weights": w3.numel(), "ms": ms,
                      "n_iter": iters, "weights_per_s": w3.numel() / ms * 1e3}))


main()
