"""Seeded synthetic tensors of the shapes BASELINE.json names."""
import numpy as np

LENET300_SHAPES = {"dense1": (784, 300), "dense2": (300, 100), "out": (100, 10)}
LENET300_THRESHOLDS = {"dense1": (1, 0.1), "dense2": (1, 0.1), "out": (0.5, 0)}  # le_net_300_100_trainer.py:23-27
LENET5_SHAPES = {"conv1": (5, 5, 1, 20), "conv2": (5, 5, 20, 50), "dense": (2450, 256), "logits": (256, 10)}


def glorot_uniform(rng, shape):
    fan_in = int(np.prod(shape[:-1]))
    fan_out = int(shape[-1])
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def lenet300_tensors(seed=0, kind="glorot"):
    """[(name, kernel, bias, (q_kernel, q_bias))] for LeNet300-100."""
    rng = np.random.RandomState(seed)
    out = []
    for name, shape in LENET300_SHAPES.items():
        if kind == "glorot":
            w = glorot_uniform(rng, shape)
        else:
            w = (rng.randn(*shape) * 0.05).astype(np.float32)
        b = (rng.randn(shape[-1]) * 0.01).astype(np.float32)
        out.append((name, w, b, LENET300_THRESHOLDS[name]))
    return out


def lenet5_tensors(seed=1):
    rng = np.random.RandomState(seed)
    out = []
    for name, shape in LENET5_SHAPES.items():
        w = (rng.randn(*shape) * 0.05).astype(np.float32)
        b = (rng.randn(shape[-1]) * 0.05).astype(np.float32)
        out.append((name, w, b, (1, 0.1)))
    return out


def gaussian(n, seed=1234, sigma=0.02):
    return (np.random.RandomState(seed).randn(n) * sigma).astype(np.float32)


def prune_np(w, q, std_smooth=True):
    """The reference expression itself (utility.py:159-162), in NumPy.  In place; returns the mask."""
    thr = np.std(w) * q if std_smooth else q
    mask = np.abs(w) < thr
    w[mask] = 0
    return mask


def prune_cases():
    """(name, tensor, quality parameter) -- pruning inputs shared by the golden generator and the tests."""
    out = []
    for kind in ("glorot", "gauss"):
        for name, w, b, (qw, qb) in lenet300_tensors(kind=kind):
            out.append(("lenet300_%s_%s_w" % (kind, name), w, qw))
            out.append(("lenet300_%s_%s_b" % (kind, name), b, qb))
    for name, w, b, (qw, qb) in lenet5_tensors():
        out.append(("lenet5_%s_w" % name, w, qw))
        out.append(("lenet5_%s_b" % name, b, qb))
    out.append(("gauss_1m_q1", gaussian(1 << 20, seed=21), 1))
    out.append(("gauss_4096x4096_q1", gaussian(4096 * 4096).reshape(4096, 4096), 1.0))
    out.append(("gauss_100003_q025", gaussian(100003, seed=5), 0.25))
    return out


def kmeans_cases(big=True):
    """(name, pruned tensor, bits, mode, forgy seed) -- k-means inputs shared by the golden generator and tests.
    Tensors are pruned first exactly as the trainer does before quantize (trainer.py:177-193 then :42-72)."""
    out = []
    for name, w, b, (qw, qb) in lenet300_tensors():  # config 1: 2-bit density
        prune_np(w, qw)
        prune_np(b, qb)
        out.append(("c1_lenet300_%s_w" % name, w, 2, "density", 0))
        out.append(("c1_lenet300_%s_b" % name, b, 2, "density", 0))
    for name, w, b, (qw, qb) in lenet5_tensors():  # config 2: 4-bit linear
        prune_np(w, qw)
        prune_np(b, qb)
        out.append(("c2_lenet5_%s_w" % name, w, 4, "linear", 0))
        if b.size >= 17:
            out.append(("c2_lenet5_%s_b" % name, b, 4, "linear", 0))
    w = gaussian(300 * 1000, seed=77).reshape(300, 1000)
    prune_np(w, 1)
    for bits, mode in [(5, "forgy"), (8, "linear"), (3, "density"), (8, "density"), (1, "linear")]:
        out.append(("pruned300k_%s%d" % (mode, bits), w.copy(), bits, mode, 0))
    w = gaussian(200 * 1000, seed=78)
    for bits, mode in [(4, "linear"), (5, "forgy"), (2, "density")]:
        out.append(("dense200k_%s%d" % (mode, bits), w.copy(), bits, mode, 3))
    if big:
        w = gaussian(2048 * 2048).reshape(2048, 2048)  # config 3 at a quarter of the size (oracle time)
        prune_np(w, 1.0)
        out.append(("c3_2048x2048_forgy5", w, 5, "forgy", 0))
        # config 4 (the benchmark's own type) at 2^22 weights: N(0, 0.02^2), std-prune q = 1, 8-bit linear
        w = gaussian(1 << 22, seed=2024)
        prune_np(w, 1.0)
        out.append(("c4_4m_linear8", w, 8, "linear", 0))
    return out
