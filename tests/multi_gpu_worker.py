"""Worker of the multi-GPU parity test: launched by torchrun, one rank per GPU.

Every rank builds the same full tensor, keeps the slice `shard_range` assigns to it, runs the sharded path and
compares with the single-rank result it computes itself on its own GPU (a second, rank-less context)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from neural_network_compression_b200 import _native as N
    from neural_network_compression_b200.common import utility as U

    n = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 22) + 12345
    failures = []
    for bits, seed in ((8, 1), (4, 2), (2, 3)):
        g = torch.Generator(device="cuda").manual_seed(1000 + seed)
        full = torch.empty(n, device="cuda").normal_(0.0, 0.02, generator=g)
        full[::9] = 0.0
        # ---- single-rank reference on this GPU
        N._tls.ctx = {}
        ref_w = full.clone()
        ref_mask, ref_km = U.compress_weight(ref_w, 1.0, True, bits, "linear")
        ref_thr = U.prune_weigth.last_threshold
        ref_ris, ref_full_km = U.get_quantized_weight(ref_w, bits, "linear")
        # ---- sharded
        N._tls.ctx = {}
        # bits 4 runs over NCCL all-reduces between the update phases, the others over the in-kernel peer exchange
        dctx = U.init_distributed(peer_exchange=(bits != 4))
        b, e = U.shard_range(n, rank, world)
        mine = full[b:e].clone()
        mask, km = U.compress_weight(mine, 1.0, True, bits, "linear")
        ok = True
        bad = []

        def check(name, cond):
            nonlocal ok
            if not cond:
                bad.append(name)
                ok = False

        check("thr", U.prune_weigth.last_threshold == ref_thr)
        check("mask", bool(torch.equal(mask, ref_mask[b:e])) and bool(torch.equal(mine, ref_w[b:e])))
        check("mean %r %r" % (km.mean, ref_km.mean), km.mean == ref_km.mean)
        check("n_nonzero %r %r" % (km.n_nonzero, ref_km.n_nonzero), km.n_nonzero == ref_km.n_nonzero)
        check("centers", km.cluster_centers_.tobytes() == ref_km.cluster_centers_.tobytes())
        check("iters %r %r" % ((km.n_iter_, km.n_relocations), (ref_km.n_iter_, ref_km.n_relocations)),
              km.n_iter_ == ref_km.n_iter_ and km.n_relocations == ref_km.n_relocations)
        check("hist", bool(np.array_equal(km.code_histogram, ref_km.code_histogram)))
        per = km.code_bits
        if (b * per) % 8 == 0:
            check("packed", bool(torch.equal(km.packed_codes, ref_km.packed_codes[b * per // 8: (e * per + 7) // 8])))
        # the reference-signature helper on the shard: labels / ris of the slice, inertia global
        ris, km2 = U.get_quantized_weight(mine, bits, "linear")
        check("labels2", bool(torch.equal(km2.labels_, ref_full_km.labels_[b:e])) and bool(torch.equal(ris, ref_ris[b:e])))
        check("inertia2", km2.inertia_ == ref_full_km.inertia_)
        # trained-quantization gradient sum over the shards == over the whole tensor (same grid-independent tolerance)
        gg = torch.empty(n, device="cuda").normal_(0.0, 1e-3, generator=g)
        gs = U.cluster_gradient_sum(gg[b:e].clone(), km2.labels_, km2.n_clusters)
        full_sum = torch.zeros(km2.n_clusters, dtype=torch.float64, device="cuda").index_add_(0, ref_full_km.labels_.long(), gg.double())
        mag = torch.zeros(km2.n_clusters, dtype=torch.float64, device="cuda").index_add_(0, ref_full_km.labels_.long(), gg.double().abs())
        check("gradsum", bool(np.all(np.abs(gs - full_sum.cpu().numpy()) <= 1e-13 * mag.cpu().numpy() + 1e-300)))
        # np.std of the sharded tensor
        m, v, s = U.weight_stats(full[b:e].clone())
        check("stats", (m, v, s) == tuple(np.float32(x) for x in (np.mean(full.cpu().numpy()), np.var(full.cpu().numpy()), np.std(full.cpu().numpy()))))
        if not ok:
            failures.append((bits, rank))
        print("rank %d bits %d iters %d reloc %d thr %.9g peer_exchange=%s %s" % (rank, bits, km.n_iter_, km.n_relocations, ref_thr,
                                                                                    dctx.peer_exchange, "OK" if ok else "MISMATCH " + "; ".join(bad)), flush=True)
    # density / forgy init on shards: global min / max + 31-bin histogram all-reduce, forgy owner gather
    g = torch.Generator(device="cuda").manual_seed(77)
    full = torch.empty(n, device="cuda").normal_(0.0, 0.02, generator=g)
    for mode, bits in (("density", 2), ("forgy", 5)):
        N._tls.ctx = {}
        ref_w = full.clone()
        np.random.seed(5)
        ref_mask, ref_km = U.compress_weight(ref_w, 1.0, True, bits, mode)
        N._tls.ctx = {}
        dctx = U.init_distributed()
        b, e = U.shard_range(n, rank, world)
        mine = full[b:e].clone()
        np.random.seed(5)
        mask, km = U.compress_weight(mine, 1.0, True, bits, mode)
        bad = []
        if not bool(torch.equal(mask, ref_mask[b:e])):
            bad.append("mask")
        if km.cluster_centers_.tobytes() != ref_km.cluster_centers_.tobytes():
            bad.append("centers")
        if (km.n_iter_, km.n_relocations) != (ref_km.n_iter_, ref_km.n_relocations):
            bad.append("iters %r %r" % ((km.n_iter_, km.n_relocations), (ref_km.n_iter_, ref_km.n_relocations)))
        if not np.array_equal(km.code_histogram, ref_km.code_histogram):
            bad.append("hist")
        if bad:
            failures.append((mode, rank))
        print("rank %d %s-%d iters %d k %d %s" % (rank, mode, bits, km.n_iter_, km.n_clusters, "OK" if not bad else "MISMATCH " + "; ".join(bad)), flush=True)
    t = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(t)
    dist.destroy_process_group()
    if int(t.item()):
        sys.exit(1)


if __name__ == "__main__":
    main()
