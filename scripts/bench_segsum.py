"""Config C5 of BASELINE.json: trained-quantization gradient sum, 2^28 gradients, 256 clusters, 8-bit codes.
One process per GPU (torchrun) or a single GPU.  Prints one JSON line (rank 0)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch

def main():
    world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    from neural_network_compression_b200 import _native as N
    from neural_network_compression_b200.common import utility as U
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        U.init_distributed(device=local)
    n = (1 << 28) // world
    g = torch.empty(n, device="cuda").normal_(0, 1e-3, generator=torch.Generator(device="cuda").manual_seed(7 + rank))
    out = {}
    for name in ("uniform", "skewed"):
        gen = torch.Generator(device="cuda").manual_seed(8 + rank)
        codes = torch.randint(0, 256, (n,), device="cuda", generator=gen, dtype=torch.int32)
        if name == "skewed":  # the code histogram of a pruned layer: 68 % of the weights in one cluster
            codes = torch.where(torch.rand(n, device="cuda", generator=gen) < 0.683, torch.full_like(codes, 126), codes)
        packed = codes.to(torch.uint8)
        ctx = N.default_context(local)
        for _ in range(3):
            res = U.cluster_gradient_sum(g, packed, 256, 8)
        torch.cuda.synchronize()
        ms = []
        for _ in range(5):
            ctx.timer_start(); res = U.cluster_gradient_sum(g, packed, 256, 8); ms.append(ctx.timer_stop())
        ref = torch.zeros(256, dtype=torch.float64, device="cuda").index_add_(0, codes.long(), g.double())
        if world > 1:
            dist.all_reduce(ref)
        err = float(np.abs(res - ref.cpu().numpy()).max() / np.abs(ref.cpu().numpy()).max())
        best = min(ms)
        out[name] = {"ms": best, "GB/s_per_gpu": 5.0 * n / best / 1e6, "rel_err_vs_fp64_index_add": err}
    if rank == 0:
        print(json.dumps({"workload": "grad segsum 2^28 fp32 gradients, 8-bit codes, 256 clusters", "n_gpus": world, **out}))
    if world > 1:
        dist.destroy_process_group()

main()
