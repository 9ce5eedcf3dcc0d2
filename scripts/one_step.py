"""One compress_weight step on a device-resident Gaussian tensor (profiling target)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_network_compression_b200.common import utility as U
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 26
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
t = torch.empty(n, device='cuda').normal_(0, 0.02, generator=torch.Generator(device='cuda').manual_seed(2024))
for rep in range(reps):
    tt = t.clone()
    torch.cuda.synchronize(); t0 = time.time()
    mask, km = U.compress_weight(tt, 1.0, True, 8, 'linear')
    torch.cuda.synchronize(); print('step %.2f ms' % (1e3 * (time.time() - t0)), km.n_iter_, {k: round(v, 3) for k, v in km.profile.items()})
