import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from neural_network_compression_b200.common import utility as U
from neural_network_compression_b200 import _native as N
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
t = torch.empty(n, device='cuda').normal_(0, 0.02, generator=torch.Generator(device='cuda').manual_seed(2024))
for rep in range(2):
    tt = t.clone()
    torch.cuda.synchronize(); t0 = time.time()
    mask, km = U.compress_weight(tt, 1.0, True, 8, 'linear')
    torch.cuda.synchronize(); print('step', time.time() - t0, km.n_iter_, km.n_relocations, {k: round(v, 3) for k, v in km.profile.items()})
