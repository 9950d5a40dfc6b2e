import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from flownet2_b200 import ops
from oracle import oracle as O
r = np.random.default_rng(0)
for (N, C, H, W, cl) in [(1, 16, 16, 16, True), (1, 256, 10, 12, False), (2, 64, 40, 56, True)]:
    a = r.standard_normal((N, C, H, W)).astype(np.float32); b = r.standard_normal((N, C, H, W)).astype(np.float32)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    if cl:
        ta, tb = ta.contiguous(memory_format=torch.channels_last), tb.contiguous(memory_format=torch.channels_last)
    got = ops.correlation(ta, tb, 20, 1, 20, 1, 2)
    torch.cuda.synchronize()
    want = O.correlation_fwd(a, b, 20, 1, 20, 1, 2, 0, False)
    print((N, C, H, W, cl), "max err", float(np.abs(got.contiguous().cpu().numpy() - want).max()), flush=True)
