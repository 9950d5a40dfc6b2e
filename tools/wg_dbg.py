import sys; sys.path.insert(0, ".")
import numpy as np, torch
from flownet2_b200 import ops as OPS
import flownet2_b200 as F
r = np.random.default_rng(0)
N, Ci, H, W, Co, k, s, p = 2, 16, 12, 20, 24, 3, 1, 1
x = torch.from_numpy(r.standard_normal((N, Ci, H, W)).astype(np.float32)).cuda().contiguous(memory_format=torch.channels_last)
w = torch.from_numpy(r.standard_normal((Co, Ci, k, k)).astype(np.float32)).cuda()
dy = torch.from_numpy(r.standard_normal((N, Co, H, W)).astype(np.float32)).cuda().contiguous(memory_format=torch.channels_last)
gx, gw, gb = OPS.conv2d_backward(x, w, dy, s, p, False, need_input_grad=False)
torch.cuda.synchronize()
print("ok", float(gw.abs().max()))
