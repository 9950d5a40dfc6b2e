#!/usr/bin/env python
"""Small driver for `ncu --set full`: the correlation layer at the bench shapes and two representative convs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flownet2_b200 import ops

torch.manual_seed(0)
cl = torch.channels_last
for (N, H, W) in [(4, 56, 128), (8, 40, 56)]:
    a = torch.randn(N, 256, H, W, device="cuda").contiguous(memory_format=cl)
    b = torch.randn(N, 256, H, W, device="cuda").contiguous(memory_format=cl)
    out = torch.empty(N, 441, H, W, device="cuda").contiguous(memory_format=cl)
    for _ in range(3):
        ops.correlation(a, b, 20, 1, 20, 1, 2, out=out)
# conv3_1 of FlowNetC (473 -> 256, 3x3) and conv2 (64 -> 128, 5x5/2) at 1024x448, batch 4
x = torch.randn(4, 473, 56, 128, device="cuda").contiguous(memory_format=cl)
w = torch.randn(256, 473, 3, 3, device="cuda") * 0.02
bias = torch.zeros(256, device="cuda")
for _ in range(3):
    ops.conv2d(x, w, bias, 1, 1, False, 0.1, 1)
x = torch.randn(4, 64, 224, 512, device="cuda").contiguous(memory_format=cl)
w = torch.randn(128, 64, 5, 5, device="cuda") * 0.02
bias = torch.zeros(128, device="cuda")
for _ in range(3):
    ops.conv2d(x, w, bias, 2, 2, False, 0.1, 1)
torch.cuda.synchronize()
print("done")
