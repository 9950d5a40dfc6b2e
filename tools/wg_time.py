#!/usr/bin/env python
"""Driver for ncu: the tensor-core weight gradient (and the fast correlation backward) at FlowNet2-C training shapes (batch 8, 448x320)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flownet2_b200 import ops

torch.manual_seed(0)
cl = torch.channels_last
# conv3_1: 473 -> 256, 3x3 at 40x56
x = torch.randn(8, 473, 40, 56, device="cuda").contiguous(memory_format=cl)
w = torch.randn(256, 473, 3, 3, device="cuda") * 0.02
dy = torch.randn(8, 256, 40, 56, device="cuda").contiguous(memory_format=cl)
for _ in range(3):
    ops.conv2d_backward(x, w, dy, 1, 1, False, need_input_grad=False)
# conv2: 64 -> 128, 5x5 stride 2 at 160x224 -> 80x112 (tap pairs, parity planes)
x = torch.randn(8, 64, 160, 224, device="cuda").contiguous(memory_format=cl)
w = torch.randn(128, 64, 5, 5, device="cuda") * 0.02
dy = torch.randn(8, 128, 80, 112, device="cuda").contiguous(memory_format=cl)
for _ in range(3):
    ops.conv2d_backward(x, w, dy, 2, 2, False, need_input_grad=False)
# correlation backward at (8, 256, 40, 56), d = 21, stride_2 = 2
a = torch.randn(8, 256, 40, 56, device="cuda").contiguous(memory_format=cl)
b = torch.randn(8, 256, 40, 56, device="cuda").contiguous(memory_format=cl)
td = torch.randn(8, 441, 40, 56, device="cuda").contiguous(memory_format=cl)
for _ in range(3):
    ops.correlation_backward(a, b, td, 20, 1, 20, 1, 2)
torch.cuda.synchronize()
print("done")
