"""Calibrates the round-toward-zero compensation of the tcgen05 conv engine: mean signed relative error of the raw
conv output (no bias, no ReLU) vs float64, per (NT, KD), with FN2_TC_COMP=0 (run once per KD via env)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from flownet2_b200 import ops
r = np.random.default_rng(0)
cl = torch.channels_last
for (N, Ci, H, W, Co, k) in [(2, 256, 16, 32, 128, 3), (2, 512, 16, 32, 256, 3), (2, 256, 16, 32, 64, 3), (2, 128, 32, 32, 32, 3), (2, 96, 32, 32, 16, 3)]:
    x = torch.from_numpy(r.standard_normal((N, Ci, H, W)).astype(np.float32)).cuda().contiguous(memory_format=cl)
    x = torch.where(x > 0, x, 0.1 * x)                      # post-leaky-ReLU-like activations
    w = torch.from_numpy((r.standard_normal((Co, Ci, k, k)) * np.sqrt(2.0 / (Ci * k * k))).astype(np.float32)).cuda()
    ref = torch.nn.functional.conv2d(x.double(), w.double(), padding=1)
    got = ops.conv2d(x, w, None, 1, 1, False, None, 2).double()
    simt = ops.conv2d(x, w, None, 1, 1, False, None, 1).double()
    den = ref.abs().mean()
    print("KD=%s COMP=%s  Ci=%4d Co=%4d (NT=%3d): tc shrink %+.3e  |err| %.3e   simt shrink %+.3e |err| %.3e" % (
        os.environ.get("FN2_TC_KD", "def"), os.environ.get("FN2_TC_COMP", "1"), Ci, Co,
        128 if Co % 128 == 0 else 64 if Co % 64 == 0 else 32 if Co % 32 == 0 else 16,
        float(((got - ref) * ref.sign()).mean() / den), float((got - ref).abs().mean() / den),
        float(((simt - ref) * ref.sign()).mean() / den), float((simt - ref).abs().mean() / den)), flush=True)
