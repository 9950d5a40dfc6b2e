"""Single-chunk calibration of the tensor-core round-toward-zero loss (run with FN2_TC_COMP=0): a 1x1 convolution with
Ci = 32*KD is exactly one accumulation chunk per output, so the loss can be measured in ulps of the chunk value.
Compares the two compensation models offline (relative factor vs ulps of the value's binade)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from flownet2_b200 import ops
r = np.random.default_rng(0)
cl = torch.channels_last
kd = int(os.environ.get("FN2_TC_KD", "4"))
for Co in (128, 64, 32, 16):
    Ci = 32 * kd
    N, H, W = 2, 32, 64
    x = torch.from_numpy(r.standard_normal((N, Ci, H, W)).astype(np.float32)).cuda().contiguous(memory_format=cl)
    x = torch.where(x > 0, x, 0.1 * x)
    w = torch.from_numpy((r.standard_normal((Co, Ci, 1, 1)) * np.sqrt(2.0 / Ci)).astype(np.float32)).cuda()
    ref = torch.nn.functional.conv2d(x.double(), w.double()).cpu().numpy().ravel()
    got = ops.conv2d(x, w, None, 1, 0, False, None, 2).double().cpu().numpy().ravel()
    g32 = got.astype(np.float32)
    binade = (g32.view(np.uint32) & np.uint32(0xff800000)).view(np.float32).astype(np.float64)   # signed 2^e
    ulp = np.abs(binade) * 2.0 ** -23
    ok = np.abs(got) > 1e-3
    loss_rel = ((ref - got) / got)[ok].mean()
    loss_ulp = ((ref - got) * np.sign(got) / ulp)[ok].mean()
    e_none = np.abs(got - ref).mean() / np.abs(ref).mean()
    e_rel = np.abs(got * (1 + loss_rel) - ref).mean() / np.abs(ref).mean()
    e_ulp = np.abs(got + loss_ulp * ulp * np.sign(got) - ref).mean() / np.abs(ref).mean()
    print("KD=%d Co=%3d: loss rel %.3e | %.3f ulp ; mean|err|/mean|ref|: none %.3e  rel-model %.3e  ulp-model %.3e" % (
        kd, Co, loss_rel, loss_ulp, e_none, e_rel, e_ulp), flush=True)
