#!/usr/bin/env python
"""Per-layer device time of one forward pass, in the style of `caffe time` (tools/caffe.cpp:346-385)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import flownet2_b200 as F  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="FlowNet2")
ap.add_argument("--width", type=int, default=1024)
ap.add_argument("--height", type=int, default=436)
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--top", type=int, default=40)
a = ap.parse_args()
net = F.Net(F.fill_template(F.model_template(a.model), a.width, a.height), None, F.TEST, batch=a.batch)
net.fill_params(1)
r = np.random.default_rng(0)
img = np.round(r.uniform(0, 255, (a.batch, 3, a.height, a.width))).astype(np.float32)
net.forward(img0=img, img1=img)
net.time_layers()
lt = net.time_layers()
work = net.layer_work()
rows = [(t, n, ty, f, b) for (n, ty, t), (_, _, f, b) in zip(lt, work)]
total = sum(r_[0] for r_ in rows)
print("total %.3f ms over %d layers, %d launches" % (total, len(rows), net.launches_per_forward))
bytype = {}
for t, n, ty, f, b in rows:
    e = bytype.setdefault(ty, [0.0, 0.0, 0.0]); e[0] += t; e[1] += f; e[2] += b
for ty, (t, f, b) in sorted(bytype.items(), key=lambda kv: -kv[1][0]):
    print("  %-18s %9.3f ms  %5.1f%%  %8.1f GFLOP/s  %8.1f GB/s" % (ty, t, 100 * t / total, f / t / 1e6 if t else 0, b / t / 1e6 if t else 0))
for t, n, ty, f, b in sorted(rows, reverse=True)[:a.top]:
    print("%-34s %-14s %8.3f ms  %9.1f GFLOP/s %8.1f GB/s  (%.2f GFLOP)" % (n, ty, t, f / t / 1e6 if t else 0, b / t / 1e6 if t else 0, f / 1e9))
