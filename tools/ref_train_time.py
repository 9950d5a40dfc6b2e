#!/usr/bin/env python
"""Times the REFERENCE's own GPU training step (its unmodified layer classes from oracle/_ref: im2col + cuBLAS convolutions, its
correlation / augmentation / loss kernels, Split layers inserted like Net::Init does) on the FlowNet2-C training graph of
BASELINE config 5, on this GPU.  Prints one JSON line.  Run as a subprocess of bench.py: a CHECK failure inside the reference
aborts the process."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    import flownet2_b200 as F
    from oracle import ref as R
    if not R.available():
        print(json.dumps({"unavailable": "oracle/_ref is not built"}))
        return
    R.set_mode(True, 0)
    R.set_seed(1701)
    cw, ch, dw, dh = 448, 320, 512, 384
    proto = F.fill_train_template(F.train_template("FlowNet2-C"), cw, ch, dw, dh, batch)
    net = R.RefNet(proto, None, phase=0, splits=True)
    r = np.random.default_rng(7)
    img0 = np.round(r.uniform(0, 255, (batch, 3, dh, dw))).astype(np.float32)
    img1 = np.clip(img0 + np.round(r.normal(0, 4, img0.shape)), 0, 255).astype(np.float32)
    gt = (4 * r.standard_normal((batch, 2, dh, dw))).astype(np.float32)
    times, losses = [], None
    warmup = 2
    for i in range(steps + warmup):
        t0 = time.perf_counter()
        net.forward(img0=img0, img1=img1, flow_gt=gt)
        net.backward()
        g = net.layers[[n for n, _, _ in net.layers].index("conv1")][2].params[0].get(diff=True)      # D2H read = synchronisation
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        losses = [float(net.blob("flow_loss%d" % l).reshape(-1)[0]) for l in (6, 5, 4, 3, 2)]
    mean = float(np.mean(times))
    print(json.dumps({"value": batch / mean, "unit": "frame-pairs/s", "ms_per_step": mean * 1e3, "steps": steps, "warmup": warmup, "ms_min": min(times) * 1e3, "ms_max": max(times) * 1e3,
                      "what": "reference layer classes (oracle/_ref) in GPU mode: FlowNet2-C training step, crop %dx%d from %dx%d, batch %d, "
                              "host-timed including its input uploads" % (cw, ch, dw, dh, batch),
                      "losses": losses, "grad_finite": bool(np.isfinite(g).all() and np.abs(g).max() > 0)}))


if __name__ == "__main__":
    main()
