#!/usr/bin/env python
"""Times the REFERENCE's own GPU forward path (its unmodified layer classes from oracle/_ref: im2col + cuBLAS convolutions, its
correlation / warp / resample kernels) on a deploy prototxt of this repo, on this GPU.  Prints one JSON line.  Run as a subprocess of
bench.py: a CHECK failure inside the reference aborts the process, which must not take the bench line down.
    python tools/ref_forward_time.py MODEL WIDTH HEIGHT BATCH [STEPS] [WARMUP]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def main():
    model, width, height, batch = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    steps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
    warmup = int(sys.argv[6]) if len(sys.argv) > 6 else 2
    import torch
    import flownet2_b200 as F
    from oracle import ref as R
    from oracle.net import synth_weights
    if not R.available():
        print(json.dumps({"unavailable": "oracle/_ref is not built"}))
        return
    R.set_mode(True, 0)
    small = F.fill_template(F.model_template(model), 64, 64)
    proto = F.fill_template(F.model_template(model), width, height)
    weights, _ = synth_weights(small, 1701, proto)
    net = R.RefNet(proto, weights, batch=batch)
    r = np.random.default_rng(3)
    a = np.round(r.uniform(0, 255, (batch, 3, height, width))).astype(np.float32)
    b = np.clip(a + np.round(r.normal(0, 3, a.shape)), 0, 255).astype(np.float32)
    times, flow = [], None
    for i in range(warmup + steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        net.forward(img0=a, img1=b)
        flow = net.blob("predict_flow_final")            # device -> host read = synchronisation
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    mean = float(np.mean(times))
    print(json.dumps({"value": batch / mean, "unit": "frame-pairs/s", "ms_per_step": mean * 1e3, "ms_min": min(times) * 1e3,
                      "ms_max": max(times) * 1e3, "steps": steps, "warmup": warmup,
                      "what": "reference layer classes (oracle/_ref) in GPU mode on the same B200: %s %dx%d, %d pairs per step, host-timed "
                              "including its H2D/D2H blob copies" % (model, width, height, batch),
                      "output_finite": bool(np.isfinite(flow).all())}))


if __name__ == "__main__":
    main()
