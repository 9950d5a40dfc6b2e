#!/usr/bin/env python
"""Parity at the BASELINE.json shapes, three ways: engine (libfn2.so) vs the float64-accumulating CPU oracle vs the
REFERENCE's own layer code on the GPU (oracle/_ref).  Prints a table; tests/test_ref_gpu.py asserts on the same numbers.

    python tools/parity_fullsize.py [corr] [C] [CSS] [FN2]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import flownet2_b200 as F  # noqa: E402
from flownet2_b200 import ops  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import ref as R  # noqa: E402
from oracle.net import OracleNet, synth_weights  # noqa: E402
from util import maxabs, rng, smooth_images  # noqa: E402

RESULTS = {}


def corr_case(N, C, H, W):
    r = rng(N * 1000 + H)
    a = r.standard_normal((N, C, H, W)).astype(np.float32)
    b = r.standard_normal((N, C, H, W)).astype(np.float32)
    ta = torch.from_numpy(a).cuda().contiguous(memory_format=torch.channels_last)
    tb = torch.from_numpy(b).cuda().contiguous(memory_format=torch.channels_last)
    got = ops.correlation(ta, tb, 20, 1, 20, 1, 2).contiguous().cpu().numpy()
    t0 = time.time()
    want = O.correlation_fwd(a, b, 20, 1, 20, 1, 2, 0, exact_order=False)
    t_or = time.time() - t0
    text = ('name: "c" type: "Correlation" bottom: "a" bottom: "b" top: "t" correlation_param { pad: 20 kernel_size: 1 '
            'max_displacement: 20 stride_1: 1 stride_2: 2 }')
    ref, = R.run_layer(text, [a, b])
    res = {"engine_vs_oracle": maxabs(got, want), "engine_vs_reference_gpu": maxabs(got, ref),
           "reference_gpu_vs_oracle": maxabs(ref, want), "absmax": float(np.abs(want).max()), "oracle_s": round(t_or, 2)}
    RESULTS["corr_%dx%dx%dx%d" % (N, C, H, W)] = res
    print("corr", (N, C, H, W), res, flush=True)


def net_case(model, w, h, batch, seed=1701):
    small = F.fill_template(F.model_template(model), 64, 64)
    proto = F.fill_template(F.model_template(model), w, h)
    weights, blob = synth_weights(small, seed, proto)
    img0, img1 = smooth_images(rng(seed), batch, h, w)
    net = F.Net(proto, blob, F.TEST, batch=batch)
    got = net.forward(img0=img0, img1=img1)["predict_flow_final"].copy()
    got2 = net.forward(img0=img0, img1=img1)["predict_flow_final"].copy()
    del net
    torch.cuda.empty_cache()
    t0 = time.time()
    rnet = R.RefNet(proto, weights, batch=batch)
    rnet.forward(img0=img0, img1=img1)
    ref = rnet.blob("predict_flow_final")
    t_ref = time.time() - t0
    del rnet
    t0 = time.time()
    want = OracleNet(proto, blob, batch=batch, f64acc=True).forward(img0=img0, img1=img1)["predict_flow_final"]
    t_or = time.time() - t0
    res = {"engine_vs_oracle": maxabs(got, want), "engine_vs_reference_gpu": maxabs(got, ref),
           "reference_gpu_vs_oracle": maxabs(ref, want), "replay_equal": bool(np.array_equal(got, got2)),
           "absmax": float(np.abs(want).max()), "oracle_s": round(t_or, 1), "reference_gpu_s": round(t_ref, 1)}
    RESULTS["%s_%dx%d_b%d" % (model, w, h, batch)] = res
    print(model, (w, h, batch), res, flush=True)


def main():
    what = sys.argv[1:] or ["corr", "C", "CSS", "FN2"]
    R.set_mode(True, 0)
    if "corr" in what:
        for shp in [(8, 256, 40, 56), (4, 256, 48, 96), (4, 256, 56, 128)]:
            corr_case(*shp)
    if "C" in what:
        net_case("FlowNet2-C", 448, 320, 1)
    if "S" in what:
        net_case("FlowNet2-S", 448, 320, 1)
    if "CSS" in what:
        net_case("FlowNet2-CSS", 768, 384, 1)
    if "FN2" in what:
        net_case("FlowNet2", 1024, 436, 1)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_fullsize.json"), "w") as f:
        json.dump(RESULTS, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
