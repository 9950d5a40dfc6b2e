#!/usr/bin/env python
"""Generates tests/golden/ref_golden.npz + ref_golden.json by running the REFERENCE's own layer code (oracle/_ref,
compiled from /root/reference by oracle/ref_shim/build_ref.py) on a GPU:

    gpurun -- python tests/golden/make_ref_golden.py gpurun_out/ref_golden      (then copy the two files to tests/golden/)

Layer cases and net cases are listed in tests/golden/ref_cases.py.  For every layer case the reference layer is set up
through Layer::SetUp, run through Layer::Forward (GPU mode: its kernels / cuBLAS calls) and, where marked, Layer::Backward.
Correlation cases are run three times to answer whether the warp-synchronous reduction of CorrelateData
(correlation_layer.cu:79-105, no __syncwarp) is deterministic on sm_100; the answer is stored in the .json.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_cases as RC  # noqa: E402
from oracle import ref as R  # noqa: E402
from oracle.net import synth_weights  # noqa: E402


def run_layer_case(name, gpu=True):
    c = RC.LAYER_CASES[name]
    bottoms, params, r = RC.case_inputs(name)
    bl = [R.Blob(shape=b.shape) for b in bottoms]
    for b, a in zip(bl, bottoms):
        b.set(a)
    tl = [R.Blob() for _ in range(c["ntop"])]
    layer = R.Layer(c["text"], 1)
    layer.setup(bl, tl)
    if params is not None:
        for p, a in zip(layer.params, params):
            p.set(np.asarray(a, np.float32).reshape(p.shape))
    layer.forward()
    out = {"top%d" % i: t.get() for i, t in enumerate(tl)}
    if c["kind"] == "aug_train":
        out["space"] = layer.debug_eigenspace()
    if c.get("keep_params"):
        for i, p in enumerate(layer.params):
            out["param%d" % i] = p.get()
    if c.get("backward"):
        for t in tl:
            t.set(r.standard_normal(t.shape).astype(np.float32), diff=True)
        for b in bl:                                   # FlowWarp / Correlation accumulate or overwrite: start from zero
            b.set(np.zeros(b.shape, np.float32), diff=True)
        layer.backward()
        for i, b in enumerate(bl):
            out["bdiff%d" % i] = b.get(diff=True)
    return out


def main():
    outbase = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "ref_golden")
    import torch
    assert torch.cuda.is_available(), "run on the GPU box"
    R.set_mode(True, 0)
    g, meta = {}, {"device": torch.cuda.get_device_name(0), "layers": {}, "nets": {}}
    for name in RC.LAYER_CASES:
        t0 = time.time()
        out = run_layer_case(name)
        info = {"shapes": {k: list(v.shape) for k, v in out.items()}}
        if RC.LAYER_CASES[name]["kind"] == "correlation":
            again = [run_layer_case(name) for _ in range(2)]
            info["deterministic_over_3_runs"] = bool(all(np.array_equal(out[k], a[k], equal_nan=True) for a in again for k in out))
        for k, v in out.items():
            g["L/%s/%s" % (name, k)] = v
        info["seconds"] = round(time.time() - t0, 3)
        meta["layers"][name] = info
        print(name, info, flush=True)
    # CPU twins of the reference where Forward_cpu exists: do the reference's two paths agree?
    R.set_mode(False)
    for name, c in RC.LAYER_CASES.items():
        if c["kind"] in ("flow_warp", "channel_norm", "conv"):
            cpu = run_layer_case(name)
            d = {k: float(np.nanmax(np.abs(cpu[k].astype(np.float64) - g["L/%s/%s" % (name, k)]))) for k in cpu}
            meta["layers"][name]["reference_cpu_vs_gpu_maxabs"] = d
            print(name, "reference CPU vs GPU", d, flush=True)
    R.set_mode(True, 0)
    import flownet2_b200 as F
    for cname, (model, w, h, batch) in RC.NET_CASES.items():
        t0 = time.time()
        small = F.fill_template(F.model_template(model), 64, 64)
        proto = F.fill_template(F.model_template(model), w, h)
        weights, _ = synth_weights(small, 1701, proto)
        from util import rng, smooth_images
        img0, img1 = smooth_images(rng(11), batch, h, w)
        net = R.RefNet(proto, weights, batch=batch)
        net.forward(img0=img0, img1=img1)
        flow = net.blob("predict_flow_final")
        g["N/%s/flow" % cname] = flow
        meta["nets"][cname] = {"model": model, "w": w, "h": h, "batch": batch, "absmax": float(np.abs(flow).max()),
                               "seconds": round(time.time() - t0, 2)}
        print(cname, meta["nets"][cname], flush=True)
    np.savez_compressed(outbase + ".npz", **g)
    with open(outbase + ".json", "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", outbase + ".npz", os.path.getsize(outbase + ".npz"), "bytes")


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    main()
