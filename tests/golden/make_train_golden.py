"""Generates tests/golden/train_golden.npz by running the REFERENCE's own layer classes (oracle/_ref, built from the sources
under /root/reference by oracle/ref_shim/build_ref.py) on the cases of train_cases.py, in GPU mode.

    python tests/golden/make_train_golden.py            (on the GPU box; the built oracle/_ref/libref_caffe.so travels there)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import train_cases as TC  # noqa: E402
from oracle import ref as R  # noqa: E402


def run_case(name):
    c = TC.TRAIN_CASES[name]
    bottoms, params, r = TC.train_inputs(name)
    bl = [R.Blob(shape=b.shape) for b in bottoms]
    for b, a in zip(bl, bottoms):
        b.set(a)
    tl = [R.Blob()]
    layer = R.Layer(c["text"], 1)
    layer.setup(bl, tl)
    if params is not None:
        for p, a in zip(layer.params, params):
            p.set(np.asarray(a, np.float32).reshape(p.shape))
    layer.forward()
    out = {"top0": tl[0].get().reshape(tl[0].shape if tl[0].shape else (1,))}
    k = c["kind"]
    if k == "conv_bwd":
        tl[0].set(r.standard_normal(tl[0].shape).astype(np.float32), diff=True)
        for p in layer.params:
            p.set(np.zeros(p.shape, np.float32), diff=True)
        layer.backward()
        out["bdiff0"] = bl[0].get(diff=True)
        for i, p in enumerate(layer.params):
            out["pdiff%d" % i] = p.get(diff=True)
    elif k == "l1loss":
        shp = tl[0].shape
        tl[0].set(np.full(shp, c["top_diff"], np.float32), diff=True)
        for b in bl:
            b.set(np.zeros(b.shape, np.float32), diff=True)
        layer.backward()
        for i, b in enumerate(bl):
            out["bdiff%d" % i] = b.get(diff=True)
    return out


def main():
    outbase = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "train_golden")
    import torch
    assert torch.cuda.is_available(), "run on the GPU box"
    R.set_mode(True, 0)
    g, meta = {}, {"device": torch.cuda.get_device_name(0), "cases": {}}
    for name in TC.TRAIN_CASES:
        out = run_case(name)
        for k, v in out.items():
            g["T/%s/%s" % (name, k)] = v
        meta["cases"][name] = {k: list(v.shape) for k, v in out.items()}
        print(name, meta["cases"][name], flush=True)
    # whole-net gradients: the reference's Net semantics (Split layers, loss weights) on the FlowNet2-C graph + five EPE losses
    import flownet2_b200 as F
    from oracle.net import synth_weights
    proto = TC.loss_net_proto()
    ins = TC.loss_net_inputs()
    small = F.fill_template(F.model_template("FlowNet2-C"), 64, 64)
    weights, _ = synth_weights(small, TC.LOSS_NET["seed"], F.fill_template(F.model_template("FlowNet2-C"), TC.LOSS_NET["w"], TC.LOSS_NET["h"]))
    net = R.RefNet(proto, weights, batch=TC.LOSS_NET["batch"], splits=True)
    net.forward(**ins)
    net.backward()
    for lvl in TC.LOSS_NET["weights"]:
        g["N/lossnet/loss%d" % lvl] = net.blob("flow_loss%d" % lvl).reshape(-1)
    ngrads = 0
    for name, typ, layer in net.layers:
        if typ in ("Convolution", "Deconvolution"):
            for i, pb in enumerate(layer.params):
                g["N/lossnet/grad/%s/%d" % (name, i)] = TC.grad_signature(name, i, pb.get(diff=True))
                ngrads += 1
    meta["lossnet"] = {"losses": {str(l): float(g["N/lossnet/loss%d" % l][0]) for l in TC.LOSS_NET["weights"]}, "param_blobs": ngrads}
    print("lossnet", meta["lossnet"], flush=True)
    np.savez_compressed(outbase + ".npz", **g)
    with open(outbase + ".json", "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", outbase + ".npz", os.path.getsize(outbase + ".npz"), "bytes")


if __name__ == "__main__":
    main()
