#!/usr/bin/env python
"""Regenerates the committed golden fixtures (run in the build container, where /root/reference exists):

    PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION=python python tests/golden/make_golden.py

Writes
  ops_golden.npz            seeded inputs + CPU-oracle outputs for every hot-path op (pins the oracle and
                            feeds the GPU parity tests with size-independent fixtures)
  ref_pb2_caffemodel.bin    a .caffemodel serialised by the REFERENCE's own python/caffe/proto/caffe_pb2.py
  ref_pb2_caffemodel.json   what it contains (names, types, shapes, sums) -> pins both wire-format readers
  ref_pb2_prototxt.json     the reference's text_format parse of our deploy templates (layer name/type/
                            bottoms/tops + a few params) -> pins the C++ and the oracle prototxt parsers
  chairs_crop.npz           64x64 crop of data/FlyingChairs_examples/0000000-{img0,img1}.ppm + gt.flo, the only
                            real-image fixture in the reference; plus header/stat of the full .flo file
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402


def ops_golden():
    r = np.random.default_rng(1701)
    g = {}
    a = r.standard_normal((2, 16, 9, 11)).astype(np.float32)
    b = r.standard_normal((2, 16, 9, 11)).astype(np.float32)
    g["corr_a"], g["corr_b"] = a, b
    g["corr_mul_p4_k1_d4_s1_s2"] = O.correlation_fwd(a, b, 4, 1, 4, 1, 2, 0)
    g["corr_mul_p3_k3_d2_s2_s1"] = O.correlation_fwd(a, b, 3, 3, 2, 2, 1, 0)
    g["corr_sub_p2_k1_d2_s1_s1"] = O.correlation_fwd(a, b, 2, 1, 2, 1, 1, 1)
    td = r.standard_normal(g["corr_mul_p4_k1_d4_s1_s2"].shape).astype(np.float32)
    g["corr_topdiff"] = td
    g["corr_bwd0"], g["corr_bwd1"] = O.correlation_bwd(a, b, td, 4, 1, 4, 1, 2)
    img = r.standard_normal((2, 3, 13, 17)).astype(np.float32)
    flow = r.uniform(-6, 6, (2, 2, 13, 17)).astype(np.float32)
    flow[0, :, 2, 3] = 100.0
    g["warp_img"], g["warp_flow"] = img, flow
    g["warp_zero"] = O.flow_warp_fwd(img, flow, False)
    g["warp_nan"] = O.flow_warp_fwd(img, flow, True)
    wd = r.standard_normal(img.shape).astype(np.float32)
    g["warp_topdiff"] = wd
    g["warp_bwd_img"], g["warp_bwd_flow"] = O.flow_warp_bwd(img, flow, wd)
    x = r.standard_normal((1, 2, 11, 15)).astype(np.float32)
    g["rs_x"] = x
    g["rs_linear_up"] = O.resample_fwd(x, 29, 37, 2, True)
    g["rs_linear_down_aa"] = O.resample_fwd(x, 5, 7, 2, True)
    g["rs_cubic_up"] = O.resample_fwd(x, 23, 31, 3, True)
    g["rs_nearest"] = O.resample_fwd(x, 20, 31, 1, True)
    im = r.uniform(0, 1, (2, 3, 12, 14)).astype(np.float32)
    mats = np.stack([O.transmat_from_coeff(14, 12, 14, 12),
                     O.transmat_from_coeff(10, 8, 14, 12, mirror=1, angle=0.3, dx=0.1, dy=-0.05, zoom_x=1.3, zoom_y=0.8)])
    g["aug_x"], g["aug_mats"] = im, mats
    g["aug_identity"] = O.spatial_augmentation(im, mats[[0, 0]], 12, 14)
    g["aug_affine"] = O.spatial_augmentation(im, mats, 8, 10)
    top, mpp, mpc = O.mean_subtract(im, 0, 3.0, 1000, False, mean_pp=np.full((3, 12, 14), 0.4, np.float32))
    g["mean_top"], g["mean_pp"], g["mean_pc"] = top, mpp, mpc
    cx = r.standard_normal((2, 5, 9, 10)).astype(np.float32)
    cw = r.standard_normal((7, 5, 3, 3)).astype(np.float32)
    cb = r.standard_normal(7).astype(np.float32)
    g["conv_x"], g["conv_w"], g["conv_b"] = cx, cw, cb
    g["conv_s2_p1"] = O.conv_fwd(cx, cw, cb, 2, 1)
    dw = r.standard_normal((5, 4, 4, 4)).astype(np.float32)
    db = r.standard_normal(4).astype(np.float32)
    g["deconv_w"], g["deconv_b"] = dw, db
    g["deconv_s2_p1"] = O.deconv_fwd(cx, dw, db, 2, 1)
    g["chnorm"] = O.channel_norm(cx)
    np.savez_compressed(os.path.join(HERE, "ops_golden.npz"), **g)
    print("ops_golden.npz:", len(g), "arrays")


def aug_golden():
    """Training-time colour augmentations of DataAugmentation (oracle restatement; the reference has no vectors for them)."""
    r = np.random.default_rng(4242)
    g = {}
    x = r.uniform(0, 1, (2, 3, 9, 13)).astype(np.float32)
    ev = np.array([0.51, 0.56, 0.65, 0.79, 0.01, -0.62, 0.35, -0.83, 0.44], np.float32)
    co = np.tile(np.array([1, 1, 1, 0, 0, 0, 1, 1, 1, 1, 1, 1, 0, 0, 0, 1, 1, 1, 1, 0, 1, 0], np.float32), (2, 1))
    co += r.uniform(-0.25, 0.25, co.shape).astype(np.float32)
    eff = np.zeros((2, 9), np.float32)
    eff[:, 4], eff[:, 5] = np.cos([0.7, 2.9]), np.sin([0.7, 2.9])
    eff[:, 6], eff[:, 7] = [1.0, -2.0], [0.3, 0.15]
    g["x"], g["eigvec"], g["eigen_coeffs"], g["effects"] = x, ev, co, eff
    g["space"] = O.chromatic_eigenspace(x, ev)
    g["eigen_out"] = O.chromatic_eigen_augmentation(x, co, g["space"], 1.0)
    g["effects_out"] = O.apply_effects(x, eff, 1.0)
    chroma = np.array([[1.2, 0.05, 0.9, 1.1, 0.95, 1.0], [0.8, -0.03, 1.15, 0.9, 1.05, 1.1]], np.float32)
    g["chroma"] = chroma
    g["chroma_out"] = O.color_contrast_augmentation(x, chroma, 1.0)
    np.savez_compressed(os.path.join(HERE, "aug_golden.npz"), **g)
    print("aug_golden.npz:", len(g), "arrays")


def ref_pb2():
    sys.path.insert(0, "/root/reference/python/caffe/proto")
    import caffe_pb2
    from google.protobuf import text_format
    r = np.random.default_rng(7)
    net = caffe_pb2.NetParameter()
    net.name = "golden"
    desc = []

    def add(name, type_, blobs):
        l = net.layer.add()
        l.name, l.type = name, type_
        d = {"name": name, "type": type_, "blobs": []}
        for kind, shape in blobs:
            arr = r.standard_normal(shape).astype(np.float32)
            bp = l.blobs.add()
            if kind == "shape":
                bp.shape.dim.extend(shape)
                bp.data.extend(arr.ravel().tolist())
            elif kind == "legacy":
                s4 = [1] * (4 - len(shape)) + list(shape)
                bp.num, bp.channels, bp.height, bp.width = s4
                bp.data.extend(arr.ravel().tolist())
                shape = s4
            elif kind == "double":
                bp.shape.dim.extend(shape)
                bp.double_data.extend(arr.ravel().astype(np.float64).tolist())
            d["blobs"].append({"shape": list(shape), "sum": float(arr.astype(np.float64).sum()),
                               "first": float(arr.ravel()[0])})
        desc.append(d)

    add("conv1", "Convolution", [("shape", (4, 3, 3, 3)), ("shape", (4,))])
    add("legacy_conv", "Convolution", [("legacy", (2, 3, 1, 1)), ("legacy", (2,))])
    add("double_ip", "InnerProduct", [("double", (3, 5))])
    add("img0s_aug", "DataAugmentation", [("shape", (1, 1, 1, 1)), ("shape", (1, 3, 4, 5)), ("shape", (1, 3, 1, 1))])
    add("no_blobs", "ReLU", [])
    with open(os.path.join(HERE, "ref_pb2_caffemodel.bin"), "wb") as f:
        f.write(net.SerializeToString())
    with open(os.path.join(HERE, "ref_pb2_caffemodel.json"), "w") as f:
        json.dump(desc, f, indent=1)

    import flownet2_b200 as F
    out = {}
    for m in ["FlowNet2-S", "FlowNet2-C", "FlowNet2-CSS", "FlowNet2-SD", "FlowNet2"]:
        txt = F.fill_template(F.model_template(m), 1024, 436)
        np_ = caffe_pb2.NetParameter()
        text_format.Merge(txt, np_)
        layers = []
        for l in np_.layer:
            e = {"name": l.name, "type": l.type, "bottom": list(l.bottom), "top": list(l.top)}
            if l.type in ("Convolution", "Deconvolution"):
                cp = l.convolution_param
                e["conv"] = [cp.num_output, list(cp.kernel_size), list(cp.stride), list(cp.pad), cp.bias_term]
            if l.type == "Correlation":
                cp = l.correlation_param
                e["corr"] = [cp.pad, cp.kernel_size, cp.max_displacement, cp.stride_1, cp.stride_2, cp.correlation_type]
            if l.type == "Resample":
                rp = l.resample_param
                e["resample"] = [rp.width, rp.height, rp.type, rp.antialias]
            if l.type == "Eltwise":
                e["coeff"] = [float(c) for c in l.eltwise_param.coeff]
            layers.append(e)
        out[m] = {"inputs": list(np_.input), "input_shape": [list(s.dim) for s in np_.input_shape], "layers": layers}
    with open(os.path.join(HERE, "ref_pb2_prototxt.json"), "w") as f:
        json.dump(out, f)
    print("ref_pb2 fixtures written")


def chairs():
    base = "/root/reference/data/FlyingChairs_examples/0000000"

    def ppm(path):
        with open(path, "rb") as f:
            assert f.readline().strip() == b"P6"
            line = f.readline()
            while line.startswith(b"#"):
                line = f.readline()
            w, h = [int(v) for v in line.split()]
            assert int(f.readline()) == 255
            return np.frombuffer(f.read(w * h * 3), np.uint8).reshape(h, w, 3)

    raw = open(base + "-gt.flo", "rb").read()
    assert raw[:4] == b"PIEH"
    w, h = np.frombuffer(raw[4:12], np.int32)
    flow = np.frombuffer(raw[12:], np.float32).reshape(h, w, 2)
    i0, i1 = ppm(base + "-img0.ppm"), ppm(base + "-img1.ppm")
    y0, x0 = 160, 224
    np.savez_compressed(os.path.join(HERE, "chairs_crop.npz"),
                        img0=i0[y0:y0 + 64, x0:x0 + 64], img1=i1[y0 - 32:y0 + 96, x0 - 32:x0 + 96],
                        flow=flow[y0:y0 + 64, x0:x0 + 64].copy(),
                        flo_header=np.array([w, h], np.int32), flo_sum=np.float64(flow.astype(np.float64).sum()),
                        flo_first_row=flow[0, :8].copy(), flo_bytes_head=np.frombuffer(raw[:12 + 8 * 8 * 2 // 2], np.uint8))
    print("chairs_crop.npz written; flow range", flow.min(), flow.max())


if __name__ == "__main__":
    import sys
    if "--aug-only" in sys.argv:
        aug_golden()
        sys.exit(0)
    ops_golden()
    aug_golden()
    ref_pb2()
    chairs()
