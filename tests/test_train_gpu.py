"""GPU parity of the gradient path (BASELINE config 5): ConvolutionLayer / DeconvolutionLayer::Backward_gpu
through the C-ABI against the float64 oracle (oracle.conv_bwd, itself pinned to the reference's ConvolutionLayer::Backward in
tests/golden/train_golden.npz), and Net::Backward over the FlowNet2-C graph against OracleNet.backward.

Tolerances are relative to the largest gradient entry of each tensor: the data gradient runs through the 3xTF32 tensor-core
engine (same error budget as the forward pass, 1e-5), the weight gradient is an fp32 FMA reduction over up to N*H*W terms.
"""
import numpy as np
import pytest

from oracle import oracle as O
from oracle.net import OracleNet
from tests.util import maxabs, rng, smooth_images

from flownet2_b200 import ops as OPS

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def rel(got, want):
    return maxabs(got, want) / max(float(np.abs(want).max()), 1e-30)


def dev(a, channels_last=False):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.contiguous(memory_format=torch.channels_last) if channels_last else t


# (name, N, Ci, H, W, Co, k, stride, pad, deconv, channels_last)
CONV_BWD_CASES = [
    ("c3x3_s1", 2, 16, 12, 20, 24, 3, 1, 1, False, True),
    ("c3x3_s1_wide", 1, 64, 24, 40, 64, 3, 1, 1, False, True),            # adjoint runs on the tensor-core engine
    ("c5x5_s2", 2, 8, 21, 27, 16, 5, 2, 2, False, True),                  # adjoint output smaller than the bottom
    ("c7x7_s2_img", 2, 3, 32, 48, 16, 7, 2, 3, False, False),             # first layer: plain NCHW bottom
    ("c1x1", 2, 32, 9, 11, 8, 1, 1, 0, False, True),
    ("c3x3_s2_wide", 1, 128, 16, 24, 256, 3, 2, 1, False, True),
    ("c3x3_s2_odd", 2, 32, 17, 23, 64, 3, 2, 1, False, True),             # odd bottom: no extra rows; adjoint has Co = 32
    ("c3x3_s2_even", 2, 32, 18, 24, 64, 3, 2, 1, False, True),            # even bottom: one extra row / column (out_pad)
    ("predict", 2, 96, 6, 10, 2, 3, 1, 1, False, True),                   # predict_flow: Co = 2
    ("d4x4_s2", 2, 32, 6, 8, 16, 4, 2, 1, True, True),
    ("d4x4_s2_flow", 2, 2, 6, 8, 2, 4, 2, 1, True, True),                 # upsample_flow: 2 -> 2
    ("d4x4_s2_wide", 1, 256, 8, 12, 128, 4, 2, 1, True, True),
]


@pytest.mark.parametrize("case", CONV_BWD_CASES, ids=[c[0] for c in CONV_BWD_CASES])
def test_conv_backward_matches_oracle(fn2, case):
    name, N, Ci, H, W, Co, k, s, p, deconv, cl = case
    r = rng(sum(ord(c) for c in name))
    x = r.standard_normal((N, Ci, H, W)).astype(np.float32)
    w = (r.standard_normal((Ci, Co, k, k) if deconv else (Co, Ci, k, k)) / np.sqrt(Ci * k * k)).astype(np.float32)
    if deconv:
        Ho, Wo = s * (H - 1) + k - 2 * p, s * (W - 1) + k - 2 * p
    else:
        Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    dy = r.standard_normal((N, Co, Ho, Wo)).astype(np.float32)
    gx, gw, gb = OPS.conv2d_backward(dev(x, cl), dev(w), dev(dy, cl), s, p, deconv)
    torch.cuda.synchronize()
    wx, ww, wb = O.conv_bwd(x, w, dy, s, p, deconv)
    assert rel(gx.cpu().numpy(), wx) <= 1e-5, ("bottom diff", rel(gx.cpu().numpy(), wx))
    assert rel(gw.cpu().numpy(), ww) <= 1e-5, ("weight diff", rel(gw.cpu().numpy(), ww))
    assert rel(gb.cpu().numpy(), wb) <= 1e-5, ("bias diff", rel(gb.cpu().numpy(), wb))


def test_conv_backward_params_accumulates(fn2):
    import ctypes as C
    r = rng(5)
    x = r.standard_normal((2, 8, 10, 12)).astype(np.float32)
    w = r.standard_normal((16, 8, 3, 3)).astype(np.float32)
    dy = r.standard_normal((2, 16, 10, 12)).astype(np.float32)
    _, gw, gb = OPS.conv2d_backward(dev(x, True), dev(w), dev(dy, True), 1, 1, False, need_input_grad=False)
    l = fn2.lib()
    d = OPS._conv_desc(dev(w), 1, 1, False, True)
    nb = C.c_size_t()
    fn2.check(l.fn2_conv_backward_params_workspace_bytes(C.byref(d), 2, 10, 12, C.byref(nb)))
    ws = torch.empty(max(nb.value, 4), dtype=torch.uint8, device="cuda")
    gw2, gb2 = gw.clone(), gb.clone()
    xx, dd = dev(x, True), dev(dy, True)
    tx, tdy = OPS.desc(xx), OPS.desc(dd)
    fn2.check(l.fn2_conv_backward_params(C.byref(d), C.byref(tx), C.byref(tdy), C.c_void_p(gw2.data_ptr()), C.c_void_p(gb2.data_ptr()),
                                         1, C.c_void_p(ws.data_ptr()), nb.value, None))
    torch.cuda.synchronize()
    assert rel(gw2.cpu().numpy(), 2 * gw.cpu().numpy()) <= 1e-6
    assert rel(gb2.cpu().numpy(), 2 * gb.cpu().numpy()) <= 1e-6


@pytest.mark.parametrize("slope", [0.0, 0.1])
def test_relu_backward_and_axpby(fn2, slope):
    r = rng(9)
    y = r.standard_normal((2, 5, 7, 9)).astype(np.float32)
    y[0, 0, 0, :3] = 0.0
    d = r.standard_normal(y.shape).astype(np.float32)
    got = OPS.relu_backward(dev(y, True), dev(d, True), slope).cpu().numpy()
    assert np.array_equal(got, O.relu_bwd(y, d, slope))
    a, b = dev(y, True), dev(d)
    want = (np.float32(0.5) * y + np.float32(2.0) * d).astype(np.float32)
    OPS.axpby(a, 0.5, b, 2.0)
    assert maxabs(b.cpu().numpy(), want) <= 1e-6


# ---- whole-net gradients ------------------------------------------------------------------------------------------------
SEEDS = ["predict_flow6", "predict_flow5", "predict_flow4", "predict_flow3", "predict_flow2"]


def test_flownet_c_backward_matches_oracle(fn2):
    w, h, batch = 192, 100, 1
    proto = fn2.fill_template(fn2.model_template("FlowNet2-C"), w, h)
    net = fn2.Net(proto, None, fn2.TEST, batch=batch)
    net.fill_params(7)
    weights = net.to_caffemodel()
    img0, img1 = smooth_images(rng(7), batch, h, w)
    net.forward(img0=img0, img1=img1)
    r = rng(8)
    seeds = {s: r.standard_normal(net.blobs[s].shape).astype(np.float32) for s in SEEDS}
    net.clear_param_diffs()
    net.backward(**seeds)
    need = dict(zip(net.layer_names, net.layer_need_backward()))
    assert need["conv1"] and need["corr"] and need["Concat2"] and not need["Resample_final"] and not need["img0s_aug"]
    onet = OracleNet(proto, weights, batch=batch, f64acc=True)
    B = onet.forward(img0=img0, img1=img1)
    # differentiate at the ENGINE's activations: a pre-activation within rounding of zero (conv3b has some: 1e-8 against 8e-8)
    # must not pick different sides of the ReLU kink in the two implementations
    for k in list(B):
        if k not in ("img0", "img1"):
            assert maxabs(net.blobs[k].data, B[k]) <= 1e-4 * max(1.0, float(np.abs(B[k]).max())), k
            B[k] = net.blobs[k].data
    D, P = onet.backward(**seeds)
    worst = {}
    for lname, grads in P.items():
        for i, g in enumerate(grads):
            got = net.param(lname, i, diff=True).reshape(g.shape)
            worst["%s[%d]" % (lname, i)] = rel(got, g)
    bad = {k: v for k, v in worst.items() if v > 2e-5}
    assert not bad, bad
    assert len(worst) >= 40
    # (concat tops are not compared: the fused ReLUs of their zero-copy children differentiate in place inside them)
    for blob in ("conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b", "corr", "conv_redir", "conv3_1", "deconv5", "conv6_1"):
        assert rel(net.get_diff(blob), D[blob]) <= 2e-5, (blob, rel(net.get_diff(blob), D[blob]))
    # parameter gradients accumulate across calls until cleared (net.cpp:935-955)
    first = net.param("conv3_1", 0, diff=True)
    net.forward(img0=img0, img1=img1)
    net.backward(**seeds)
    assert rel(net.param("conv3_1", 0, diff=True), 2 * first) <= 1e-6
    net.clear_param_diffs()
    net.sync()
    assert not net.param("conv3_1", 0, diff=True).any()


def test_backward_without_fusion_or_aliasing_is_the_same(fn2, monkeypatch):
    """The fused conv+ReLU derivative and the zero-copy concat gradient slices against the plain layer-by-layer execution."""
    w, h, batch = 128, 64, 2
    proto = fn2.fill_template(fn2.model_template("FlowNet2-C"), w, h)
    img0, img1 = smooth_images(rng(3), batch, h, w)
    r = rng(4)
    res = []
    for plain in (False, True):
        if plain:
            monkeypatch.setenv("FN2_NO_FUSE", "1")
            monkeypatch.setenv("FN2_NO_ALIAS", "1")
        net = fn2.Net(proto, None, fn2.TEST, batch=batch)
        net.fill_params(5)
        net.forward(img0=img0, img1=img1)
        if not res:
            seeds = {s: r.standard_normal(net.blobs[s].shape).astype(np.float32) for s in SEEDS}
        net.clear_param_diffs()
        net.backward(**seeds)
        res.append({n: net.param(n, 0, diff=True) for n, t in zip(net.layer_names, net.layer_types) if t in ("Convolution", "Deconvolution")
                    and net.layer_need_backward()[net.layer_names.index(n)]})
    assert len(res[0]) >= 20
    for k in res[0]:
        assert rel(res[0][k], res[1][k]) <= 1e-5, (k, rel(res[0][k], res[1][k]))
