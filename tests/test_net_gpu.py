"""GPU parity: whole-net forward through the Caffe-surface engine (fn2_net_* C-ABI) vs the CPU
oracle executing the same prototxt with the same weights (exported as .caffemodel bytes and
re-read by the oracle's own wire-format parser).

Tolerance on predict_flow_final: 1e-4 max-abs (north_star).  Intermediate blobs are compared
relative to their magnitude (1e-4) to localise a failure.
"""
import numpy as np
import pytest

from oracle.net import OracleNet
from tests.util import maxabs, rng, smooth_images

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def run_pair(fn2, model, width, height, batch, seed=1701, check_blobs=()):
    proto = fn2.fill_template(fn2.model_template(model), width, height)
    net = fn2.Net(proto, None, fn2.TEST, batch=batch)
    net.fill_params(seed)
    weights = net.to_caffemodel()
    img0, img1 = smooth_images(rng(seed), batch, height, width)
    out = net.forward(img0=img0, img1=img1)
    got = out["predict_flow_final"]
    assert got.shape == (batch, 2, height, width)
    onet = OracleNet(proto, weights, batch=batch, f64acc=True)
    ref = onet.forward(img0=img0, img1=img1)
    want = ref["predict_flow_final"]
    report = []
    for name in check_blobs:
        g, w = net.blobs[name].data, ref[name]
        report.append((name, maxabs(g, w), float(np.abs(w).max())))
    return net, got, want, report


@pytest.mark.parametrize("model,w,h,batch", [
    ("FlowNet2-S", 128, 96, 2),          # 96 rows -> adapted 128: exercises both Resample directions
    ("FlowNet2-C", 128, 128, 1),
    ("FlowNet2-C", 192, 100, 2),
])
def test_single_net_parity(fn2, model, w, h, batch):
    blobs = ["img0_nomean_resize", "conv3_1", "conv6_1", "predict_flow6", "predict_flow2"]
    net, got, want, report = run_pair(fn2, model, w, h, batch, check_blobs=blobs)
    for name, err, mag in report:
        assert err <= 1e-4 * max(1.0, mag), (name, err, mag)
    assert np.isfinite(got).all()
    assert np.abs(want).max() > 1e-3, "degenerate test: flow is ~0"
    assert maxabs(got, want) <= 1e-4, maxabs(got, want)


def test_css_parity(fn2):
    blobs = ["net1_predict_flow2", "net2_in_img1_warped", "net2_in_err_norm", "net2_predict_flow2", "net3_predict_flow2"]
    net, got, want, report = run_pair(fn2, "FlowNet2-CSS", 128, 128, 1, check_blobs=blobs)
    for name, err, mag in report:
        assert err <= 2e-4 * max(1.0, mag), (name, err, mag)
    assert maxabs(got, want) <= 1e-4 * max(1.0, np.abs(want).max()), (maxabs(got, want), np.abs(want).max())


def test_full_flownet2_parity(fn2):
    net, got, want, report = run_pair(fn2, "FlowNet2", 128, 64, 1, check_blobs=["fuse_input", "fuse_predict_flow0"])
    for name, err, mag in report:
        assert err <= 2e-4 * max(1.0, mag), (name, err, mag)
    assert maxabs(got, want) <= 1e-4 * max(1.0, np.abs(want).max()), (maxabs(got, want), np.abs(want).max())


def test_graph_replay_matches_eager(fn2):
    proto = fn2.fill_template(fn2.model_template("FlowNet2-C"), 128, 64)
    net = fn2.Net(proto, None, fn2.TEST, batch=2)
    net.fill_params(3)
    img0, img1 = smooth_images(rng(3), 2, 64, 128)
    first = net.forward(img0=img0, img1=img1)["predict_flow_final"]        # eager (allocations)
    second = net.forward(img0=img0, img1=img1)["predict_flow_final"]       # captured + replayed
    third = net.forward(img0=img0, img1=img1)["predict_flow_final"]        # replayed
    assert maxabs(first, second) == 0.0 and maxabs(first, third) == 0.0
    assert net.launches_per_forward > 0
    other0, other1 = smooth_images(rng(4), 2, 64, 128)
    changed = net.forward(img0=other0, img1=other1)["predict_flow_final"]
    assert maxabs(first, changed) > 0.0


def test_caffemodel_roundtrip_and_name_matching(fn2):
    proto = fn2.fill_template(fn2.model_template("FlowNet2-S"), 64, 64)
    a = fn2.Net(proto, None, fn2.TEST)
    a.fill_params(5)
    blob = a.to_caffemodel()
    b = fn2.Net(proto, blob, fn2.TEST)
    # conv weights survive the round trip bit for bit; DataAugmentation only takes the iteration count and
    # the per-channel mean when mean_per_pixel is false (adjust_blobs, data_augmentation_layer.cpp:172-183)
    from oracle.net import parse_caffemodel
    wa, wb = parse_caffemodel(blob), parse_caffemodel(b.to_caffemodel())
    assert wa.keys() == wb.keys()
    for k in wa:
        idx = [0, 2] if k.endswith("_aug") else range(len(wa[k]))
        for i in idx:
            assert np.array_equal(wa[k][i], wb[k][i]), (k, i)
    img0, img1 = smooth_images(rng(5), 1, 64, 64)
    fa = a.forward(img0=img0, img1=img1)["predict_flow_final"]
    fb = b.forward(img0=img0, img1=img1)["predict_flow_final"]
    assert maxabs(fa, fb) == 0.0


def test_flo_roundtrip_and_reference_files(fn2, tmp_path):
    flow = rng(6).standard_normal((2, 13, 17)).astype(np.float32)
    p = str(tmp_path / "x.flo")
    fn2.write_flo(p, flow)
    raw = open(p, "rb").read()
    assert raw[:4] == b"PIEH" and np.frombuffer(raw[4:12], np.int32).tolist() == [17, 13]
    # python twin of the reference writer: (H,W,2) interleaved (scripts/run-flownet.py:117-124)
    assert np.array_equal(np.frombuffer(raw[12:], np.float32).reshape(13, 17, 2), flow.transpose(1, 2, 0))
    assert np.array_equal(fn2.read_flo(p), flow)
