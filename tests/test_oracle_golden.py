"""CPU tests of the oracle: against the reference's own known-answer tests (conv/deconv/ReLU), against
analytic identities for the layers the reference never tests (Correlation / FlowWarp / Resample /
DataAugmentation -- "parity unpinned"), against an independent NumPy restatement, and
against the committed golden fixtures (regression pin)."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.util import maxabs, rng

GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ---------------------------------------------------------------------------------------------------
# reference known-answer tests
# ---------------------------------------------------------------------------------------------------
def caffe_conv_naive(x, w, b, stride, pad):
    """Independent naive conv like the reference's in-test caffe_conv (test_convolution_layer.cpp:22-139)."""
    N, Ci, H, W = x.shape
    Co, _, kh, kw = w.shape
    Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    xp = np.zeros((N, Ci, H + 2 * pad, W + 2 * pad), np.float64)
    xp[:, :, pad:pad + H, pad:pad + W] = x
    out = np.zeros((N, Co, Ho, Wo), np.float64)
    for r in range(kh):
        for s in range(kw):
            patch = xp[:, :, r:r + stride * Ho:stride, s:s + stride * Wo:stride]
            out += np.einsum("nchw,oc->nohw", patch, w[:, :, r, s].astype(np.float64))
    return out + (b[None, :, None, None] if b is not None else 0)


def test_conv_vs_naive_1e4():
    # TestSimpleConvolution: 2x3x6x4 input, 4 outputs, kernel 3 stride 2, bias 0.1, EXPECT_NEAR 1e-4
    r = rng()
    x = r.standard_normal((2, 3, 6, 4)).astype(np.float32)
    w = r.standard_normal((4, 3, 3, 3)).astype(np.float32)
    b = np.full(4, 0.1, np.float32)
    assert maxabs(O.conv_fwd(x, w, b, 2, 0), caffe_conv_naive(x, w, b, 2, 0)) < 1e-4
    # Test1x1Convolution (:443)
    w1 = r.standard_normal((4, 3, 1, 1)).astype(np.float32)
    assert maxabs(O.conv_fwd(x, w1, b, 1, 0), caffe_conv_naive(x, w1, b, 1, 0)) < 1e-4
    # padded, f64 accumulation variant
    assert maxabs(O.conv_fwd(x, w, b, 1, 1, f64acc=True), caffe_conv_naive(x, w, b, 1, 1)) < 1e-5


def test_sobel_separable_identity():
    # TestSobelConvolution (test_convolution_layer.cpp:498-589): 3x3 Sobel == (3x1) o (1x3)
    x = rng(2).standard_normal((2, 1, 9, 7)).astype(np.float32)
    k = np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], np.float32)
    full = O.conv_fwd(x, k[None, None], None, 1, 0)
    col = O.conv_fwd(x, np.array([1, 2, 1], np.float32).reshape(1, 1, 3, 1), None, 1, 0)
    sep = O.conv_fwd(col, np.array([-1, 0, 1], np.float32).reshape(1, 1, 1, 3), None, 1, 0)
    assert maxabs(full, sep) < 1e-4


def test_deconv_all_ones_known_answer():
    # TestSimpleDeconvolution (test_deconvolution_layer.cpp:91-137): 3.1 / +3 / +9 on overlaps
    x = np.ones((2, 3, 6, 4), np.float32)
    out = O.deconv_fwd(x, np.ones((3, 4, 3, 3), np.float32), np.full(4, 0.1, np.float32), 2, 0)
    assert out.shape == (2, 4, 13, 9)
    for h in range(13):
        for w in range(9):
            expected = 3.1
            ho = h % 2 == 0 and 0 < h < 12
            wo = w % 2 == 0 and 0 < w < 8
            expected += 9 if (ho and wo) else (3 if (ho or wo) else 0)
            assert np.allclose(out[:, :, h, w], expected, atol=1e-4)


def test_leaky_relu():
    # test_neuron_layer.cpp:232-260
    x = rng(3).standard_normal((2, 3, 4, 5)).astype(np.float32)
    y = O.relu(x, 0.01)
    assert np.array_equal(y[x >= 0], x[x >= 0]) and np.allclose(y[x < 0], x[x < 0] * 0.01, rtol=1e-6)


# ---------------------------------------------------------------------------------------------------
# analytic identities for the unpinned layers
# ---------------------------------------------------------------------------------------------------
def correlation_numpy(a, b, pad, k, md, s1, s2):
    """Second, independent restatement (vectorised NumPy, float64) of correlation_layer.cu:46-114."""
    N, C, H, W = a.shape
    kr, gr = (k - 1) // 2, md // s2
    D = 2 * gr + 1
    border = md + kr
    pH, pW = H + 2 * pad, W + 2 * pad
    th = int(np.ceil((pH - 2 * border) / s1)); tw = int(np.ceil((pW - 2 * border) / s1))
    extra = md + k + s1 * max(th, tw)
    ap = np.zeros((N, C, pH + extra, pW + extra)); bp = np.zeros_like(ap)
    ap[:, :, pad:pad + H, pad:pad + W] = a; bp[:, :, pad:pad + H, pad:pad + W] = b
    out = np.zeros((N, D * D, th, tw))
    ys, xs = np.arange(th) * s1 + md, np.arange(tw) * s1 + md
    for tc in range(D * D):
        o, p = (tc % D - gr) * s2, (tc // D - gr) * s2
        acc = 0
        for j in range(k):
            for i in range(k):
                acc = acc + (ap[:, :, ys[:, None] + j, xs[None, :] + i] * bp[:, :, ys[:, None] + p + j, xs[None, :] + o + i]).sum(1)
        out[:, tc] = acc / (k * k * C)
    return out


@pytest.mark.parametrize("cfg", [(4, 1, 4, 1, 2), (3, 3, 2, 1, 1), (4, 3, 2, 2, 2), (2, 1, 4, 1, 2)])
def test_correlation_vs_numpy(cfg):
    pad, k, md, s1, s2 = cfg
    r = rng(4)
    a = r.standard_normal((2, 12, 14, 15)).astype(np.float32)
    b = r.standard_normal((2, 12, 14, 15)).astype(np.float32)
    want = correlation_numpy(a, b, pad, k, md, s1, s2)
    assert maxabs(O.correlation_fwd(a, b, pad, k, md, s1, s2, 0, True), want) < 2e-6
    assert maxabs(O.correlation_fwd(a, b, pad, k, md, s1, s2, 0, False), want) < 1e-6


def test_correlation_identities():
    a = rng(5).standard_normal((1, 32, 16, 18)).astype(np.float32)
    t = O.correlation_fwd(a, a, 20, 1, 20, 1, 2)
    assert t.shape == (1, 441, 16, 18)
    assert maxabs(t[0, 220], (a[0].astype(np.float64) ** 2).mean(0)) < 1e-5     # centre channel = mean(x^2)
    shape = O.correlation_shape(40, 56, 20, 1, 20, 1, 2)
    assert shape[:3] == (441, 40, 56)
    with pytest.raises(ValueError):
        O.correlation_shape(4, 4, 0, 1, 8, 1, 1)


def test_correlation_backward_finite_difference():
    # the reference's GradientChecker pattern (test_gradient_check_util.hpp:19-28)
    r = rng(6)
    a = r.standard_normal((1, 3, 7, 8)).astype(np.float32)
    b = r.standard_normal((1, 3, 7, 8)).astype(np.float32)
    cfg = (2, 1, 2, 1, 1)
    td = r.standard_normal(O.correlation_fwd(a, b, *cfg).shape).astype(np.float32)
    g0, g1 = O.correlation_bwd(a, b, td, *cfg)
    f = lambda aa, bb: float((O.correlation_fwd(aa, bb, *cfg, 0, False).astype(np.float64) * td).sum())
    eps = 1e-2
    for idx in [(0, 0, 0, 0), (0, 1, 3, 4), (0, 2, 6, 7)]:
        d = np.zeros_like(a); d[idx] = eps
        assert abs((f(a + d, b) - f(a - d, b)) / (2 * eps) - g0[idx]) < 2e-3
        assert abs((f(a, b + d) - f(a, b - d)) / (2 * eps) - g1[idx]) < 2e-3


def test_flow_warp_identities():
    r = rng(7)
    img = r.standard_normal((2, 3, 9, 11)).astype(np.float32)
    z = np.zeros((2, 2, 9, 11), np.float32)
    assert np.array_equal(O.flow_warp_fwd(img, z), img)                         # zero flow = identity
    f = z.copy(); f[:, 0] = 2.0; f[:, 1] = -1.0                                 # integer flow = pure shift
    w = O.flow_warp_fwd(img, f)
    assert np.array_equal(w[:, :, 1:, :-2], img[:, :, :-1, 2:])
    assert (w[:, :, 0, :] == 0).all() and (w[:, :, :, -2:] == 0).all()           # out of range -> fill ZERO
    assert np.isnan(O.flow_warp_fwd(img, f, True)[:, :, 0, :]).all()            # ... or NaN
    fn = z.copy(); fn[0, 0, 4, 4] = np.nan
    assert (O.flow_warp_fwd(img, fn)[0, :, 4, 4] == 0).all()                    # NaN flow -> fill


def test_flow_warp_real_data_sanity():
    d = np.load(os.path.join(GOLD, "chairs_crop.npz"))
    img0 = d["img0"].astype(np.float32).transpose(2, 0, 1)[None]
    img1 = d["img1"].astype(np.float32).transpose(2, 0, 1)[None]                 # 128x128 around the crop
    flow = np.zeros((1, 2, 128, 128), np.float32)
    flow[0, :, 32:96, 32:96] = d["flow"].transpose(2, 0, 1)
    warped = O.flow_warp_fwd(img1, flow)[:, :, 32:96, 32:96]
    err_warp = np.abs(warped - img0).mean()
    err_plain = np.abs(img1[:, :, 32:96, 32:96] - img0).mean()
    assert err_warp < 0.5 * err_plain                                            # warping by GT flow explains the motion


def test_resample_identities():
    x = rng(8).standard_normal((2, 3, 10, 12)).astype(np.float32)
    for t in (1, 2, 3):
        assert np.array_equal(O.resample_fwd(x, 10, 12, t), x)                    # same size = identity
    c = np.full((1, 1, 7, 9), 3.25, np.float32)
    for t, oh, ow in [(2, 20, 31), (3, 20, 31), (2, 3, 4), (3, 3, 4)]:
        assert maxabs(O.resample_fwd(c, oh, ow, t), np.full((1, 1, oh, ow), 3.25)) < 1e-6   # partition of unity
    # the swapped half-pixel offsets (resample_layer.cu:62-63): x uses fy, y uses fx
    ramp = np.tile(np.arange(8, dtype=np.float32), (4, 1))[None, None]
    up = O.resample_fwd(ramp, 4, 16, 2)            # fx = 0.5, fy = 1: x_in = x_out*0.5 + 1/2 - 0.5
    assert maxabs(up[0, 0, 1, 2:14], np.arange(2, 14) * 0.5) < 1e-6


def test_spatial_augmentation_edge_blend():
    # deploy "identity": dim-1.05 clamp blends the last two rows/cols 0.05/0.95 (data_augmentation_layer.cu:45-46)
    x = rng(9).uniform(0, 1, (1, 3, 6, 8)).astype(np.float32)
    m = O.transmat_from_coeff(8, 6, 8, 6)[None]
    assert np.array_equal(m[0], np.array([1, 0, 0, 1, 0, 0], np.float32))
    y = O.spatial_augmentation(x, m, 6, 8)
    assert np.array_equal(y[:, :, :-1, :-1], x[:, :, :-1, :-1])
    want_last_col = np.float32(1 - np.float32(0.95)) * x[0, :, 0, -2] + np.float32(0.95) * x[0, :, 0, -1]
    assert maxabs(y[0, :, 0, -1], want_last_col) < 1e-6


def test_mean_running_update():
    top = rng(10).uniform(0, 1, (2, 3, 4, 5)).astype(np.float32)
    out, mpp, mpc = O.mean_subtract(top, 0, num_iter=1.0, recompute_mean=10, mean_per_pixel=False)
    assert maxabs(mpp, top.mean(0)) < 1e-6 and maxabs(mpc, top.mean((0, 2, 3))) < 1e-6
    assert maxabs(out, top - mpc[None, :, None, None]) < 1e-6
    out2, mpp2, _ = O.mean_subtract(top, 0, num_iter=11.0, recompute_mean=10, mean_per_pixel=True, mean_pp=mpp, mean_pc=mpc)
    assert np.array_equal(mpp2, mpp) and maxabs(out2, top - mpp[None]) < 1e-7    # frozen after recompute_mean iterations


# ---------------------------------------------------------------------------------------------------
# committed golden fixtures (regression pin; generated by tests/golden/make_golden.py)
# ---------------------------------------------------------------------------------------------------
def test_ops_golden_fixture():
    g = np.load(os.path.join(GOLD, "ops_golden.npz"))
    a, b = g["corr_a"], g["corr_b"]
    eq = lambda x, y: maxabs(x, y) == 0.0
    assert eq(O.correlation_fwd(a, b, 4, 1, 4, 1, 2, 0), g["corr_mul_p4_k1_d4_s1_s2"])
    assert eq(O.correlation_fwd(a, b, 3, 3, 2, 2, 1, 0), g["corr_mul_p3_k3_d2_s2_s1"])
    assert eq(O.correlation_fwd(a, b, 2, 1, 2, 1, 1, 1), g["corr_sub_p2_k1_d2_s1_s1"])
    d0, d1 = O.correlation_bwd(a, b, g["corr_topdiff"], 4, 1, 4, 1, 2)
    assert eq(d0, g["corr_bwd0"]) and eq(d1, g["corr_bwd1"])
    assert eq(O.flow_warp_fwd(g["warp_img"], g["warp_flow"], False), g["warp_zero"])
    assert eq(O.flow_warp_fwd(g["warp_img"], g["warp_flow"], True), g["warp_nan"])
    bi, bf = O.flow_warp_bwd(g["warp_img"], g["warp_flow"], g["warp_topdiff"])
    assert eq(bi, g["warp_bwd_img"]) and eq(bf, g["warp_bwd_flow"])
    x = g["rs_x"]
    assert eq(O.resample_fwd(x, 29, 37, 2, True), g["rs_linear_up"])
    assert eq(O.resample_fwd(x, 5, 7, 2, True), g["rs_linear_down_aa"])
    assert eq(O.resample_fwd(x, 23, 31, 3, True), g["rs_cubic_up"])
    assert eq(O.resample_fwd(x, 20, 31, 1, True), g["rs_nearest"])
    assert eq(O.spatial_augmentation(g["aug_x"], g["aug_mats"][[0, 0]], 12, 14), g["aug_identity"])
    assert eq(O.spatial_augmentation(g["aug_x"], g["aug_mats"], 8, 10), g["aug_affine"])
    assert eq(O.conv_fwd(g["conv_x"], g["conv_w"], g["conv_b"], 2, 1), g["conv_s2_p1"])
    assert eq(O.deconv_fwd(g["conv_x"], g["deconv_w"], g["deconv_b"], 2, 1), g["deconv_s2_p1"])
    assert eq(O.channel_norm(g["conv_x"]), g["chnorm"])


def test_oracle_training_augmentations_identities():
    """Chromatic-eigen with an orthonormal basis and default coefficients is the identity (up to the [0, max] clamp); the
    shadow effect darkens exactly the half-plane (x - W/2) nx + (y - H/2) ny > distance."""
    r = np.random.default_rng(5)
    x = r.uniform(0.05, 0.95, (2, 3, 6, 9)).astype(np.float32)
    q, _ = np.linalg.qr(r.standard_normal((3, 3)))
    space = O.chromatic_eigenspace(x, q.astype(np.float32).reshape(9))
    assert np.allclose(space[3:6], x.mean(axis=(0, 2, 3)), atol=1e-6)
    ident = np.tile(np.array([1, 1, 1, 0, 0, 0, 1, 1, 1, 1, 1, 1, 0, 0, 0, 1, 1, 1, 1, 0, 1, 0], np.float32), (2, 1))
    y = O.chromatic_eigen_augmentation(x, ident, space, 1.0)
    assert np.abs(y - x).max() < 2e-6
    eff = np.zeros((2, 9), np.float32)
    eff[:, 4] = 1.0; eff[:, 6] = 1.5; eff[:, 7] = 0.25                 # normal (1, 0), distance 1.5, strength 0.25
    z = O.apply_effects(x, eff, 1.0)
    xs = np.arange(9) - 9 // 2
    dark = xs - 1.5 > 0
    assert np.array_equal(z[..., ~dark], x[..., ~dark])
    assert np.allclose(z[..., dark], np.clip(x[..., dark] - 0.25, 0, 1), atol=0)


def test_aug_golden_fixture():
    """Regression pin of the training-time colour augmentations (generated by tests/golden/make_golden.py --aug-only); libm's
    powf / cosf may differ in the last ulp between glibc versions, hence 1e-6."""
    g = np.load(os.path.join(GOLD, "aug_golden.npz"))
    x = g["x"]
    space = O.chromatic_eigenspace(x, g["eigvec"])
    assert maxabs(space, g["space"]) <= 1e-6
    assert maxabs(O.chromatic_eigen_augmentation(x, g["eigen_coeffs"], g["space"], 1.0), g["eigen_out"]) <= 1e-6
    assert maxabs(O.apply_effects(x, g["effects"], 1.0), g["effects_out"]) == 0.0
    assert maxabs(O.color_contrast_augmentation(x, g["chroma"], 1.0), g["chroma_out"]) <= 1e-6
