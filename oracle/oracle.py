"""ctypes front end of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under flownet2_b200/ may import this module.

Every function takes/returns fp32 NCHW numpy arrays (the Caffe Blob layout) and cites the
reference lines it restates in oracle.c.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile liboracle.so (gcc, seconds).  Called by __graft_entry__.build()."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
    return _LIB


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def correlation_shape(H, W, pad, kernel_size, max_disp, stride1, stride2):
    out = (C.c_int * 5)()
    rc = lib().fn2o_correlation_shape(H, W, pad, kernel_size, max_disp, stride1, stride2, out)
    if rc:
        raise ValueError("correlation: invalid configuration (rc=%d)" % rc)
    return tuple(out)  # top_channels, top_h, top_w, grid_radius, grid_width


def correlation_fwd(b0, b1, pad, kernel_size, max_disp, stride1, stride2, corr_type=0,
                    exact_order=True):
    b0, p0 = _f(b0)
    b1, p1 = _f(b1)
    N, Cc, H, W = b0.shape
    tc, th, tw, _, _ = correlation_shape(H, W, pad, kernel_size, max_disp, stride1, stride2)
    top = np.empty((N, tc, th, tw), np.float32)
    rc = lib().fn2o_correlation_fwd(p0, p1, top.ctypes.data_as(C.POINTER(C.c_float)), N, Cc, H, W,
                                    pad, kernel_size, max_disp, stride1, stride2, corr_type,
                                    1 if exact_order else 0)
    assert rc == 0
    return top


def correlation_bwd(b0, b1, topdiff, pad, kernel_size, max_disp, stride1, stride2):
    b0, p0 = _f(b0)
    b1, p1 = _f(b1)
    td, ptd = _f(topdiff)
    N, Cc, H, W = b0.shape
    d0 = np.empty_like(b0)
    d1 = np.empty_like(b1)
    rc = lib().fn2o_correlation_bwd(p0, p1, ptd, d0.ctypes.data_as(C.POINTER(C.c_float)),
                                    d1.ctypes.data_as(C.POINTER(C.c_float)), N, Cc, H, W, pad,
                                    kernel_size, max_disp, stride1, stride2)
    assert rc == 0
    return d0, d1


def correlation1d_shape(H, W, pad, kernel_size, max_disp, stride1, stride2, single_direction=0):
    out = (C.c_int * 5)()
    rc = lib().fn2o_correlation1d_shape(H, W, pad, kernel_size, max_disp, stride1, stride2, single_direction, out)
    if rc:
        raise ValueError("correlation1d: invalid configuration (rc=%d)" % rc)
    return tuple(out)  # top_channels, top_h, top_w, grid_radius, x_shift


def correlation1d_fwd(b0, b1, pad, kernel_size, max_disp, stride1, stride2, single_direction=0, corr_type=0):
    b0, p0 = _f(b0)
    b1, p1 = _f(b1)
    N, Cc, H, W = b0.shape
    tc, th, tw, _, _ = correlation1d_shape(H, W, pad, kernel_size, max_disp, stride1, stride2, single_direction)
    top = np.empty((N, tc, th, tw), np.float32)
    rc = lib().fn2o_correlation1d_fwd(p0, p1, top.ctypes.data_as(C.POINTER(C.c_float)), N, Cc, H, W, pad, kernel_size, max_disp,
                                      stride1, stride2, single_direction, corr_type)
    assert rc == 0
    return top


def correlation_bwd_ex(b0, b1, topdiff, pad, kernel_size, max_disp, stride1, stride2, corr_type=0, one_d=False, single_direction=0):
    """Gradients of Correlation (2-D) / Correlation1D, MULTIPLY or SUBTRACT, float64 accumulation."""
    b0, p0 = _f(b0)
    b1, p1 = _f(b1)
    td, ptd = _f(topdiff)
    N, Cc, H, W = b0.shape
    d0 = np.empty_like(b0)
    d1 = np.empty_like(b1)
    rc = lib().fn2o_correlation_bwd_ex(p0, p1, ptd, d0.ctypes.data_as(C.POINTER(C.c_float)), d1.ctypes.data_as(C.POINTER(C.c_float)),
                                       N, Cc, H, W, pad, kernel_size, max_disp, stride1, stride2, corr_type, 1 if one_d else 0,
                                       single_direction)
    assert rc == 0
    return d0, d1


def flow_warp_fwd(image, flow, fill_nan=False):
    image, pi = _f(image)
    flow, pf = _f(flow)
    N, Cc, H, W = image.shape
    assert flow.shape == (N, 2, H, W)
    out = np.empty_like(image)
    lib().fn2o_flow_warp_fwd(pi, pf, out.ctypes.data_as(C.POINTER(C.c_float)), N, Cc, H, W,
                             1 if fill_nan else 0)
    return out


def flow_warp_bwd(image, flow, warped_diff):
    image, pi = _f(image)
    flow, pf = _f(flow)
    wd, pw = _f(warped_diff)
    N, Cc, H, W = image.shape
    di = np.empty_like(image)
    df = np.empty_like(flow)
    lib().fn2o_flow_warp_bwd(pi, pf, pw, di.ctypes.data_as(C.POINTER(C.c_float)),
                             df.ctypes.data_as(C.POINTER(C.c_float)), N, Cc, H, W)
    return di, df


RESAMPLE_NEAREST, RESAMPLE_LINEAR, RESAMPLE_CUBIC = 1, 2, 3


def resample_fwd(x, out_h, out_w, rtype=RESAMPLE_LINEAR, antialias=True):
    x, px = _f(x)
    N, Cc, H, W = x.shape
    out = np.empty((N, Cc, out_h, out_w), np.float32)
    rc = lib().fn2o_resample_fwd(px, out.ctypes.data_as(C.POINTER(C.c_float)), N * Cc, H, W,
                                 out_h, out_w, rtype, 1 if antialias else 0)
    assert rc == 0
    return out


def transmat_from_coeff(width, height, bottomwidth, bottomheight, mirror=0.0, angle=0.0, dx=0.0,
                        dy=0.0, zoom_x=1.0, zoom_y=1.0):
    out = (C.c_float * 6)()
    lib().fn2o_transmat_from_coeff(C.c_float(mirror), C.c_float(angle), C.c_float(dx),
                                   C.c_float(dy), C.c_float(zoom_x), C.c_float(zoom_y),
                                   width, height, bottomwidth, bottomheight, out)
    return np.array(list(out), np.float32)


def spatial_augmentation(src, mats, dest_h, dest_w):
    src, ps = _f(src)
    mats, pm = _f(mats)
    N, Cc, H, W = src.shape
    assert mats.shape == (N, 6)
    dst = np.empty((N, Cc, dest_h, dest_w), np.float32)
    lib().fn2o_spatial_augmentation(ps, dst.ctypes.data_as(C.POINTER(C.c_float)), pm, N, Cc, H, W,
                                    dest_h, dest_w)
    return dst


def color_contrast_augmentation(data, chroma, max_multiplier=1.0):
    data = np.array(data, np.float32, copy=True, order="C")
    chroma, pc = _f(chroma)
    N, Cc, H, W = data.shape
    assert Cc == 3 and chroma.shape == (N, 6)
    lib().fn2o_color_contrast_augmentation(data.ctypes.data_as(C.POINTER(C.c_float)), pc, N, H, W,
                                           C.c_float(max_multiplier))
    return data


def chromatic_eigenspace(data, eigvec9):
    """-> 25 floats (tChromaticEigenSpace: mean_eig, mean_rgb, max_abs_eig, max_rgb, min_rgb, max_l, eigvec)."""
    data, pd = _f(data)
    ev, pe = _f(np.asarray(eigvec9, np.float32).reshape(9))
    N, Cc, H, W = data.shape
    assert Cc == 3
    space = np.zeros(25, np.float32)
    lib().fn2o_chromatic_eigenspace(pd, N, H, W, pe, space.ctypes.data_as(C.POINTER(C.c_float)))
    return space


def chromatic_eigen_augmentation(data, coeffs, space, max_multiplier=1.0):
    data = np.array(data, np.float32, copy=True, order="C")
    coeffs, pc = _f(coeffs)
    space, ps = _f(space)
    N, Cc, H, W = data.shape
    assert Cc == 3 and coeffs.shape == (N, 22) and space.shape == (25,)
    lib().fn2o_chromatic_eigen_augmentation(data.ctypes.data_as(C.POINTER(C.c_float)), pc, ps, N, H, W, C.c_float(max_multiplier))
    return data


def apply_effects(data, effects, max_multiplier=1.0):
    data = np.array(data, np.float32, copy=True, order="C")
    effects, pe = _f(effects)
    N, Cc, H, W = data.shape
    assert effects.shape == (N, 9)
    lib().fn2o_apply_effects(data.ctypes.data_as(C.POINTER(C.c_float)), pe, N, Cc, H, W, C.c_float(max_multiplier))
    return data


def mean_subtract(top, mode, num_iter=0.0, recompute_mean=0, mean_per_pixel=False, mean_pp=None,
                  mean_pc=None):
    """Returns (top, mean_pp, mean_pc) after the mean step (copies; inputs untouched)."""
    top = np.array(top, np.float32, copy=True, order="C")
    N, Cc, H, W = top.shape
    mean_pp = np.zeros((Cc, H, W), np.float32) if mean_pp is None else np.array(mean_pp, np.float32, copy=True)
    mean_pc = np.zeros((Cc,), np.float32) if mean_pc is None else np.array(mean_pc, np.float32, copy=True)
    lib().fn2o_mean_subtract(top.ctypes.data_as(C.POINTER(C.c_float)), N, Cc, H, W, mode,
                             C.c_float(num_iter), recompute_mean, 1 if mean_per_pixel else 0,
                             mean_pp.ctypes.data_as(C.POINTER(C.c_float)),
                             mean_pc.ctypes.data_as(C.POINTER(C.c_float)))
    return top, mean_pp, mean_pc


def _pair(v):
    return (v, v) if np.isscalar(v) else tuple(v)


def conv_fwd(x, weight, bias=None, stride=1, pad=0, dilation=1, group=1, f64acc=False):
    x, px = _f(x)
    weight, pw = _f(weight)
    N, Ci, H, W = x.shape
    Co, Cig, kh, kw = weight.shape
    assert Cig * group == Ci
    sh, sw = _pair(stride)
    ph, pw_ = _pair(pad)
    dh, dw = _pair(dilation)
    Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    Wo = (W + 2 * pw_ - (dw * (kw - 1) + 1)) // sw + 1
    out = np.empty((N, Co, Ho, Wo), np.float32)
    pb = None
    if bias is not None:
        bias, pb = _f(bias)
    rc = lib().fn2o_conv_fwd(px, pw, pb, out.ctypes.data_as(C.POINTER(C.c_float)), N, Ci, H, W, Co,
                             kh, kw, ph, pw_, sh, sw, dh, dw, group, 1 if f64acc else 0)
    assert rc == 0, rc
    return out


def deconv_fwd(x, weight, bias=None, stride=1, pad=0, group=1, f64acc=False):
    x, px = _f(x)
    weight, pw = _f(weight)
    N, Ci, H, W = x.shape
    Ci2, Cog, kh, kw = weight.shape
    assert Ci2 == Ci
    Co = Cog * group
    sh, sw = _pair(stride)
    ph, pw_ = _pair(pad)
    Ho = sh * (H - 1) + kh - 2 * ph
    Wo = sw * (W - 1) + kw - 2 * pw_
    out = np.empty((N, Co, Ho, Wo), np.float32)
    pb = None
    if bias is not None:
        bias, pb = _f(bias)
    rc = lib().fn2o_deconv_fwd(px, pw, pb, out.ctypes.data_as(C.POINTER(C.c_float)), N, Ci, H, W,
                               Co, kh, kw, ph, pw_, sh, sw, group, 1 if f64acc else 0)
    assert rc == 0, rc
    return out


def conv_bwd(x, weight, top_diff, stride=1, pad=0, deconv=False):
    """Gradients of ConvolutionLayer / DeconvolutionLayer (conv_layer.cpp:40-73, deconv_layer.cpp:40-74 ->
    base_conv_layer.cpp:329-350 backward_cpu_gemm / weight_cpu_gemm / backward_cpu_bias), restated tap by tap with float64
    accumulation: -> (bottom_diff, weight_diff, bias_diff), group 1, no dilation.  The deconvolution is the convolution with
    the roles of bottom and top exchanged (its forward is conv's backward_cpu_gemm, its data gradient conv's forward_cpu_gemm)."""
    x = np.asarray(x, np.float64)
    w = np.asarray(weight, np.float64)
    dy = np.asarray(top_diff, np.float64)
    sh, sw = _pair(stride)
    ph, pw = _pair(pad)
    kh, kw = w.shape[2:]
    if deconv:
        small, big = x, dy            # weight [C_small][C_big][kh][kw]
    else:
        small, big = dy, x            # weight [C_small][C_big][kh][kw] with small = top
    N, Cb, H, W = big.shape
    Hs, Ws = small.shape[2:]
    bp = np.zeros((N, Cb, H + 2 * ph + sh, W + 2 * pw + sw))
    bp[:, :, ph:ph + H, pw:pw + W] = big
    dw = np.zeros_like(w)
    gbig = np.zeros_like(bp)
    for ky in range(kh):
        for kx in range(kw):
            win = (slice(None), slice(None), slice(ky, ky + sh * (Hs - 1) + 1, sh), slice(kx, kx + sw * (Ws - 1) + 1, sw))
            # small[n,a,i,j] pairs with big_padded[n,b,i*s+ky,j*s+kx] through w[a,b,ky,kx]
            dw[:, :, ky, kx] = np.einsum("naij,nbij->ab", small, bp[win])
            gbig[win] += np.einsum("naij,ab->nbij", small, w[:, :, ky, kx])
    gbig = gbig[:, :, ph:ph + H, pw:pw + W]
    if deconv:
        # bottom = small: its gradient is the convolution of the top diff with the same weights
        dx = np.zeros_like(x)
        for ky in range(kh):
            for kx in range(kw):
                win = (slice(None), slice(None), slice(ky, ky + sh * (Hs - 1) + 1, sh), slice(kx, kx + sw * (Ws - 1) + 1, sw))
                dx += np.einsum("nbij,ab->naij", bp[win], w[:, :, ky, kx])
        db = dy.sum(axis=(0, 2, 3))
        return dx.astype(np.float32), dw.astype(np.float32), db.astype(np.float32)
    db = dy.sum(axis=(0, 2, 3))
    return gbig.astype(np.float32), dw.astype(np.float32), db.astype(np.float32)


def relu_bwd(top_data, top_diff, negative_slope=0.0):
    """relu_layer.cpp:26-41 evaluated from the top data (the layer runs in place in every FlowNet prototxt)."""
    t = np.asarray(top_data, np.float32)
    d = np.asarray(top_diff, np.float32)
    return (d * ((t > 0) + np.float32(negative_slope) * (t <= 0))).astype(np.float32)


def relu(x, negative_slope=0.0):
    x, px = _f(x)
    out = np.empty_like(x)
    lib().fn2o_relu(px, out.ctypes.data_as(C.POINTER(C.c_float)), C.c_size_t(x.size),
                    C.c_float(negative_slope))
    return out


def eltwise_sum(bottoms, coeffs=None):
    bs = [np.ascontiguousarray(b, np.float32) for b in bottoms]
    coeffs = [1.0] * len(bs) if coeffs is None or len(coeffs) == 0 else list(coeffs)
    arr = (C.POINTER(C.c_float) * len(bs))(*[b.ctypes.data_as(C.POINTER(C.c_float)) for b in bs])
    cf = (C.c_float * len(bs))(*coeffs)
    top = np.empty_like(bs[0])
    lib().fn2o_eltwise_sum(arr, cf, len(bs), top.ctypes.data_as(C.POINTER(C.c_float)),
                           C.c_size_t(top.size))
    return top


def channel_norm(x):
    x, px = _f(x)
    N, Cc, H, W = x.shape
    out = np.empty((N, 1, H, W), np.float32)
    lib().fn2o_channel_norm(px, out.ctypes.data_as(C.POINTER(C.c_float)), N, Cc, H, W)
    return out
