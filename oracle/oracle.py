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


# ---- training-side neighbours (SURVEY.md 8 "next" row 2) --------------------------------------------------------------------
def l1loss_fwd(b0, b1=None, l2_per_location=False, l2_prescale_by_channels=False, normalize_by_num_entries=False,
               epsilon=1e-2, plateau=0.0):
    """L1LossLayer::Forward_gpu (l1loss_layer.cu:67-143; sub-layers set up in l1loss_layer.cpp:14-61): float32 element
    operations in the reference's order, the two dot products accumulated in float64.  -> (loss, normalize_coeff)."""
    f = np.float32
    d = np.asarray(b0, f) if b1 is None else (np.asarray(b0, f) - np.asarray(b1, f))
    ok = d == d                                                        # FindNotNaNs
    N, Cc = d.shape[0], d.shape[1]
    norm = f(ok.sum(dtype=np.float64)) / f(Cc) if normalize_by_num_entries else f(N)
    if l2_per_location:
        dz = np.where(ok, d, f(0))                                     # KillMasked
        w = f(1) / f(Cc) if l2_prescale_by_channels else f(1)
        ssum = np.zeros((N,) + d.shape[2:], f)
        for c in range(Cc):                                            # Power(2), then the 1x1 sum convolution
            ssum = (ssum + w * (dz[:, c] * dz[:, c])).astype(f)
        if plateau > 0:
            ssum = np.where(np.abs(ssum) < f(plateau) * f(plateau), f(0), ssum)
        root = np.sqrt((ssum + f(epsilon)).astype(f)).astype(f)        # Power(0.5, shift = epsilon)
        dot = root.sum(dtype=np.float64)
    else:
        keep = ok & ~(np.abs(d) < f(plateau)) if plateau > 0 else ok   # MaskPlateauValues
        dot = np.abs(np.where(keep, d, f(0))).sum(dtype=np.float64)
    return f(f(dot) / norm), norm


def l1loss_bwd(b0, b1=None, top_diff=1.0, l2_per_location=False, l2_prescale_by_channels=False, normalize_by_num_entries=False,
               epsilon=1e-2, plateau=0.0):
    """L1LossLayer::Backward_gpu (l1loss_layer.cu:146-187; Power backward power_layer.cu:40-83) -> (bottom0 diff, bottom1 diff)."""
    f = np.float32
    d = np.asarray(b0, f) if b1 is None else (np.asarray(b0, f) - np.asarray(b1, f))
    ok = d == d
    N, Cc = d.shape[0], d.shape[1]
    norm = f(ok.sum(dtype=np.float64)) / f(Cc) if normalize_by_num_entries else f(N)
    alpha = f(top_diff) / norm
    if l2_per_location:
        dz = np.where(ok, d, f(0))
        w = f(1) / f(Cc) if l2_prescale_by_channels else f(1)
        ssum = np.zeros((N,) + d.shape[2:], f)
        for c in range(Cc):
            ssum = (ssum + w * (dz[:, c] * dz[:, c])).astype(f)
        kill = (np.abs(ssum) < f(plateau) * f(plateau)) if plateau > 0 else np.zeros(ssum.shape, bool)
        ssum = np.where(kill, f(0), ssum)
        base = (ssum + f(epsilon)).astype(f)
        with np.errstate(divide="ignore", invalid="ignore"):
            gs = ((np.sqrt(base).astype(f) / base).astype(f) * f(0.5)).astype(f) * alpha
        gs = np.where(kill, f(0), gs).astype(f)
        g = ((f(2) * dz) * (w * gs)[:, None]).astype(f)
        g = np.where(ok, g, f(0))
    else:
        keep = ok & ~(np.abs(d) < f(plateau)) if plateau > 0 else ok
        dz = np.where(keep, d, f(0))
        g = np.where(keep, alpha * np.where(dz > 0, f(1), f(-1)), f(0)).astype(f)
    return g.astype(f), (-g).astype(f)


def downsample_fwd(x, top_h, top_w):
    """DownsampleFeatures (downsample_layer.cu:15-80), float32 like the kernel."""
    f = np.float32
    x = np.asarray(x, f)
    N, Cc, H, W = x.shape
    if (H, W) == (top_h, top_w):
        return x.copy()
    wscale, hscale = f(W - 1) / f(top_w - 1), f(H - 1) / f(top_h - 1)
    wrad, hrad = int(np.ceil(wscale)), int(np.ceil(hscale))
    bx = (np.arange(top_w, dtype=f) / f(top_w - 1)) * f(W - 1)
    by = (np.arange(top_h, dtype=f) / f(top_h - 1)) * f(H - 1)
    rnd = lambda v: np.where(v >= 0, np.floor(v + f(0.5)), np.ceil(v - f(0.5))).astype(np.int64)      # roundf
    ix, iy = rnd(bx), rnd(by)
    val = np.zeros((N, Cc, top_h, top_w), f)
    wsum = np.zeros((top_h, top_w), f)
    wsum_n = np.zeros((N, Cc, top_h, top_w), f)
    wnan = np.zeros((N, Cc, top_h, top_w), f)
    for yo in range(-hrad, hrad + 1):
        sy = iy + yo
        wy = np.maximum(f(0), f(1) - np.abs(sy.astype(f) - by) / hscale).astype(f)
        vy = (sy >= 0) & (sy < H)
        for xo in range(-wrad, wrad + 1):
            sx = ix + xo
            wx = np.maximum(f(0), f(1) - np.abs(sx.astype(f) - bx) / wscale).astype(f)
            vx = (sx >= 0) & (sx < W)
            valid = vy[:, None] & vx[None, :]
            wgt = np.where(valid, (wx[None, :] * wy[:, None]).astype(f), f(0))
            samp = x[:, :, np.clip(sy, 0, H - 1)[:, None], np.clip(sx, 0, W - 1)[None, :]]
            isn = np.isnan(samp)
            wnan = (wnan + np.where(isn, wgt, f(0))).astype(f)
            wg = np.where(isn, f(0), wgt).astype(f)
            val = (val + np.where(isn, f(0), samp) * wg).astype(f)
            wsum_n = (wsum_n + wg).astype(f)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = np.where(wnan / wsum_n > f(0.5), f(np.nan), val / wsum_n)
    return out.astype(f)


AUG_COEFF_DEFAULT = np.array([0, 0, 0, 0, 1, 1, 1, 0, 1, 1, 1, 1] + [1, 1, 1, 0, 0, 0, 1, 1, 1, 1, 1, 1, 0, 0, 0, 1, 1, 1, 1, 0, 1, 0] + [0] * 8,
                             np.float32)     # AugmentationCoeff, caffe.proto:436-486, in descriptor order


def aug_array_to_values(arr):
    """array_to_coeff (augmentation_layer_base.cpp:368-379): fields with default 0 as is, the others through exp."""
    arr = np.asarray(arr, np.float32).reshape(-1, 42)
    return np.where(np.abs(AUG_COEFF_DEFAULT) < 1e-3, arr, np.exp(arr).astype(np.float32)).astype(np.float32)


def transmat_inverse(m):
    """tTransMat::inverse (augmentation_layer_base.cpp:51-68) on the NAME-ordered record t0..t5, float32."""
    f = np.float32
    a, b, c, d, e, ff = [f(v) for v in (m[0], m[1], m[2], m[3], m[4], m[5])]
    den = f(a * d - b * c)
    return np.array([d / den, -b / den, -c / den, a / den, f(c * ff - d * e) / den, f(b * e - a * ff) / den], f)


def flow_augmentation(flow, params1, params2, crop_w, crop_h):
    """FlowAugmentationLayer::Forward_gpu (flow_augmentation_layer.cu:24-66, :121-166): matrices from the two coefficient blobs
    (the second inverted), then per output pixel  p1 = M1 p,  p2 = p1 + flow[nearest(p1)],  p3 = M2^-1 p2,  out = p3 - p."""
    f = np.float32
    flow = np.asarray(flow, f)
    N, _, H, W = flow.shape
    v1, v2 = aug_array_to_values(params1), aug_array_to_values(params2)
    out = np.empty((N, 2, crop_h, crop_w), f)
    flat = flow.reshape(-1)
    xs, ys = np.meshgrid(np.arange(crop_w, dtype=f), np.arange(crop_h, dtype=f))
    for n in range(N):
        a = transmat_from_coeff(crop_w, crop_h, W, H, *[float(v1[n, i]) for i in (0, 3, 1, 2, 4, 5)])
        b = transmat_inverse(transmat_from_coeff(crop_w, crop_h, W, H, *[float(v2[n, i]) for i in (0, 3, 1, 2, 4, 5)]))
        x1 = ((xs * a[0]).astype(f) + (ys * a[2]).astype(f)).astype(f) + a[4]
        y1 = ((xs * a[1]).astype(f) + (ys * a[3]).astype(f)).astype(f) + a[5]
        sx, sy = (x1 + f(0.5)).astype(np.int64), (y1 + f(0.5)).astype(np.int64)       # C cast: toward zero
        iu = np.clip(W * (H * (2 * n + 0) + sy) + sx, 0, flat.size - 1)
        iv = np.clip(W * (H * (2 * n + 1) + sy) + sx, 0, flat.size - 1)
        x2, y2 = (x1 + flat[iu]).astype(f), (y1 + flat[iv]).astype(f)
        x3 = ((x2 * b[0]).astype(f) + (y2 * b[2]).astype(f)).astype(f) + b[4]
        y3 = ((x2 * b[1]).astype(f) + (y2 * b[3]).astype(f)).astype(f) + b[5]
        out[n, 0], out[n, 1] = x3 - xs, y3 - ys
    return out


def _aug_det_value(gen, as_bool=False):
    """caffe_rng_generate (util/rng.cpp:8-114) for generators without randomness (spread 0, prob 0 or 1): the deterministic part of
    the reference's sampler -- its random stream (boost::mt19937 through an unpinned boost) is not reproducible."""
    f = np.float32
    t = gen.get("rand_type", "uniform")
    assert float(gen.get("spread", 0)) == 0, "oracle: only deterministic generators"
    if t in ("uniform", "gaussian"):
        v = f(gen.get("mean", 0))
        if gen.get("exp", False):
            v = f(np.exp(v))
    elif t == "bernoulli":
        assert float(gen.get("prob", 0)) in (0.0, 1.0)
        v = f(1 if float(gen.get("prob", 0)) > 0 else 0)
    elif t in ("uniform_bernoulli", "gaussian_bernoulli"):
        assert float(gen.get("prob", 0)) == 1.0
        v = f(gen.get("mean", 0))
        if gen.get("exp", False):
            v = f(np.exp(v))
    else:
        raise ValueError(t)
    if as_bool:
        v = f(1 if v != 0 else 0)
    if gen.get("discretize", False):
        v = f(np.round(v))
    v = f(f(gen.get("multiplier", 1)) * v)
    return f(1 if v != 0 else 0) if as_bool else v


def generate_augmentation_parameters(in_params, mode, gens, crop_w, crop_h, bottom_w, bottom_h, num=None):
    """GenerateAugmentationParametersLayer::Forward_gpu (generate_augmentation_parameters_layer.cu:16-117) for deterministic
    generators.  in_params: (N, 42) array form or None ("regenerate"); gens: {generator name: {field: value}} as in
    AugmentationParameter (caffe.proto:489-546).  -> (N, 42) float32 array form."""
    f = np.float32
    D = AUG_COEFF_DEFAULT
    def to_arr(v):
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.where(np.abs(D) < 1e-3, v, np.log(v).astype(f)).astype(f)
    N = num if in_params is None else np.asarray(in_params).reshape(-1, 42).shape[0]
    out = np.empty((N, 42), f)
    has = lambda *names: any(n in gens for n in names)
    spatial = has("mirror", "rotate", "zoom", "translate", "squeeze", "translate_x", "translate_y")
    chromatic = has("brightness", "gamma", "contrast", "color")
    effect = has("fog_size", "fog_amount", "motion_blur_angle", "motion_blur_size", "shadow_angle", "shadow_distance", "shadow_strength", "noise")
    eigen = has("lmult_pow", "lmult_mult", "lmult_add", "sat_pow", "sat_mult", "sat_add", "col_pow", "col_mult", "col_add", "ladd_pow",
                "ladd_mult", "ladd_add", "col_rotate")
    g = lambda n, as_bool=False: _aug_det_value(gens[n], as_bool)

    def gen_spatial():                                   # generate_spatial_coeffs, augmentation_layer_base.cpp:73-99
        v = D.copy()
        if "mirror" in gens: v[0] = g("mirror", True)
        if "translate" in gens: v[1] = g("translate"); v[2] = g("translate")
        if "translate_x" in gens: v[1] = g("translate_x")
        if "translate_y" in gens: v[2] = g("translate_y")
        if "rotate" in gens: v[3] = g("rotate")
        if "zoom" in gens: v[4] = g("zoom"); v[5] = v[4]
        if "squeeze" in gens:
            sq = g("squeeze"); v[4] = f(v[4] * sq); v[5] = f(v[5] / sq)
        return v

    def gen_group(v, which):                             # :252-336; returns the record with the group's fields set
        v = v.copy()
        if which == "chromatic":
            for n, i in (("gamma", 6), ("brightness", 7), ("contrast", 8)):
                if n in gens: v[i] = g(n)
            if "color" in gens: v[9:12] = g("color")
        elif which == "eigen":
            pairs = (("ladd_pow", (12,)), ("col_pow", (13, 14)), ("ladd_add", (15,)), ("col_add", (16, 17)), ("ladd_mult", (18,)),
                     ("col_mult", (19, 20)), ("sat_pow", (22, 23)), ("sat_add", (25, 26)), ("sat_mult", (28, 29)), ("lmult_pow", (30,)),
                     ("lmult_mult", (32,)), ("lmult_add", (31,)), ("col_rotate", (33,)))
            for n, idx in pairs:
                if n in gens:
                    for i in idx: v[i] = g(n)
        else:
            z = lambda n: g(n) if n in gens else f(0)    # an absent generator of a present group draws its default (mean 0)
            if has("fog_amount", "fog_size"): v[34] = z("fog_amount"); v[35] = z("fog_size")
            if has("motion_blur_angle", "motion_blur_size"): v[36] = z("motion_blur_angle"); v[37] = z("motion_blur_size")
            if has("shadow_angle", "shadow_distance", "shadow_strength"):
                v[38] = z("shadow_angle"); v[39] = z("shadow_distance"); v[40] = z("shadow_strength")
            if "noise" in gens: v[41] = g("noise")
        return v

    def valid(v):                                        # the 4-corner test of generate_valid_spatial_coeffs, :118-160
        good = 0
        for x in sorted({0, max(0, crop_w - 1)}) if crop_w > 1 else [0]:
            for y in sorted({0, max(0, crop_h - 1)}) if crop_h > 1 else [0]:
                x1 = f(-x + 0.5 * crop_w) if v[0] else f(x - 0.5 * crop_w)
                y1 = f(y - 0.5 * crop_h)
                x2 = f(np.cos(v[3]) * x1 - np.sin(v[3]) * y1); y2 = f(np.sin(v[3]) * x1 + np.cos(v[3]) * y1)
                x2 = f(x2 + v[1] * f(crop_w)); y2 = f(y2 + v[2] * f(crop_h))
                x2 = f(x2 / v[4]); y2 = f(y2 / v[5])
                x2 = f(x2 + 0.5 * bottom_w); y2 = f(y2 + 0.5 * bottom_h)
                if not (np.floor(x2) < 0 or np.floor(x2) > bottom_w - 2 or np.floor(y2) < 0 or np.floor(y2) > bottom_h - 2):
                    good += 1
        return good == 4

    for n in range(N):
        use_in = mode in ("add", "replace") and in_params is not None
        v = aug_array_to_values(np.asarray(in_params).reshape(-1, 42)[n])[0] if use_in else D.copy()
        if spatial:
            if mode == "replace":
                v[0:6] = D[0:6]
            base = to_arr(v)
            cand = aug_array_to_values(to_arr(gen_spatial()) + base)[0]      # deterministic: every try draws the same record
            v = cand if valid(cand) else aug_array_to_values(base)[0]
        o = to_arr(v)
        fresh = mode in ("regenerate", "replace") or in_params is None
        for on, which in ((chromatic, "chromatic"), (eigen, "eigen"), (effect, "effect")):
            if not on:
                continue
            if fresh:
                v = gen_group(v, which)
                o = to_arr(v)
            else:
                o = (o + to_arr(gen_group(D.copy(), which))).astype(f)
        out[n] = o
    return out


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
