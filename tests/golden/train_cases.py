"""Case list of the training-side layers (gradients of Convolution / Deconvolution, L1Loss, Downsample, FlowAugmentation,
GenerateAugmentationParameters), shared by tests/golden/make_train_golden.py (runs the REFERENCE's layer classes, oracle/_ref, on
the GPU box -- L1Loss / Downsample / FlowAugmentation / GenerateAugmentationParameters have no CPU path in the reference,
l1loss_layer.cpp:93-102 etc. -- and writes tests/golden/train_golden.npz) and by the tests that compare the oracle (CPU) and the
CUDA kernels (GPU) with those vectors.  Inputs are regenerated from the seeds, only the reference's outputs are stored."""
import numpy as np

from ref_cases import coeff_blob, rng_for, COEFF_DEFAULT  # noqa: F401


def _conv(name, deconv, N, Ci, H, W, Co, k, s, p):
    typ = "Deconvolution" if deconv else "Convolution"
    text = ('name: "%s" type: "%s" bottom: "x" top: "y" convolution_param { num_output: %d kernel_size: %d stride: %d pad: %d }'
            % (name, typ, Co, k, s, p))
    wshape = (Ci, Co, k, k) if deconv else (Co, Ci, k, k)
    return dict(kind="conv_bwd", text=text, args=(s, p, deconv),
                inputs=lambda r: [r.standard_normal((N, Ci, H, W)).astype(np.float32)],
                params=lambda r: [(r.standard_normal(wshape) / np.sqrt(Ci * k * k)).astype(np.float32), r.standard_normal(Co).astype(np.float32)])


def _flow_pair(r, N, H, W, nan_blocks=True):
    a = (3 * r.standard_normal((N, 2, H, W))).astype(np.float32)
    b = (3 * r.standard_normal((N, 2, H, W))).astype(np.float32)
    if nan_blocks:                                   # invalid ground truth: NaN in both channels (and one single-channel NaN)
        b[0, :, 1:4, 2:7] = np.nan
        b[-1, 0, H - 2, W - 3] = np.nan
    a[0, :, 5, 5] = b[0, :, 5, 5]                    # exact zero difference (sign(0) = -1, sqrt(eps))
    a[0, 0, 6, 1:4] = b[0, 0, 6, 1:4] + np.float32(0.01)   # below any plateau
    a[0, 1, 6, 1:4] = b[0, 1, 6, 1:4] - np.float32(0.02)
    return a, b


def _l1(name, N=2, H=12, W=16, single=False, **kw):
    fields = " ".join("%s: %s" % (k, str(v).lower() if isinstance(v, bool) else v) for k, v in kw.items())
    bots = 'bottom: "a"' if single else 'bottom: "a" bottom: "b"'
    text = 'name: "%s" type: "L1Loss" %s top: "loss" l1_loss_param { %s }' % (name, bots, fields)

    def inputs(r):
        a, b = _flow_pair(r, N, H, W, nan_blocks=not single)
        if single:
            a = (a - b).astype(np.float32)
            a[0, :, 1:3, 2:5] = np.nan
            return [a]
        return [a, b]
    return dict(kind="l1loss", text=text, inputs=inputs, args=kw, top_diff=1.7)


def _down(name, shape, th, tw, two=False, nan=True):
    if two:
        text = 'name: "%s" type: "Downsample" bottom: "x" bottom: "like" top: "y"' % name
    else:
        text = 'name: "%s" type: "Downsample" bottom: "x" top: "y" downsample_param { top_height: %d top_width: %d }' % (name, th, tw)

    def inputs(r):
        x = r.standard_normal(shape).astype(np.float32)
        if nan:
            x[0, :, 3:12, 5:19] = np.nan            # block of invalid flow: some outputs become NaN, some average around it
            x[-1, 0, 0, 0] = np.nan
        return [x] + ([np.zeros((shape[0], 1, th, tw), np.float32)] if two else [])
    return dict(kind="downsample", text=text, inputs=inputs, args=(th, tw))


def _mild_coeffs(r, N, scale):
    v = np.tile(COEFF_DEFAULT, (N, 1)).astype(np.float64)
    v[:, 0] = r.integers(0, 2, N) if scale > 0 else 0
    v[:, 1:3] = r.uniform(-0.03, 0.03, (N, 2)) * scale
    v[:, 3] = r.uniform(-0.05, 0.05, N) * scale
    v[:, 4:6] = np.exp(r.uniform(-0.03, 0.03, (N, 2)) * scale)
    with np.errstate(divide="ignore", invalid="ignore"):
        arr = np.where(np.abs(COEFF_DEFAULT) < 1e-3, v, np.log(v))
    return arr.astype(np.float32).reshape(N, 42, 1, 1)


def _flowaug(name, N, H, W, cw, ch, scale=1.0):
    text = ('name: "%s" type: "FlowAugmentation" bottom: "flow" bottom: "p1" bottom: "p2" top: "y" '
            'augmentation_param { crop_width: %d crop_height: %d }' % (name, cw, ch))
    return dict(kind="flow_aug", text=text, args=(cw, ch),
                inputs=lambda r: [(4 * r.standard_normal((N, 2, H, W))).astype(np.float32), _mild_coeffs(r, N, scale), _mild_coeffs(r, N, scale)])


# deterministic generators only (spread 0, prob 1): the reference's random STREAM (boost::mt19937 through an unpinned boost) is
# not reproducible, the layer's mode logic and array arithmetic are
_GENS = ('translate { rand_type: "uniform" mean: 0.04 spread: 0 } zoom { rand_type: "uniform" exp: true mean: 0.1 spread: 0 } '
         'rotate { rand_type: "uniform_bernoulli" mean: 0.05 spread: 0 prob: 1.0 } '
         'gamma { rand_type: "uniform" exp: true mean: 0.2 spread: 0 } brightness { rand_type: "uniform" mean: 0.1 spread: 0 } '
         'lmult_mult { rand_type: "uniform" exp: true mean: -0.1 spread: 0 } col_rotate { rand_type: "uniform" mean: 0.3 spread: 0 } '
         'noise { rand_type: "uniform" mean: 0.02 spread: 0 }')


def _genaug(name, mode, image_bottom=False):
    size = "" if image_bottom else "bottomwidth: 64 bottomheight: 48 "
    text = ('name: "%s" type: "GenerateAugmentationParameters" bottom: "p" top: "q" augmentation_param { augment_during_test: true '
            'mode: "%s" crop_width: 48 crop_height: 32 %s%s }' % (name, mode, size, _GENS))
    if image_bottom:
        inputs = lambda r: [r.standard_normal((3, 3, 48, 64)).astype(np.float32)]
    else:
        inputs = lambda r: [coeff_blob(r, 3, effects=False)]
    return dict(kind="gen_aug", text=text, inputs=inputs, args=(mode, image_bottom))


TRAIN_CASES = {
    # ---- ConvolutionLayer / DeconvolutionLayer::Backward (conv_layer.cu:26-58, deconv_layer.cu:26-55) ----------------------
    "cb_3x3": _conv("cb_3x3", False, 2, 16, 12, 20, 24, 3, 1, 1),
    "cb_5x5_s2": _conv("cb_5x5_s2", False, 2, 8, 21, 27, 16, 5, 2, 2),
    "cb_7x7_s2": _conv("cb_7x7_s2", False, 1, 3, 32, 48, 16, 7, 2, 3),
    "cb_3x3_s2_even": _conv("cb_3x3_s2_even", False, 2, 32, 18, 24, 64, 3, 2, 1),
    "cb_1x1": _conv("cb_1x1", False, 2, 32, 9, 11, 8, 1, 1, 0),
    "db_4x4_s2": _conv("db_4x4_s2", True, 2, 32, 6, 8, 16, 4, 2, 1),
    "db_4x4_flow": _conv("db_4x4_flow", True, 2, 2, 6, 8, 2, 4, 2, 1),
    # ---- L1Loss (l1loss_layer.cu:67-192) ------------------------------------------------------------------------------------
    "l1_plain": _l1("l1_plain"),
    "l1_epe": _l1("l1_epe", l2_per_location=True),                                                   # the FlowNet2 training loss
    "l1_epe_norm_plateau": _l1("l1_epe_norm_plateau", l2_per_location=True, normalize_by_num_entries=True, plateau=0.5, epsilon=0.001),
    "l1_epe_prescale": _l1("l1_epe_prescale", l2_per_location=True, l2_prescale_by_channels=True),
    "l1_plateau_norm": _l1("l1_plateau_norm", plateau=0.3, normalize_by_num_entries=True),
    "l1_single": _l1("l1_single", single=True, l2_per_location=True, normalize_by_num_entries=True),
    # ---- Downsample (downsample_layer.cu:15-80) -----------------------------------------------------------------------------
    "ds_4x": _down("ds_4x", (2, 2, 32, 48), 8, 12),
    "ds_frac": _down("ds_frac", (2, 2, 32, 48), 7, 13),
    "ds_2x_like": _down("ds_2x_like", (2, 2, 32, 48), 16, 24, two=True),
    "ds_same": _down("ds_same", (1, 2, 8, 12), 8, 12, nan=False),
    # ---- FlowAugmentation (flow_augmentation_layer.cu:24-166) ---------------------------------------------------------------
    "fa_mild": _flowaug("fa_mild", 3, 40, 56, 32, 24),
    "fa_identity": _flowaug("fa_identity", 2, 24, 32, 24, 16, scale=0.0),
    # ---- GenerateAugmentationParameters (generate_augmentation_parameters_layer.cu:16-117) ----------------------------------
    "ga_add": _genaug("ga_add", "add"),
    "ga_replace": _genaug("ga_replace", "replace"),
    "ga_regenerate": _genaug("ga_regenerate", "add", image_bottom=True),
}


def train_inputs(name):
    """-> (bottoms, params or None, rng positioned after the inputs) for a case, deterministic."""
    c = TRAIN_CASES[name]
    r = rng_for(name)
    bottoms = c["inputs"](r)
    params = c["params"](r) if c.get("params") else None
    return bottoms, params, r


# ---- whole-net gradients against the reference's own Net semantics ------------------------------------------------------------
# FlowNet2-C deploy graph up to predict_flow2 + one L1Loss (end-point error) per pyramid level against ground-truth inputs: what
# Net::Backward (Split layers inserted, net.cpp:640-655) makes of it in the reference, for OracleNet.backward and the engine.
LOSS_NET = dict(w=192, h=100, batch=1, seed=1701, weights={6: 0.32, 5: 0.08, 4: 0.02, 3: 0.01, 2: 0.005})
# small parameter blobs are stored whole, the others as 4 seeded random projections + their L2 norm
LOSS_NET_FULL = ("predict_flow6", "predict_flow2", "upsample_flow6to5", "conv_redir", "conv1")


def loss_net_proto():
    import re
    import flownet2_b200 as F
    w, h, batch = LOSS_NET["w"], LOSS_NET["h"], LOSS_NET["batch"]
    proto = F.fill_template(F.model_template("FlowNet2-C"), w, h)
    cut = proto.index('layer {\n  name: "Eltwise_final_x20"')
    body = proto[:cut]
    aw, ah = (w + 63) // 64 * 64, (h + 63) // 64 * 64
    first_layer = body.index("layer {")
    gts, losses = "", ""
    for lvl, wgt in LOSS_NET["weights"].items():
        sw, sh = aw >> lvl, ah >> lvl
        gts += 'input: "gt%d"\ninput_shape {\n  dim: %d\n  dim: 2\n  dim: %d\n  dim: %d\n}\n' % (lvl, batch, sh, sw)
        losses += ('layer {\n  name: "flow_loss%d"\n  type: "L1Loss"\n  bottom: "predict_flow%d"\n  bottom: "gt%d"\n  top: "flow_loss%d"\n'
                   '  loss_weight: %g\n  l1_loss_param {\n    l2_per_location: true\n  }\n}\n' % (lvl, lvl, lvl, lvl, wgt))
    text = body[:first_layer] + gts + body[first_layer:] + losses
    if batch != 1:
        text = re.sub(r"dim: 1\n  dim: 3", "dim: %d\n  dim: 3" % batch, text)
    return text


def loss_net_inputs():
    import sys
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from util import rng, smooth_images
    w, h, batch = LOSS_NET["w"], LOSS_NET["h"], LOSS_NET["batch"]
    img0, img1 = smooth_images(rng(21), batch, h, w)
    r = rng(22)
    aw, ah = (w + 63) // 64 * 64, (h + 63) // 64 * 64
    ins = {"img0": img0, "img1": img1}
    for lvl in LOSS_NET["weights"]:
        g = (0.5 * r.standard_normal((batch, 2, ah >> lvl, aw >> lvl))).astype(np.float32)
        if lvl == 2:
            g[0, :, 3:6, 10:20] = np.nan                     # invalid ground truth
        ins["gt%d" % lvl] = g
    return ins


def grad_signature(name, blob_index, g):
    """What the golden file keeps of one parameter gradient."""
    g = np.asarray(g, np.float64).reshape(-1)
    if name in LOSS_NET_FULL:
        return g.astype(np.float32)
    r = np.random.default_rng([77, blob_index] + [ord(c) for c in name])
    proj = r.standard_normal((4, g.size)) / np.sqrt(g.size)
    return np.concatenate([proj @ g, [np.sqrt((g * g).sum())]]).astype(np.float32)
