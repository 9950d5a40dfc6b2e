"""CPU oracle for whole-net forward passes (TEST INFRASTRUCTURE, see oracle/oracle.c header).

An independent, deliberately simple restatement of the reference's graph semantics
(src/caffe/net.cpp:40-286 in-order execution, in-place tops :394-400; weight matching by layer
name :752-802) on top of oracle.py's per-layer functions.  It has its own prototxt text-format
parser and its own .caffemodel wire-format reader so that nothing is shared with the product's
C++ parsers.
"""
import struct

import numpy as np

from . import oracle as O


# ------------------------------------------------------------------------------------------------
# text format
# ------------------------------------------------------------------------------------------------
def _tokens(text):
    i, n = 0, len(text)
    while i < n:
        c = text[i]
        if c in " \t\r\n,;":
            i += 1
        elif c == "#":
            while i < n and text[i] != "\n":
                i += 1
        elif c in "{}:<>[]":
            yield c
            i += 1
        elif c in "\"'":
            j = i + 1
            out = []
            while text[j] != c:
                if text[j] == "\\":
                    j += 1
                out.append(text[j])
                j += 1
            yield ("str", "".join(out))
            i = j + 1
        else:
            j = i
            while j < n and text[j] not in " \t\r\n,;{}:<>[]#\"'":
                j += 1
            yield ("tok", text[i:j])
            i = j


def parse_prototxt(text):
    toks = list(_tokens(text))
    pos = [0]

    def parse_msg(closer):
        fields = []
        while pos[0] < len(toks):
            t = toks[pos[0]]
            if t == closer:
                pos[0] += 1
                return fields
            name = t[1]
            pos[0] += 1
            if toks[pos[0]] == ":":
                pos[0] += 1
            t = toks[pos[0]]
            if t in ("{", "<"):
                pos[0] += 1
                fields.append((name, parse_msg("}" if t == "{" else ">")))
            elif t == "[":
                pos[0] += 1
                while toks[pos[0]] != "]":
                    fields.append((name, toks[pos[0]][1]))
                    pos[0] += 1
                pos[0] += 1
            else:
                fields.append((name, t[1]))
                pos[0] += 1
        assert closer is None, "unbalanced braces"
        return fields

    return parse_msg(None)


def getall(msg, name):
    return [v for k, v in msg if k == name]


def get(msg, name, default=None):
    v = getall(msg, name)
    return v[0] if v else default


# ------------------------------------------------------------------------------------------------
# wire format (.caffemodel): NetParameter.layer=100 -> name=1, type=2, blobs=7 -> shape=7{dim=1},
# data=5 (packed float), legacy num/channels/height/width = 1..4   (caffe.proto:10-22,94,313-331)
# ------------------------------------------------------------------------------------------------
def _varint(b, i):
    v, s = 0, 0
    while True:
        x = b[i]
        i += 1
        v |= (x & 0x7F) << s
        if not x & 0x80:
            return v, i
        s += 7


def _fields(b):
    i = 0
    while i < len(b):
        key, i = _varint(b, i)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _varint(b, i)
        elif wt == 2:
            ln, i = _varint(b, i)
            v = b[i:i + ln]
            i += ln
        elif wt == 5:
            v = b[i:i + 4]
            i += 4
        elif wt == 1:
            v = b[i:i + 8]
            i += 8
        else:
            raise ValueError("wire type %d" % wt)
        yield fn, wt, v


def parse_caffemodel(data):
    layers = {}
    for fn, wt, v in _fields(memoryview(data)):
        if fn != 100:
            continue
        name, blobs = None, []
        for f2, w2, v2 in _fields(v):
            if f2 == 1:
                name = bytes(v2).decode()
            elif f2 == 7:
                shape, legacy, arr = None, {}, None
                for f3, w3, v3 in _fields(v2):
                    if f3 == 7:
                        for f4, w4, v4 in _fields(v3):
                            if f4 == 1 and w4 == 2:
                                shape, j = [], 0
                                while j < len(v4):
                                    d, j = _varint(v4, j)
                                    shape.append(d)
                    elif f3 == 5 and w3 == 2:
                        arr = np.frombuffer(v3, dtype="<f4").copy()
                    elif f3 == 8 and w3 == 2 and arr is None:      # double_data (blob.cpp:472-476)
                        arr = np.frombuffer(v3, dtype="<f8").astype(np.float32)
                    elif f3 in (1, 2, 3, 4) and w3 == 0:
                        legacy[f3] = v3
                if shape is None:
                    shape = [legacy.get(k, 1) for k in (1, 2, 3, 4)]
                blobs.append(arr.reshape(shape))
        layers[name] = blobs
    return layers


def _enc_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _enc_field(fn, payload):
    return _enc_varint((fn << 3) | 2) + _enc_varint(len(payload)) + payload


def write_caffemodel(layers):
    """layers: [(name, type, [ndarray, ...])] -> bytes of a binary NetParameter (inverse of parse_caffemodel; BlobProto with
    shape=7{dim=1 packed} and data=5 packed float, caffe.proto:10-22; LayerParameter name=1 type=2 blobs=7 in NetParameter.layer=100)."""
    out = bytearray()
    for name, typ, blobs in layers:
        body = _enc_field(1, name.encode()) + _enc_field(2, typ.encode())
        for a in blobs:
            a = np.ascontiguousarray(a, dtype="<f4")
            dims = b"".join(_enc_varint(int(d)) for d in a.shape)
            body += _enc_field(7, _enc_field(7, _enc_field(1, dims)) + _enc_field(5, a.tobytes()))
        out += _enc_field(100, body)
    return bytes(out)


def synth_weights(prototxt_text, seed, real_prototxt=None):
    """Deterministic synthetic weights for a deploy prototxt (He-initialised conv/deconv, zero bias, DataAugmentation state past
    the recompute_mean threshold): {layer name: [blobs]} plus the same as caffemodel bytes.  The shapes come from one oracle
    forward at the prototxt's own input size, so call it with a SMALL fill of the template; conv shapes do not depend on H, W.
    real_prototxt: the fill of the same template at the size that will actually run; layers whose filler depends on the
    template variables (the diagonal scale convolution, $SCALE_WIDTH$/$SCALE_HEIGHT$) are taken from it."""
    net = OracleNet(prototxt_text, None, batch=1, synth_seed=seed)
    ins = {n: np.zeros(s, np.float32) for n, s in zip(net.input_names, net.input_shapes)}
    net.forward(**ins)
    types = {get(l, "name"): get(l, "type") for l in net.layers}
    ordered = [(get(l, "name"), types[get(l, "name")], net.weights[get(l, "name")]) for l in net.layers if get(l, "name") in net.weights]
    for name, typ, blobs in ordered:
        if typ == "DataAugmentation":          # restore the iteration counter the zero-input forward advanced; drop the sized mean
            rm = float(blobs[0].reshape(-1)[0]) - 1
            blobs[0] = np.full((1, 1, 1, 1), rm, np.float32)
            blobs[1] = np.full((1, blobs[1].shape[1], 1, 1), 0.4, np.float32)
    if real_prototxt is not None:
        for l in getall(parse_prototxt(real_prototxt), "layer"):
            wf = get(get(l, "convolution_param", []), "weight_filler", [])
            if get(wf, "type") == "diagonal":
                dv = [float(x) for x in getall(wf, "diag_val")]
                for name, typ, blobs in ordered:
                    if name == get(l, "name"):
                        for i in range(min(blobs[0].shape[0], blobs[0].shape[1])):
                            blobs[0][i, i, :, :] = dv[i] if i < len(dv) else dv[-1]
    return {n: b for n, _, b in ordered}, write_caffemodel(ordered)


# ------------------------------------------------------------------------------------------------
# forward
# ------------------------------------------------------------------------------------------------
def _hw(msg, rep, h, w, default):
    if get(msg, h) is not None or get(msg, w) is not None:
        return int(get(msg, h, default)), int(get(msg, w, default))
    v = [int(x) for x in getall(msg, rep)]
    if not v:
        return default, default
    return (v[0], v[0]) if len(v) == 1 else (v[0], v[1])


class OracleNet(object):
    def __init__(self, prototxt_text, caffemodel_bytes, batch=0, f64acc=False, synth_seed=None):
        self.root = parse_prototxt(prototxt_text)
        self.weights = parse_caffemodel(caffemodel_bytes) if caffemodel_bytes else {}
        self.synth_seed = synth_seed
        self.f64acc = f64acc
        self.input_names = getall(self.root, "input")
        self.input_shapes = [[int(d) for d in getall(s, "dim")] for s in getall(self.root, "input_shape")]
        self.layers = getall(self.root, "layer")
        for l in self.layers:
            if get(l, "type") == "Input":
                self.input_names += getall(l, "top")
                self.input_shapes += [[int(d) for d in getall(s, "dim")] for s in getall(get(l, "input_param", []), "shape")]
        if batch:
            for s in self.input_shapes:
                s[0] = batch
        self.blobs = {}

    def forward(self, **inputs):
        B = self.blobs = {}
        for n, s in zip(self.input_names, self.input_shapes):
            a = np.ascontiguousarray(inputs[n], np.float32)
            assert list(a.shape) == s, (n, a.shape, s)
            B[n] = a
        for l in self.layers:
            if get(l, "type") == "Input":
                continue
            bots = [B[b] for b in getall(l, "bottom")]
            for t, v in zip(getall(l, "top"), self.run_layer(l, bots)):
                B[t] = v
        return B

    def run_layer(self, l, bots):
        """One layer of the prototxt on numpy bottoms -> list of numpy tops (layer semantics: see the per-type comments)."""
        typ, name = get(l, "type"), get(l, "name")
        tops = getall(l, "top")
        B = {}
        if typ == "Eltwise":
            ep = get(l, "eltwise_param", [])
            assert get(ep, "operation", "SUM") == "SUM"
            B[tops[0]] = O.eltwise_sum(bots, [float(c) for c in getall(ep, "coeff")])
        elif typ == "DataAugmentation":
            B[tops[0]] = self._aug(l, name, bots[0])
        elif typ == "Resample":
            rp = get(l, "resample_param", [])
            if len(bots) == 2:
                oh, ow = bots[1].shape[2], bots[1].shape[3]
            else:
                oh, ow = int(get(rp, "height")), int(get(rp, "width"))
            rtype = {"NEAREST": 1, "LINEAR": 2, "CUBIC": 3}[get(rp, "type", "LINEAR")]
            B[tops[0]] = O.resample_fwd(bots[0], oh, ow, rtype, get(rp, "antialias", "true") == "true")
        elif typ in ("Convolution", "Deconvolution"):
            cp = get(l, "convolution_param")
            kh, kw = _hw(cp, "kernel_size", "kernel_h", "kernel_w", 0)
            sh, sw = _hw(cp, "stride", "stride_h", "stride_w", 1)
            ph, pw = _hw(cp, "pad", "pad_h", "pad_w", 0)
            has_bias = get(cp, "bias_term", "true") == "true"
            if name not in self.weights:           # synthetic weights (timing-only runs)
                assert self.synth_seed is not None, "no weights for layer " + name
                co, ci = int(get(cp, "num_output")), bots[0].shape[1]
                r = np.random.default_rng(self.synth_seed + len(self.weights))
                shp = (co, ci, kh, kw) if typ == "Convolution" else (ci, co, kh, kw)
                wf = get(cp, "weight_filler", [])
                if get(wf, "type") == "diagonal":          # DiagonalFiller, filler.hpp:265-290
                    dv = [float(x) for x in getall(wf, "diag_val")]
                    wsyn = np.zeros(shp, np.float32)
                    for i in range(min(shp[0], shp[1])):
                        wsyn[i, i, :, :] = dv[i] if i < len(dv) else (dv[-1] if dv else 1.0)
                else:
                    wsyn = (r.standard_normal(shp) * np.sqrt(2.0 / (ci * kh * kw))).astype(np.float32)
                self.weights[name] = [wsyn] + ([np.zeros(co, np.float32)] if has_bias else [])
            w = self.weights[name][0]
            b = self.weights[name][1].reshape(-1) if has_bias else None
            for bot, top in zip(bots, tops):
                if typ == "Convolution":
                    B[top] = O.conv_fwd(bot, w, b, (sh, sw), (ph, pw), f64acc=self.f64acc)
                else:
                    B[top] = O.deconv_fwd(bot, w, b, (sh, sw), (ph, pw), f64acc=self.f64acc)
        elif typ == "ReLU":
            B[tops[0]] = O.relu(bots[0], float(get(get(l, "relu_param", []), "negative_slope", 0)))
        elif typ == "Concat":
            B[tops[0]] = np.concatenate(bots, axis=1)
        elif typ == "Correlation":
            cp = get(l, "correlation_param")
            B[tops[0]] = O.correlation_fwd(bots[0], bots[1], int(get(cp, "pad", 0)), int(get(cp, "kernel_size")),
                                           int(get(cp, "max_displacement")), int(get(cp, "stride_1", 1)),
                                           int(get(cp, "stride_2", 1)),
                                           {"MULTIPLY": 0, "SUBTRACT": 1}[get(cp, "correlation_type", "MULTIPLY")],
                                           exact_order=not self.f64acc)
        elif typ == "FlowWarp":
            fp = get(l, "flow_warp_param", [])
            B[tops[0]] = O.flow_warp_fwd(bots[0], bots[1], get(fp, "fill_value", "ZERO") == "NOT_A_NUMBER")
        elif typ == "ChannelNorm":
            B[tops[0]] = O.channel_norm(bots[0])
        elif typ == "L1Loss":
            loss, _ = O.l1loss_fwd(bots[0], bots[1] if len(bots) > 1 else None, **self._l1_args(l))
            B[tops[0]] = np.array([loss], np.float32)
        elif typ == "Downsample":
            dp = get(l, "downsample_param", [])
            th, tw = (bots[1].shape[2], bots[1].shape[3]) if len(bots) > 1 else (int(get(dp, "top_height")), int(get(dp, "top_width")))
            B[tops[0]] = O.downsample_fwd(bots[0], th, tw)
        elif typ == "Silence":
            pass
        else:
            raise NotImplementedError(typ)
        return [B[t] for t in tops]

    @staticmethod
    def _l1_args(l):
        lp = get(l, "l1_loss_param", [])
        tf = lambda k: get(lp, k, "false") == "true"
        return dict(l2_per_location=tf("l2_per_location"), l2_prescale_by_channels=tf("l2_prescale_by_channels"),
                    normalize_by_num_entries=tf("normalize_by_num_entries"), epsilon=float(get(lp, "epsilon", 1e-2)),
                    plateau=float(get(lp, "plateau", 0)))

    # ---- gradients ------------------------------------------------------------------------------------------------------
    def backward(self, **seeds):
        """Net::Backward (net.cpp:640-655) over the blobs of the last forward(): seeds = {blob name: top diff}.  No Split layers
        are inserted; a blob read by several layers simply sums the contributions (what SplitLayer::Backward does,
        split_layer.cpp:38-52).  -> ({blob: diff}, {layer name: [param diffs]}); parameter diffs start from zero."""
        B = self.blobs
        D = {k: np.asarray(v, np.float32).copy() for k, v in seeds.items()}
        for l in self.layers:                   # loss tops are seeded with their loss_weight (layer.hpp:455-478; default 1 for top 0)
            if get(l, "type", "").endswith("Loss"):
                t = getall(l, "top")[0]
                if t not in D:
                    lw = getall(l, "loss_weight")
                    D[t] = np.full(B[t].shape, float(lw[0]) if lw else 1.0, np.float32)
        P = {}
        for l in reversed(self.layers):
            typ = get(l, "type")
            if typ == "Input":
                continue
            tops, bots = getall(l, "top"), getall(l, "bottom")
            if not any(t in D for t in tops):
                continue
            tds = [D[t] if t in D else np.zeros_like(B[t]) for t in tops]
            grads = self.backward_layer(l, [B[b] for b in bots], [B[t] for t in tops], tds, P)
            for b, g in zip(bots, grads):
                if g is None:
                    continue
                if b in tops or b not in D:
                    D[b] = g
                else:
                    D[b] = D[b] + g
        return D, P

    def backward_layer(self, l, bots, tops, tds, P):
        typ, name = get(l, "type"), get(l, "name")
        if typ in ("Convolution", "Deconvolution"):
            cp = get(l, "convolution_param")
            sh, sw = _hw(cp, "stride", "stride_h", "stride_w", 1)
            ph, pw = _hw(cp, "pad", "pad_h", "pad_w", 0)
            has_bias = get(cp, "bias_term", "true") == "true"
            w = self.weights[name][0]
            out, dw, db = [], 0, 0
            for x, td in zip(bots, tds):
                gx, gw, gb = O.conv_bwd(x, w, td, (sh, sw), (ph, pw), typ == "Deconvolution")
                out.append(gx)
                dw, db = dw + gw.astype(np.float64), db + gb.astype(np.float64)
            P[name] = [np.asarray(dw, np.float32)] + ([np.asarray(db, np.float32)] if has_bias else [])
            return out
        if typ == "ReLU":
            return [O.relu_bwd(tops[0], tds[0], float(get(get(l, "relu_param", []), "negative_slope", 0)))]
        if typ == "Eltwise":
            coeffs = [float(c) for c in getall(get(l, "eltwise_param", []), "coeff")] or [1.0] * len(bots)
            return [np.float32(c) * tds[0] for c in coeffs]
        if typ == "Concat":
            out, c0 = [], 0
            for b in bots:
                out.append(tds[0][:, c0:c0 + b.shape[1]].copy())
                c0 += b.shape[1]
            return out
        if typ == "Correlation":
            cp = get(l, "correlation_param")
            ct = {"MULTIPLY": 0, "SUBTRACT": 1}[get(cp, "correlation_type", "MULTIPLY")]
            g0, g1 = O.correlation_bwd_ex(bots[0], bots[1], tds[0], int(get(cp, "pad", 0)), int(get(cp, "kernel_size")),
                                          int(get(cp, "max_displacement")), int(get(cp, "stride_1", 1)), int(get(cp, "stride_2", 1)), ct)
            return [g0, g1]
        if typ == "FlowWarp":
            g0, g1 = O.flow_warp_bwd(bots[0], bots[1], tds[0])
            return [g0, g1]
        if typ == "L1Loss":
            g0, g1 = O.l1loss_bwd(bots[0], bots[1] if len(bots) > 1 else None, float(np.asarray(tds[0]).reshape(-1)[0]), **self._l1_args(l))
            return [g0, g1][:len(bots)]
        if typ in ("DataAugmentation", "Resample", "Silence", "Downsample"):
            return [None] * len(bots)
        raise NotImplementedError("backward of " + typ)

    def _aug(self, l, name, x):
        ap = get(l, "augmentation_param")
        N, C, H, W = x.shape
        cw, ch = get(ap, "crop_width"), get(ap, "crop_height")
        if cw is not None and ch is not None:
            cw, ch = int(cw), int(ch)
            mats = np.stack([O.transmat_from_coeff(cw, ch, W, H)] * N)
            top = O.spatial_augmentation(x, mats, ch, cw)
        else:
            top = x.copy()
        rm = int(get(ap, "recompute_mean", 0))
        mpp_flag = get(ap, "mean_per_pixel", "true") == "true"
        if rm > 0:
            if name not in self.weights:
                assert self.synth_seed is not None, "no weights for layer " + name
                self.weights[name] = [np.full((1, 1, 1, 1), rm + 1, np.float32),
                                      np.full((1,) + top.shape[1:], 0.4, np.float32),
                                      np.full((1, top.shape[1], 1, 1), 0.4, np.float32)]
            st = self.weights[name]
            if st[1].size != top[0].size:        # adjust_blobs' size-mismatch path: channel average, replicated (.cpp:189-201)
                pc = st[1].reshape(top.shape[1], -1).mean(axis=1, dtype=np.float64).astype(np.float32)
                st[1] = np.broadcast_to(pc.reshape(1, -1, 1, 1), (1,) + top.shape[1:]).copy()
            st[0] = np.float32(int(st[0].reshape(-1)[0]) + 1).reshape(st[0].shape)     # :353-354
            num_iter = float(st[0].reshape(-1)[0])
            top, mpp, mpc = O.mean_subtract(top, 0, num_iter, rm, mpp_flag, st[1].reshape(top.shape[1:]),
                                            st[2].reshape(-1))
            st[1], st[2] = mpp.reshape(st[1].shape), mpc.reshape(st[2].shape)
        else:
            means = [float(m) for m in getall(ap, "mean")]
            if len(means) == 3 and not mpp_flag:
                top, _, _ = O.mean_subtract(top, 1, mean_pc=np.array(means, np.float32))
        return top
