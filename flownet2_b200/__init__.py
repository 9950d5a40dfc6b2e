"""flownet2_b200 -- Blackwell-native FlowNet2 forward path.

Python here is only a ctypes front end over the C-ABI in include/fn2.h (libfn2.so: hand-written
sm_100a CUDA + the Caffe-surface C++ host).  It mirrors the slice of pycaffe that the reference's
driver uses (scripts/run-flownet.py:64-98): ``Net(prototxt, weights, phase)``, ``net.inputs``,
``net.forward(img0=..., img1=...)``, ``net.blobs[name].data``.

There is NO CPU fallback: importing works anywhere (so the CPU test-suite can check symbols), but
any compute call without the built library or without a GPU raises.
"""
import ctypes as C
import math
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
TRAIN, TEST = 0, 1


class Fn2Error(RuntimeError):
    pass


class fn2_tensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("n", C.c_int32), ("c", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("sn", C.c_int64), ("sc", C.c_int64), ("sh", C.c_int64), ("sw", C.c_int64)]


class fn2_conv_desc(C.Structure):
    _fields_ = [("ci", C.c_int32), ("co", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
                ("stride_h", C.c_int32), ("stride_w", C.c_int32), ("pad_h", C.c_int32), ("pad_w", C.c_int32),
                ("deconv", C.c_int32), ("has_bias", C.c_int32), ("relu", C.c_int32), ("negative_slope", C.c_float),
                ("engine", C.c_int32), ("input_guard_bytes", C.c_int32), ("out_pad_h", C.c_int32), ("out_pad_w", C.c_int32)]


def lib_path():
    return os.environ.get("FN2_LIB") or os.path.join(_HERE, "libfn2.so")      # FN2_LIB: A/B a differently built library


def lib():
    """Load libfn2.so (built in-tree by flownet2_b200.build).  Fails loudly if it is missing."""
    global _LIB
    if _LIB is None:
        p = lib_path()
        if not os.path.exists(p):
            raise Fn2Error("libfn2.so is not built (%s); run `python -m flownet2_b200.build` -- there is no "
                           "CPU/PyTorch fallback" % p)
        l = C.CDLL(p)
        l.fn2_last_error.restype = C.c_char_p
        l.fn2_version.restype = C.c_char_p
        l.fn2_launch_count.restype = C.c_uint64
        for name in ("fn2_net_input_name", "fn2_net_output_name", "fn2_net_blob_name", "fn2_net_layer_name",
                     "fn2_net_layer_type"):
            getattr(l, name).restype = C.c_char_p
            getattr(l, name).argtypes = [C.c_void_p, C.c_int]
        l.fn2_net_stream.restype = C.c_void_p
        l.fn2_net_stream.argtypes = [C.c_void_p]
        l.fn2_net_create_batch.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        l.fn2_net_destroy.argtypes = [C.c_void_p]
        l.fn2_net_destroy.restype = None
        for name in ("fn2_net_num_inputs", "fn2_net_num_outputs", "fn2_net_num_blobs", "fn2_net_num_layers",
                     "fn2_net_forward", "fn2_net_sync", "fn2_net_params_changed", "fn2_net_launches_per_forward",
                     "fn2_net_graph_active", "fn2_net_backward", "fn2_net_clear_param_diffs", "fn2_net_launches_per_backward"):
            getattr(l, name).argtypes = [C.c_void_p]
        l.fn2_net_param_diff_arena.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        l.fn2_net_set_diff.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        l.fn2_net_get_diff.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        l.fn2_net_param_shape.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int)]
        l.fn2_net_get_param.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_void_p]
        l.fn2_net_layer_need_backward.argtypes = [C.c_void_p, C.c_int]
        l.fn2_net_copy_trained_layers.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        l.fn2_net_to_caffemodel.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]
        l.fn2_net_to_hdf5.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]
        l.fn2_caffemodel_to_hdf5.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t)]
        l.fn2_hdf5_summary.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.POINTER(C.c_size_t)]
        l.fn2_net_fill_params.argtypes = [C.c_void_p, C.c_uint64]
        l.fn2_net_param_arena.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        l.fn2_net_blob_shape.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
        for name in ("fn2_net_set_input", "fn2_net_get_blob", "fn2_net_set_input_device", "fn2_net_get_blob_device"):
            getattr(l, name).argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        l.fn2_net_time_layers.argtypes = [C.c_void_p, C.c_void_p]
        l.fn2_net_layer_work.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        T = C.POINTER(fn2_tensor)
        l.fn2_correlation_forward.argtypes = [T, T, T, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_void_p, C.c_size_t, C.c_void_p]
        l.fn2_correlation_backward.argtypes = [T, T, T, T, T, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        l.fn2_correlation_workspace_bytes.argtypes = [C.c_int] * 10 + [C.POINTER(C.c_size_t)]
        l.fn2_correlation_shape.argtypes = [C.c_int] * 7 + [C.POINTER(C.c_int)] * 3
        l.fn2_flow_warp_forward.argtypes = [T, T, T, C.c_int, C.c_void_p]
        l.fn2_flow_warp_backward.argtypes = [T, T, T, T, T, C.c_void_p]
        l.fn2_resample_forward.argtypes = [T, T, C.c_int, C.c_int, C.c_void_p]
        l.fn2_spatial_augmentation.argtypes = [T, T, C.c_void_p, C.c_void_p]
        l.fn2_color_contrast_augmentation.argtypes = [T, C.c_void_p, C.c_float, C.c_void_p]
        l.fn2_mean_update.argtypes = [T, T, C.c_void_p, C.c_float, C.c_void_p]
        l.fn2_mean_subtract.argtypes = [T, T, C.c_void_p, C.c_int, C.c_void_p]
        D = C.POINTER(fn2_conv_desc)
        l.fn2_conv_packed_floats.argtypes = [D, C.c_int, C.POINTER(C.c_size_t)]
        l.fn2_conv_pack_weights.argtypes = [D, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        l.fn2_conv_out_shape.argtypes = [D, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        l.fn2_conv_forward.argtypes = [D, T, C.c_void_p, C.c_void_p, T, C.c_void_p, C.c_size_t, C.c_void_p]
        l.fn2_conv_workspace_bytes.argtypes = [D, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
        l.fn2_relu_forward.argtypes = [T, T, C.c_float, C.c_void_p]
        l.fn2_relu_backward.argtypes = [T, T, T, C.c_float, C.c_int, C.c_void_p]
        l.fn2_axpby.argtypes = [T, C.c_float, T, C.c_float, C.c_void_p]
        l.fn2_conv_backward_params_workspace_bytes.argtypes = [D, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
        l.fn2_conv_backward_params.argtypes = [D, T, T, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        l.fn2_conv_backward_data_desc.argtypes = [D, C.c_int, C.c_int, D, C.POINTER(C.c_int)]
        l.fn2_conv_flip_transpose_weights.argtypes = [D, C.c_void_p, C.c_void_p, C.c_void_p]
        l.fn2_eltwise_sum.argtypes = [C.POINTER(T), C.POINTER(C.c_float), C.c_int, T, C.c_void_p]
        l.fn2_channel_norm_forward.argtypes = [T, T, C.c_void_p]
        l.fn2_copy.argtypes = [T, T, C.c_void_p]
        l.fn2_fill.argtypes = [T, C.c_float, C.c_void_p]
        l.fn2_write_flo.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
        l.fn2_read_flo.argtypes = [C.c_char_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_size_t]
        _LIB = l
    return _LIB


def check(rc):
    if rc != 0:
        raise Fn2Error("fn2 error %d: %s" % (rc, lib().fn2_last_error().decode("utf-8", "replace")))


def launch_count():
    return int(lib().fn2_launch_count())


# ------------------------------------------------------------------------------------------------
# Template handling of scripts/run-flownet.py:39-60
# ------------------------------------------------------------------------------------------------
def template_vars(width, height, divisor=64.0):
    v = {"TARGET_WIDTH": width, "TARGET_HEIGHT": height}
    v["ADAPTED_WIDTH"] = int(math.ceil(width / divisor) * divisor)
    v["ADAPTED_HEIGHT"] = int(math.ceil(height / divisor) * divisor)
    v["SCALE_WIDTH"] = width / float(v["ADAPTED_WIDTH"])
    v["SCALE_HEIGHT"] = height / float(v["ADAPTED_HEIGHT"])
    return v


def fill_template(text, width, height):
    for key, value in template_vars(width, height).items():
        text = text.replace("$%s$" % key, str(value))
    return text


def train_template(name="FlowNet2-C"):
    """Text of models/<name>_train.prototxt.template (the authored FlowNet2-C training graph of BASELINE config 5)."""
    p = os.path.join(os.path.dirname(_HERE), "models", "%s_train.prototxt.template" % name)
    with open(p) as f:
        return f.read()


def fill_train_template(text, crop_width, crop_height, data_width, data_height, batch):
    """Training template variables: $TARGET_*$ = the augmentation crop (network input), $DATA_*$ = the frames read, $BATCH$."""
    for key, value in (("TARGET_WIDTH", crop_width), ("TARGET_HEIGHT", crop_height), ("DATA_WIDTH", data_width),
                       ("DATA_HEIGHT", data_height), ("BATCH", batch)):
        text = text.replace("$%s$" % key, str(value))
    return text


def model_template(name):
    """Text of models/<name>_deploy.prototxt.template (FlowNet2, FlowNet2-C, -S, -CSS, -SD)."""
    p = os.path.join(os.path.dirname(_HERE), "models", "%s_deploy.prototxt.template" % name)
    with open(p) as f:
        return f.read()


# ------------------------------------------------------------------------------------------------
# Net (pycaffe-like)
# ------------------------------------------------------------------------------------------------
class _BlobView(object):
    def __init__(self, net, name):
        self._net, self._name = net, name

    @property
    def shape(self):
        s = (C.c_int * 4)()
        check(lib().fn2_net_blob_shape(self._net._h, self._name.encode(), s))
        return tuple(s)

    @property
    def data(self):
        out = np.empty(self.shape, np.float32)
        check(lib().fn2_net_get_blob(self._net._h, self._name.encode(), out.ctypes.data_as(C.c_void_p)))
        return out


class _Blobs(object):
    def __init__(self, net):
        self._net = net

    def __getitem__(self, name):
        if name not in self._net.blob_names:
            raise KeyError(name)
        return _BlobView(self._net, name)

    def __iter__(self):
        return iter(self._net.blob_names)

    def __contains__(self, name):
        return name in self._net.blob_names


class Net(object):
    """caffe.Net(prototxt, weights, phase) look-alike (python/caffe/pycaffe.py:78-124).

    prototxt: path or text (``$VARS$`` already substituted, see fill_template).  weights: path or
    bytes of a .caffemodel, or None (then call fill_params for synthetic weights).  batch > 0
    overrides dim 0 of the Input shapes.
    """

    def __init__(self, prototxt, weights=None, phase=TEST, batch=0):
        l = lib()
        if "\n" not in prototxt and os.path.exists(prototxt):
            with open(prototxt) as f:
                prototxt = f.read()
        h = C.c_void_p()
        check(l.fn2_net_create_batch(prototxt.encode(), int(phase), int(batch), C.byref(h)))
        self._h = h
        self.inputs = [l.fn2_net_input_name(h, i).decode() for i in range(l.fn2_net_num_inputs(h))]
        self.outputs = [l.fn2_net_output_name(h, i).decode() for i in range(l.fn2_net_num_outputs(h))]
        self.blob_names = [l.fn2_net_blob_name(h, i).decode() for i in range(l.fn2_net_num_blobs(h))]
        self.layer_names = [l.fn2_net_layer_name(h, i).decode() for i in range(l.fn2_net_num_layers(h))]
        self.layer_types = [l.fn2_net_layer_type(h, i).decode() for i in range(l.fn2_net_num_layers(h))]
        self.blobs = _Blobs(self)
        if weights is not None:
            self.copy_from(weights)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().fn2_net_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def copy_from(self, weights):
        if isinstance(weights, str):
            with open(weights, "rb") as f:
                weights = f.read()
        buf = (C.c_char * len(weights)).from_buffer_copy(weights)
        check(lib().fn2_net_copy_trained_layers(self._h, buf, len(weights)))

    def fill_params(self, seed=1701):
        check(lib().fn2_net_fill_params(self._h, int(seed)))

    def to_caffemodel(self):
        n = C.c_size_t(0)
        check(lib().fn2_net_to_caffemodel(self._h, None, C.byref(n)))
        buf = (C.c_char * n.value)()
        check(lib().fn2_net_to_caffemodel(self._h, buf, C.byref(n)))
        return bytes(buf[:n.value])

    def to_hdf5(self):
        """Net::ToHDF5 (net.cpp:905-960): the weights as the bytes of a .caffemodel.h5 file."""
        n = C.c_size_t(0)
        check(lib().fn2_net_to_hdf5(self._h, None, C.byref(n)))
        buf = (C.c_char * n.value)()
        check(lib().fn2_net_to_hdf5(self._h, buf, C.byref(n)))
        return bytes(buf[:n.value])

    def param_arena(self):
        p, n = C.c_void_p(), C.c_size_t()
        check(lib().fn2_net_param_arena(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def params_changed(self):
        check(lib().fn2_net_params_changed(self._h))

    def set_input(self, name, array):
        a = np.ascontiguousarray(array, dtype=np.float32)
        if tuple(a.shape) != self.blobs[name].shape:
            raise Fn2Error("input %s has shape %s, net expects %s" % (name, a.shape, self.blobs[name].shape))
        self._keep = getattr(self, "_keep", {})
        self._keep[name] = a       # must outlive the async H2D copy
        check(lib().fn2_net_set_input(self._h, name.encode(), a.ctypes.data_as(C.c_void_p)))

    def set_input_ptr(self, name, host_ptr):
        """host_ptr: address of a (pinned) float32 NCHW buffer of the blob's size."""
        check(lib().fn2_net_set_input(self._h, name.encode(), C.c_void_p(host_ptr)))

    def set_input_device(self, name, dev_ptr):
        check(lib().fn2_net_set_input_device(self._h, name.encode(), C.c_void_p(dev_ptr)))

    def get_blob_ptr(self, name, host_ptr):
        check(lib().fn2_net_get_blob(self._h, name.encode(), C.c_void_p(host_ptr)))

    def get_blob_device(self, name, dev_ptr):
        check(lib().fn2_net_get_blob_device(self._h, name.encode(), C.c_void_p(dev_ptr)))

    def forward_async(self):
        check(lib().fn2_net_forward(self._h))

    # ---- gradients (pycaffe.py:127-175 Net.backward; net.cpp:640-655) ----
    def backward_async(self):
        check(lib().fn2_net_backward(self._h))

    def clear_param_diffs(self):
        check(lib().fn2_net_clear_param_diffs(self._h))

    def set_diff(self, name, array):
        a = np.ascontiguousarray(array, dtype=np.float32)
        if tuple(a.shape) != self.blobs[name].shape:
            raise Fn2Error("diff of %s has shape %s, net expects %s" % (name, a.shape, self.blobs[name].shape))
        check(lib().fn2_net_set_diff(self._h, name.encode(), a.ctypes.data_as(C.c_void_p)))

    def get_diff(self, name):
        out = np.empty(self.blobs[name].shape, np.float32)
        self.sync()
        check(lib().fn2_net_get_diff(self._h, name.encode(), out.ctypes.data_as(C.c_void_p)))
        return out

    def backward(self, **kwargs):
        """net.backward(flow=top_diff) -> {input blob name: gradient}; parameter gradients through net.param(layer, i, diff=True)."""
        for k, v in kwargs.items():
            self.set_diff(k, v)
        self.backward_async()
        self.sync()
        return {}

    def param(self, layer, index=0, diff=False):
        shp = (C.c_int * 4)()
        check(lib().fn2_net_param_shape(self._h, layer.encode(), int(index), shp))
        out = np.empty(tuple(int(x) for x in shp), np.float32)
        check(lib().fn2_net_get_param(self._h, layer.encode(), int(index), 1 if diff else 0, out.ctypes.data_as(C.c_void_p)))
        return out

    def param_diff_arena(self):
        p, n = C.c_void_p(), C.c_size_t()
        check(lib().fn2_net_param_diff_arena(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def layer_need_backward(self):
        return [bool(lib().fn2_net_layer_need_backward(self._h, i)) for i in range(len(self.layer_names))]

    @property
    def launches_per_backward(self):
        return int(lib().fn2_net_launches_per_backward(self._h))

    def sync(self):
        check(lib().fn2_net_sync(self._h))

    def forward(self, **kwargs):
        """net.forward(img0=..., img1=...) -> {output blob name: ndarray} (pycaffe.py:78-124)."""
        for k, v in kwargs.items():
            if k not in self.inputs:
                raise Fn2Error("Input blob arguments do not match net inputs.")
            self.set_input(k, v)
        self.forward_async()
        return {o: self.blobs[o].data for o in self.outputs}

    @property
    def stream(self):
        return lib().fn2_net_stream(self._h)

    def time_layers(self):
        ms = (C.c_float * len(self.layer_names))()
        check(lib().fn2_net_time_layers(self._h, ms))
        return list(zip(self.layer_names, self.layer_types, [float(x) for x in ms]))

    def layer_work(self):
        """[(name, type, algorithmic flops, algorithmic bytes)] per layer."""
        out = []
        for i, (n, t) in enumerate(zip(self.layer_names, self.layer_types)):
            f, b = C.c_double(), C.c_double()
            check(lib().fn2_net_layer_work(self._h, i, C.byref(f), C.byref(b)))
            out.append((n, t, f.value, b.value))
        return out

    @property
    def graph_active(self):
        return bool(lib().fn2_net_graph_active(self._h))

    @property
    def launches_per_forward(self):
        return int(lib().fn2_net_launches_per_forward(self._h))


# ------------------------------------------------------------------------------------------------
# .flo IO (src/caffe/util/output.cpp:16-64, scripts/run-flownet.py:100-126)
# ------------------------------------------------------------------------------------------------
def write_flo(path, flow_2hw):
    a = np.ascontiguousarray(flow_2hw, np.float32)
    assert a.ndim == 3 and a.shape[0] == 2
    check(lib().fn2_write_flo(path.encode(), a.ctypes.data_as(C.c_void_p), a.shape[1], a.shape[2]))


def read_flo(path):
    h, w = C.c_int(), C.c_int()
    check(lib().fn2_read_flo(path.encode(), None, C.byref(h), C.byref(w), 0))
    out = np.empty((2, h.value, w.value), np.float32)
    check(lib().fn2_read_flo(path.encode(), out.ctypes.data_as(C.c_void_p), C.byref(h), C.byref(w), out.size))
    return out


def image_to_blob(img_hwc_rgb_uint8):
    """HWC RGB uint8 -> (1,3,H,W) BGR float 0..255, scripts/run-flownet.py:30-35."""
    a = np.asarray(img_hwc_rgb_uint8)
    if a.ndim < 3:
        return a[np.newaxis, np.newaxis, :, :].astype(np.float32)
    return a[np.newaxis, :, :, :].transpose(0, 3, 1, 2)[:, [2, 1, 0], :, :].astype(np.float32)
