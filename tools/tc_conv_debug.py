import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from flownet2_b200 import ops
from oracle import oracle as O
r = np.random.default_rng(0)
cl = torch.channels_last
cases = [  # N, Ci, H, W, Co, k, stride, pad, deconv
    (1, 32, 8, 16, 64, 1, 1, 0, False),
    (1, 64, 8, 16, 128, 3, 1, 1, False),
    (2, 96, 9, 13, 128, 3, 1, 1, False),
    (1, 64, 16, 20, 128, 5, 2, 2, False),
    (2, 473, 10, 14, 256, 3, 1, 1, False),
    (1, 256, 6, 7, 64, 4, 2, 1, True),
    (1, 1024, 3, 4, 512, 4, 2, 1, True),
    (2, 3, 20, 28, 64, 7, 2, 3, False),      # conv1 of FlowNetC: Ci = 3
    (1, 12, 20, 28, 64, 7, 2, 3, False),     # conv1 of the stacked FlowNetS: Ci = 12
    (1, 11, 18, 22, 64, 3, 1, 1, False),     # fusion conv0
    (1, 82, 16, 24, 16, 3, 1, 1, False),     # fusion interconv0: Co = 16
    (1, 162, 9, 12, 32, 3, 1, 1, False),     # fusion interconv1: Co = 32
    (1, 162, 9, 11, 16, 4, 2, 1, True),      # fusion deconv0: Co = 16
]
for case in cases:
    N, Ci, H, W, Co, k, s, p, dec = case
    x = r.standard_normal((N, Ci, H, W)).astype(np.float32)
    w = (r.standard_normal((Ci, Co, k, k) if dec else (Co, Ci, k, k)) * np.sqrt(2.0 / (Ci * k * k))).astype(np.float32)
    b = r.standard_normal(Co).astype(np.float32)
    fn = O.deconv_fwd if dec else O.conv_fwd
    want = O.relu(fn(x, w, b, s, p, f64acc=True), 0.1)
    Cp = (Ci + 31) // 32 * 32 if Ci >= 32 else (Ci + 3) // 4 * 4      # engine blob padding policy
    buf = torch.zeros((N, Cp, H, W), device="cuda").contiguous(memory_format=cl)
    buf[:, :Ci] = torch.from_numpy(x).cuda()
    tx = buf[:, :Ci]
    got = ops.conv2d(tx, torch.from_numpy(w).cuda(), torch.from_numpy(b).cuda(), s, p, dec, 0.1, 2)
    torch.cuda.synchronize()
    g = got.contiguous().cpu().numpy()
    simt = ops.conv2d(tx, torch.from_numpy(w).cuda(), torch.from_numpy(b).cuda(), s, p, dec, 0.1, 1).contiguous().cpu().numpy()
    sc = max(1.0, np.abs(want).max())
    print(case, "tc err %.3e  simt err %.3e  (scale %.2f)  mean signed tc %.2e" % (np.abs(g - want).max() / sc, np.abs(simt - want).max() / sc, sc,
          float(((g - want) * np.sign(want)).mean() / np.abs(want).mean())), flush=True)
# timing on a big layer: conv3_1-like and conv4_1-like
for (N, Ci, H, W, Co, k, s, p) in [(4, 473, 56, 128, 256, 3, 1, 1), (4, 64, 224, 512, 128, 5, 2, 2), (4, 12, 448, 1024, 64, 7, 2, 3),
                                   (4, 82, 448, 1024, 16, 3, 1, 1), (4, 1024, 7, 16, 1024, 3, 1, 1)]:
    Cp = (Ci + 31) // 32 * 32 if Ci >= 32 else (Ci + 3) // 4 * 4
    x = torch.randn(N, Cp, H, W, device="cuda").contiguous(memory_format=cl)[:, :Ci]
    w = torch.randn(Co, Ci, k, k, device="cuda") * 0.02
    b = torch.zeros(Co, device="cuda")
    for eng in (2, 1):
        for _ in range(2):
            ops.conv2d(x, w, b, s, p, False, 0.1, eng)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # time only the forward call: pack once
        import ctypes as C
        from flownet2_b200 import lib, check, fn2_conv_desc
        l = lib()
        d = fn2_conv_desc(Ci, Co, k, k, s, s, p, p, 0, 1, 1, 0.1, eng)
        nf = C.c_size_t(); check(l.fn2_conv_packed_floats(C.byref(d), Ci, C.byref(nf)))
        packed = torch.empty(nf.value, device="cuda")
        check(l.fn2_conv_pack_weights(C.byref(d), Ci, C.c_void_p(w.data_ptr()), C.c_void_p(packed.data_ptr()), None))
        Ho = (H + 2 * p - k) // s + 1; Wo = (W + 2 * p - k) // s + 1
        out = torch.empty(N, Co, Ho, Wo, device="cuda").contiguous(memory_format=cl)
        dx, do = ops.desc(x), ops.desc(out)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for _ in range(3):
            check(l.fn2_conv_forward(C.byref(d), C.byref(dx), C.c_void_p(packed.data_ptr()), C.c_void_p(b.data_ptr()), C.byref(do), None, 0, st))
        e0.record()
        for _ in range(10):
            check(l.fn2_conv_forward(C.byref(d), C.byref(dx), C.c_void_p(packed.data_ptr()), C.c_void_p(b.data_ptr()), C.byref(do), None, 0, st))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2.0 * N * Co * Ho * Wo * Ci * k * k
        print("engine %d  %s: %.3f ms  %.1f TFLOP/s" % (eng, (N, Ci, H, W, Co, k, s), ms, fl / ms / 1e9), flush=True)
