"""Times the four fusion-tail layers of FlowNet2 @1024x436 b4 (few output channels at full / half resolution) on the
tcgen05 engines: FN2_TN=1 (default) taps-on-N engine, FN2_TN=0 the per-tap engine of round 1."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C, torch
from flownet2_b200 import ops, lib, check, fn2_conv_desc
cl = torch.channels_last
l = lib()
LAYERS = [("fuse_interconv0", 4, 82, 448, 1024, 16, 3, 1, 1, 0), ("fuse_deconv0", 4, 162, 224, 512, 16, 4, 2, 1, 1),
          ("fuse_interconv1", 4, 162, 224, 512, 32, 3, 1, 1, 0), ("fuse_deconv1", 4, 128, 112, 256, 32, 4, 2, 1, 1)]
for (name, N, Ci, H, W, Co, k, s, p, dc) in LAYERS:
    Cp = (Ci + 31) // 32 * 32
    x = torch.randn(N, Cp, H, W, device="cuda").contiguous(memory_format=cl)[:, :Ci]
    w = torch.randn((Ci, Co, k, k) if dc else (Co, Ci, k, k), device="cuda") * 0.02
    b = torch.zeros(Co, device="cuda")
    d = fn2_conv_desc(Ci, Co, k, k, s, s, p, p, dc, 1, 1, 0.1, 2, 0)
    nf = C.c_size_t(); check(l.fn2_conv_packed_floats(C.byref(d), Cp, C.byref(nf)))
    packed = torch.empty(nf.value, device="cuda")
    check(l.fn2_conv_pack_weights(C.byref(d), Cp, C.c_void_p(w.data_ptr()), C.c_void_p(packed.data_ptr()), None))
    ho, wo = C.c_int(), C.c_int(); check(l.fn2_conv_out_shape(C.byref(d), H, W, C.byref(ho), C.byref(wo)))
    out = torch.empty(N, Co, ho.value, wo.value, device="cuda").contiguous(memory_format=cl)
    dx, do = ops.desc(x), ops.desc(out)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    wb = C.c_size_t(); check(l.fn2_conv_workspace_bytes(C.byref(d), N, H, W, C.byref(wb)))
    wsbuf = torch.empty(max(wb.value, 4) // 4, device="cuda")
    run = lambda: check(l.fn2_conv_forward(C.byref(d), C.byref(dx), C.c_void_p(packed.data_ptr()), C.c_void_p(b.data_ptr()), C.byref(do),
                                           C.c_void_p(wsbuf.data_ptr()), wb.value, st))
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gf = 2.0 * N * Co * k * k * Ci * (ho.value * wo.value if not dc else H * W) / 1e9
    print("FN2_TN=%s %-16s %.3f ms  %.1f TFLOP/s algorithmic" % (os.environ.get("FN2_TN", "1"), name, ms, gf / ms), flush=True)
