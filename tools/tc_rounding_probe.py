#!/usr/bin/env python
"""Does the B200 tensor-core FP32 accumulator round to nearest or toward zero?

TF32 GEMM (cuBLAS) on inputs that are exactly representable in TF32 and strictly positive, so every product
is exact and every partial sum is positive: round-toward-zero accumulation shows up as a NEGATIVE mean signed
error that grows with K; round-to-nearest as a zero-mean error.  Decides whether the tcgen05 conv engine may
chain its K loop inside the tensor core or must drain short chains and add in FP32 registers."""
import torch

torch.manual_seed(0)
dev = "cuda"


def tf32_exact(x):
    # keep 10 explicit mantissa bits
    i = x.view(torch.int32)
    return (i & ~0x1FFF).view(torch.float32)


for K in (64, 512, 4096, 16384):
    a = tf32_exact(torch.rand(512, K, device=dev) + 0.5)
    b = tf32_exact(torch.rand(K, 512, device=dev) + 0.5)
    ref = a.double() @ b.double()
    torch.backends.cuda.matmul.allow_tf32 = True
    tc = (a @ b).double()
    torch.backends.cuda.matmul.allow_tf32 = False
    fp = (a @ b).double()
    e_tc = ((tc - ref) / ref)
    e_fp = ((fp - ref) / ref)
    print("K=%6d  TF32 tensor core: mean rel err %+.3e (std %.2e)   FP32 SIMT: mean %+.3e (std %.2e)   [2^-24 = 5.96e-08]"
          % (K, e_tc.mean().item(), e_tc.std().item(), e_fp.mean().item(), e_fp.std().item()))
# zero-mean inputs: what matters for real layers
for K in (512, 4608):
    a = tf32_exact(torch.randn(512, K, device=dev))
    b = tf32_exact(torch.randn(K, 512, device=dev))
    ref = a.double() @ b.double()
    torch.backends.cuda.matmul.allow_tf32 = True
    tc = (a @ b).double()
    torch.backends.cuda.matmul.allow_tf32 = False
    fp = (a @ b).double()
    s = ref.abs().mean()
    print("K=%6d randn: TF32-TC |err|/mean|ref| mean %.3e, shrink %+0.3e   FP32: %.3e, shrink %+0.3e"
          % (K, ((tc - ref).abs().mean() / s).item(), (((tc - ref) * ref.sign()).mean() / s).item(),
             ((fp - ref).abs().mean() / s).item(), (((fp - ref) * ref.sign()).mean() / s).item()))
