// Streaming (HBM-bound) layers of the FlowNet2 forward path: FlowWarp, Resample,
// DataAugmentation kernels, ReLU, Eltwise, ChannelNorm, strided copy/fill.
//
// This translation unit is compiled with -fmad=false: every expression below is evaluated
// with separately rounded multiplies and adds in the order written, which is also how the CPU
// oracle (oracle/oracle.c, -ffp-contract=off) evaluates the reference arithmetic, so these
// layers are compared bit-for-bit.  They are bandwidth-bound; FMA contraction buys nothing.
//
// All kernels take strided tensor views (fn2::T4) so the same code serves the reference's NCHW
// blobs (drop-in use) and the engine's NHWC activations / concat views.
#include <math.h>
#include <cfloat>

#include "fn2_common.cuh"

namespace fn2 {

// ---------------------------------------------------------------------------------------------
// FlowWarp forward.  Reference: flow_warp_layer.cpp:57-117 (CPU), flow_warp_layer.cu:59-122.
// One thread per output pixel; the reference's NCHW->NHWC transpose pass (.cu:24-52) is not
// needed: taps are read straight from the strided view.
// ---------------------------------------------------------------------------------------------
__global__ void flow_warp_fwd_kernel(T4 img, T4 flow, T4 out, float fill, int vec4) {
    const long long total = (long long)out.n * out.h * out.w;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % out.w);
        const int y = (int)((idx / out.w) % out.h);
        const int n = (int)(idx / ((long long)out.w * out.h));
        const float fx = flow.p[flow.off(n, 0, y, x)];
        const float fy = flow.p[flow.off(n, 1, y, x)];
        const float x2 = (float)x + fx;
        const float y2 = (float)y + fy;
        const int width = img.w, height = img.h;
        if (x2 >= 0 && y2 >= 0 && x2 < width && y2 < height) {      // flow_warp_layer.cpp:86
            const int ix2_L = (int)x2;
            const int iy2_T = (int)y2;
            const int ix2_R = min(ix2_L + 1, width - 1);
            const int iy2_B = min(iy2_T + 1, height - 1);
            const float alpha = x2 - ix2_L;
            const float beta = y2 - iy2_T;
            const float cTL = (1 - alpha) * (1 - beta);
            const float cTR = alpha * (1 - beta);
            const float cBL = (1 - alpha) * beta;
            const float cBR = alpha * beta;
            const long long oTL = img.off(n, 0, iy2_T, ix2_L), oTR = img.off(n, 0, iy2_T, ix2_R);
            const long long oBL = img.off(n, 0, iy2_B, ix2_L), oBR = img.off(n, 0, iy2_B, ix2_R);
            if (vec4) {
                // channel-fast image with <= 4 channels in 16-byte pixels: one 128-bit load per tap (same arithmetic per channel)
                const float4 TL = __ldg(reinterpret_cast<const float4*>(img.p + oTL)), TR = __ldg(reinterpret_cast<const float4*>(img.p + oTR));
                const float4 BL = __ldg(reinterpret_cast<const float4*>(img.p + oBL)), BR = __ldg(reinterpret_cast<const float4*>(img.p + oBR));
                const float tl[4] = {TL.x, TL.y, TL.z, TL.w}, tr[4] = {TR.x, TR.y, TR.z, TR.w};
                const float bl[4] = {BL.x, BL.y, BL.z, BL.w}, br[4] = {BR.x, BR.y, BR.z, BR.w};
#pragma unroll
                for (int c = 0; c < 4; c++)
                    if (c < img.c) out.p[out.off(n, c, y, x)] = ((cTL * tl[c] + cTR * tr[c]) + cBL * bl[c]) + cBR * br[c];
                continue;
            }
            for (int c = 0; c < img.c; c++) {
                const float TL = __ldg(img.p + oTL + c * img.sc);
                const float TR = __ldg(img.p + oTR + c * img.sc);
                const float BL = __ldg(img.p + oBL + c * img.sc);
                const float BR = __ldg(img.p + oBR + c * img.sc);
                out.p[out.off(n, c, y, x)] = ((cTL * TL + cTR * TR) + cBL * BL) + cBR * BR;
            }
        } else {
            for (int c = 0; c < img.c; c++) out.p[out.off(n, c, y, x)] = fill;
        }
    }
}

// FlowWarp backward.  Reference: flow_warp_layer.cu:170-229 / flow_warp_layer.cpp:120-198.
// image_diff must be zero-filled by the caller side (done in the entry point).  Scatter via
// atomicAdd exactly like the reference (order-nondeterministic in the last ulp).
__global__ void flow_warp_bwd_kernel(T4 img, T4 flow, T4 wdiff, T4 idiff, T4 fdiff) {
    const long long total = (long long)img.n * img.h * img.w;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % img.w);
        const int y = (int)((idx / img.w) % img.h);
        const int n = (int)(idx / ((long long)img.w * img.h));
        const float x2 = (float)x + flow.p[flow.off(n, 0, y, x)];
        const float y2 = (float)y + flow.p[flow.off(n, 1, y, x)];
        const int width = img.w, height = img.h;
        float du = 0.f, dv = 0.f;
        if (x2 >= 0 && y2 >= 0 && x2 < width && y2 < height) {
            const int L = (int)x2, T = (int)y2;
            const int R = min(L + 1, width - 1), B = min(T + 1, height - 1);
            const float alpha = x2 - L, beta = y2 - T;
            const float gy = B - y2, gx = R - x2;
            for (int c = 0; c < img.c; c++) {
                const float wd = wdiff.p[wdiff.off(n, c, y, x)];
                atomicAdd(idiff.p + idiff.off(n, c, T, L), wd * (1 - alpha) * (1 - beta));
                atomicAdd(idiff.p + idiff.off(n, c, T, R), wd * alpha * (1 - beta));
                atomicAdd(idiff.p + idiff.off(n, c, B, L), wd * (1 - alpha) * beta);
                atomicAdd(idiff.p + idiff.off(n, c, B, R), wd * alpha * beta);
                const float TL = img.p[img.off(n, c, T, L)], TR = img.p[img.off(n, c, T, R)];
                const float BL = img.p[img.off(n, c, B, L)], BR = img.p[img.off(n, c, B, R)];
                float tu = 0; tu += gy * (TR - TL); tu += (1 - gy) * (BR - BL);
                float tv = 0; tv += gx * (BL - TL); tv += (1 - gx) * (BR - TR);
                du += wd * tu;
                dv += wd * tv;
            }
        }
        fdiff.p[fdiff.off(n, 0, y, x)] = du;
        fdiff.p[fdiff.off(n, 1, y, x)] = dv;
    }
}

// ---------------------------------------------------------------------------------------------
// Warp block between two stacked networks (FlowNet2-CSS / FlowNet2): the chain
//   Resample(LINEAR, up-sampling) -> FlowWarp -> Eltwise(img0 - warped) -> ChannelNorm, and Eltwise(coeff * flow)
// as ONE pass: one thread per full-resolution pixel computes the interpolated flow, samples the second image, and writes all
// five tops (most of them channel ranges of the next network's input concat).  Every value goes through exactly the
// expressions of the stand-alone kernels below (resample_kernel<2>, flow_warp_fwd_kernel, eltwise_sum_kernel with coefficients
// (1, -1) resp. (coeff), channel_norm_kernel; this file is compiled -fmad=false), so the tops are bit-identical to the layer
// chain's.  References: resample_layer.cu:40-95, flow_warp_layer.cu:59-122, eltwise_layer.cu:36-60, channel_norm_layer.cu:17-30.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float triangleCoeffW(float x) {
    if (-1 <= x && x < 0) return x + 1;
    if (0 <= x && x <= 1) return 1 - x;
    return 0;
}
__global__ void warp_block_kernel(T4 fin, T4 img0, T4 img1, T4 ffull, T4 warped, T4 err, T4 errn, T4 fscaled, float fx, float fy,
                                  float coeff, float fill, int vec4) {
    const long long total = (long long)ffull.n * ffull.h * ffull.w;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int x_out = (int)(idx % ffull.w);
        const int y_out = (int)((idx / ffull.w) % ffull.h);
        const int n = (int)(idx / ((long long)ffull.w * ffull.h));
        // ---- Resample (LINEAR, no antialias: up-sampling), 2 channels: resample_kernel<2> with the window of <= 8 columns
        const float x_in = (x_out * fx + fy / 2.0f) - 0.5f;
        const float y_in = (y_out * fy + fx / 2.0f) - 0.5f;
        const int x_in_round = (int)roundf(x_in);
        const int y_in_round = (int)roundf(y_in);
        const float ax = 1.0f / 1.0f, ay = 1.0f / 1.0f;
        const int rx = (fx < 1.0f) ? 2 : (int)ceilf(2.0f / ax);
        const int ry = (fy < 1.0f) ? 2 : (int)ceilf(2.0f / ay);
        const int xl = max(x_in_round - rx, (int)floorf(x_in - 1.0f / ax)), xh = min(x_in_round + rx, (int)ceilf(x_in + 1.0f / ax));
        const int yl = max(y_in_round - ry, (int)floorf(y_in - 1.0f / ay)), yh = min(y_in_round + ry, (int)ceilf(y_in + 1.0f / ay));
        float tx_[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float dx = x_in - (xl + k);
            tx_[k] = (ax * triangleCoeffW(ax * dx)) * ay;
        }
        float sum0 = 0.f, sum1 = 0.f, wsum = 0.f;
        for (int y = yl; y <= yh; y++) {
            if (y < 0 || y >= fin.h) continue;
            const float dy = y_in - y;
            const float cy = triangleCoeffW(ay * dy);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int x = xl + k;
                if (x > xh) break;
                if (x < 0 || x >= fin.w) continue;
                const float w = tx_[k] * cy;
                if (w == 0.f) continue;
                const float* ip = fin.p + fin.off(n, 0, y, x);
                sum0 = sum0 + w * __ldg(ip);
                sum1 = sum1 + w * __ldg(ip + fin.sc);
                wsum += w;
            }
        }
        const float f0 = (!wsum) ? 0 : (sum0 / wsum), f1 = (!wsum) ? 0 : (sum1 / wsum);
        ffull.p[ffull.off(n, 0, y_out, x_out)] = f0;
        ffull.p[ffull.off(n, 1, y_out, x_out)] = f1;
        // ---- Eltwise(coeff): acc = 0 + coeff * v
        fscaled.p[fscaled.off(n, 0, y_out, x_out)] = 0.f + coeff * f0;
        fscaled.p[fscaled.off(n, 1, y_out, x_out)] = 0.f + coeff * f1;
        // ---- FlowWarp of img1 (flow_warp_fwd_kernel)
        const float x2 = (float)x_out + f0;
        const float y2 = (float)y_out + f1;
        const int width = img1.w, height = img1.h;
        float wv[4] = {fill, fill, fill, fill};
        if (x2 >= 0 && y2 >= 0 && x2 < width && y2 < height) {
            const int ix2_L = (int)x2;
            const int iy2_T = (int)y2;
            const int ix2_R = min(ix2_L + 1, width - 1);
            const int iy2_B = min(iy2_T + 1, height - 1);
            const float alpha = x2 - ix2_L;
            const float beta = y2 - iy2_T;
            const float cTL = (1 - alpha) * (1 - beta);
            const float cTR = alpha * (1 - beta);
            const float cBL = (1 - alpha) * beta;
            const float cBR = alpha * beta;
            const long long oTL = img1.off(n, 0, iy2_T, ix2_L), oTR = img1.off(n, 0, iy2_T, ix2_R);
            const long long oBL = img1.off(n, 0, iy2_B, ix2_L), oBR = img1.off(n, 0, iy2_B, ix2_R);
            if (vec4) {
                const float4 TL = __ldg(reinterpret_cast<const float4*>(img1.p + oTL)), TR = __ldg(reinterpret_cast<const float4*>(img1.p + oTR));
                const float4 BL = __ldg(reinterpret_cast<const float4*>(img1.p + oBL)), BR = __ldg(reinterpret_cast<const float4*>(img1.p + oBR));
                const float tl[4] = {TL.x, TL.y, TL.z, TL.w}, tr[4] = {TR.x, TR.y, TR.z, TR.w};
                const float bl[4] = {BL.x, BL.y, BL.z, BL.w}, br[4] = {BR.x, BR.y, BR.z, BR.w};
#pragma unroll
                for (int c = 0; c < 4; c++) wv[c] = ((cTL * tl[c] + cTR * tr[c]) + cBL * bl[c]) + cBR * br[c];
            } else {
                for (int c = 0; c < img1.c; c++) {
                    const float TL = __ldg(img1.p + oTL + c * img1.sc);
                    const float TR = __ldg(img1.p + oTR + c * img1.sc);
                    const float BL = __ldg(img1.p + oBL + c * img1.sc);
                    const float BR = __ldg(img1.p + oBR + c * img1.sc);
                    wv[c] = ((cTL * TL + cTR * TR) + cBL * BL) + cBR * BR;
                }
            }
        }
        // ---- Eltwise(1, -1): acc = (0 + 1 * a) + (-1) * b, then ChannelNorm
        float norm = 0;
        for (int c = 0; c < img1.c; c++) {
            warped.p[warped.off(n, c, y_out, x_out)] = wv[c];
            float acc = 0.f + 1.0f * img0.p[img0.off(n, c, y_out, x_out)];
            acc = acc + (-1.0f) * wv[c];
            err.p[err.off(n, c, y_out, x_out)] = acc;
            norm = norm + acc * acc;
        }
        errn.p[errn.off(n, 0, y_out, x_out)] = sqrtf(norm);
    }
}

// ---------------------------------------------------------------------------------------------
// Resample.  Reference: resample_layer.cu:14-33 (filters), :40-95 InterpolationKernel,
// :98-125 NearestNeighborKernel.  Including the swapped half-pixel offsets (:62-63).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float bicubicCoeff(float x_) {
    float x = fabsf(x_);
    if (x <= 1.0f) return x * x * (1.5f * x - 2.5f) + 1.0f;
    else if (x < 2.0f) return x * (x * (-0.5f * x + 2.5f) - 4.0f) + 2.0f;
    else return 0.0f;
}
__device__ __forceinline__ float triangleCoeff(float x) {
    if (-1 <= x && x < 0) return x + 1;
    if (0 <= x && x <= 1) return 1 - x;
    return 0;
}

// One thread per output pixel, all channels: the filter weights depend on the pixel only, and every channel accumulates
// its taps in the reference's order (y outer, x inner), so each channel's result is bit-identical to the per-element
// formulation.  Taps whose weight is exactly zero are skipped: sum + 0*v == sum and wsum + 0 == wsum for finite v (the
// triangle filter of an up-sampling has 21 zero taps out of 25).
template <int TYPE>   // 1 nearest, 2 linear(triangle), 3 cubic
__global__ void resample_kernel(T4 in, T4 out, float fx, float fy, int antialias) {
    const long long total = (long long)out.n * out.h * out.w;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x_out = (int)(idx % out.w);
        const int y_out = (int)((idx / out.w) % out.h);
        const int n = (int)(idx / ((long long)out.w * out.h));
        const float x_in = (x_out * fx + fy / 2.0f) - 0.5f;
        const float y_in = (y_out * fy + fx / 2.0f) - 0.5f;
        const int x_in_round = (int)roundf(x_in);
        const int y_in_round = (int)roundf(y_in);
        if (TYPE == 1) {
            // the reference does not clamp (resample_layer.cu:120-123) and can read out of
            // bounds; clamp instead (documented deviation)
            const int xr = min(max(x_in_round, 0), in.w - 1);
            const int yr = min(max(y_in_round, 0), in.h - 1);
            for (int c = 0; c < out.c; c++) out.p[out.off(n, c, y_out, x_out)] = in.p[in.off(n, c, yr, xr)];
            continue;
        }
        const int kernel_width = (TYPE == 3) ? 4 : 2;
        const float ax = 1.0f / (antialias ? fx : 1.0f);
        const float ay = 1.0f / (antialias ? fy : 1.0f);
        const int rx = (fx < 1.0f) ? 2 : (int)ceilf((float)kernel_width / ax);
        const int ry = (fy < 1.0f) ? 2 : (int)ceilf((float)kernel_width / ay);
        // the filters vanish outside |ax*dx| < support: only the taps of the reference's window that can carry weight are
        // visited (every skipped tap has weight exactly 0, see above)
        const float sup = (TYPE == 3) ? 2.0f : 1.0f;
        const int xl = max(x_in_round - rx, (int)floorf(x_in - sup / ax)), xh = min(x_in_round + rx, (int)ceilf(x_in + sup / ax));
        const int yl = max(y_in_round - ry, (int)floorf(y_in - sup / ay)), yh = min(y_in_round + ry, (int)ceilf(y_in + sup / ay));
        // x factors ((ax * cx) * ay, the first two products of the reference's weight expression) once per pixel when the
        // window is narrow (always, except for strong down-sampling)
        constexpr int MAXT = 8;
        float tx_[MAXT];
        const bool pre = xh - xl < MAXT;
        if (pre) {
#pragma unroll
            for (int k = 0; k < MAXT; k++) {
                const float dx = x_in - (xl + k);
                const float cx = (TYPE == 3) ? bicubicCoeff(ax * dx) : triangleCoeff(ax * dx);
                tx_[k] = (ax * cx) * ay;
            }
        }
        for (int c0 = 0; c0 < out.c; c0 += 4) {
            const int nc = min(4, out.c - c0);
            float sum[4] = {0.f, 0.f, 0.f, 0.f};
            float wsum = 0;
            for (int y = yl; y <= yh; y++) {
                if (y < 0 || y >= in.h) continue;
                const float dy = y_in - y;
                const float cy = (TYPE == 3) ? bicubicCoeff(ay * dy) : triangleCoeff(ay * dy);
#pragma unroll
                for (int k = 0; k < MAXT; k++) {
                    const int x = xl + k;
                    if (!pre || x > xh) break;
                    if (x < 0 || x >= in.w) continue;
                    const float w = tx_[k] * cy;
                    if (w == 0.f) continue;
                    const float* ip = in.p + in.off(n, c0, y, x);
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (j < nc) sum[j] = sum[j] + w * __ldg(ip + (long long)j * in.sc);
                    wsum += w;
                }
                if (!pre)
                for (int x = xl; x <= xh; x++) {
                    if (x < 0 || x >= in.w) continue;
                    const float dx = x_in - x;
                    const float cx = (TYPE == 3) ? bicubicCoeff(ax * dx) : triangleCoeff(ax * dx);
                    const float w = (ax * cx) * ay * cy;
                    if (w == 0.f) continue;
                    const float* ip = in.p + in.off(n, c0, y, x);
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (j < nc) sum[j] = sum[j] + w * __ldg(ip + (long long)j * in.sc);
                    wsum += w;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (j < nc) out.p[out.off(n, c0 + j, y_out, x_out)] = (!wsum) ? 0 : (sum[j] / wsum);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// DataAugmentation device kernels.
// SpatialAugmentation: data_augmentation_layer.cu:25-70.  The reference forms the flat source
// index in float arithmetic (:52), exact while the blob has < 2^24 elements; integer arithmetic
// here (identical in that range).  The min(idx, src_count) clamps of :53-55 can never trigger
// for width,height >= 2 because of the dim-1.05 clamp.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float clampf(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }

__global__ void spatial_aug_kernel(T4 src, T4 dst, const float* __restrict__ mats) {
    const long long total = (long long)dst.n * dst.h * dst.w;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % dst.w);
        const int y = (int)((idx / dst.w) % dst.h);
        const int n = (int)(idx / ((long long)dst.w * dst.h));
        const float* m = mats + 6 * n;
        float xpos = (x * m[0] + y * m[2]) + m[4];
        float ypos = (x * m[1] + y * m[3]) + m[5];
        xpos = clampf(xpos, 0.0f, (float)(src.w) - 1.05f);
        ypos = clampf(ypos, 0.0f, (float)(src.h) - 1.05f);
        const float tlx = floorf(xpos);
        const float tly = floorf(ypos);
        const int ix = (int)tlx, iy = (int)tly;
        const float xdist = xpos - tlx;
        const float ydist = ypos - tly;
        const float cTL = (1 - xdist) * (1 - ydist), cBR = xdist * ydist;
        const float cBL = (1 - xdist) * ydist, cTR = xdist * (1 - ydist);
        const long long oTL = src.off(n, 0, iy, ix);
        for (int c = 0; c < src.c; c++) {
            const float* s = src.p + oTL + c * src.sc;
            const float TL = __ldg(s), TR = __ldg(s + src.sw);
            const float BL = __ldg(s + src.sh), BR = __ldg(s + src.sh + src.sw);
            // term order TL, BR, BL, TR (data_augmentation_layer.cu:62-65)
            dst.p[dst.off(n, c, y, x)] = ((cTL * TL + cBR * BR) + cBL * BL) + cTR * TR;
        }
    }
}

// ColorContrastAugmentation: data_augmentation_layer.cu:73-117 (in place, 3 channels).
__global__ void color_contrast_kernel(T4 d, const float* __restrict__ chroma, float max_multiplier) {
    const long long total = (long long)d.n * d.h * d.w;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % d.w);
        const int y = (int)((idx / d.w) % d.h);
        const int n = (int)(idx / ((long long)d.w * d.h));
        const float* ch = chroma + 6 * n;
        float rgb[3];
        float mean_in = 0, mean_out = 0;
        for (int c = 0; c < 3; c++) {
            rgb[c] = d.p[d.off(n, c, y, x)];
            mean_in += rgb[c];
            rgb[c] *= ch[3 + c];
            mean_out += rgb[c];
        }
        const float brightness_coeff = mean_in / (mean_out + 0.01f);
        for (int c = 0; c < 3; c++) {
            float v = clampf(rgb[c] * brightness_coeff, 0.f, 1.f);
            v = powf(v, ch[0]);
            v = v + ch[1];
            v = 0.5f + (v - 0.5f) * ch[2];
            d.p[d.off(n, c, y, x)] = clampf(v, 0.f, max_multiplier);
        }
    }
}

// Running mean update: data_augmentation_layer.cu:600-607 (scal, axpy per sample, scal).
__global__ void mean_update_kernel(T4 top, T4 mean, float num_iter) {
    const long long total = (long long)top.c * top.h * top.w;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % top.w);
        const int y = (int)((idx / top.w) % top.h);
        const int c = (int)(idx / ((long long)top.w * top.h));
        float* mp = mean.p + mean.off(0, c, y, x);
        float m = *mp * (num_iter - 1.0f);
        const float inv = 1.0f / (float)top.n;
        for (int n = 0; n < top.n; n++) m = m + inv * top.p[top.off(n, c, y, x)];
        *mp = m * (1.0f / num_iter);
    }
}
// Per-channel average of the mean image (caffe_gpu_gemv with ones, :608); one CTA per channel,
// double accumulation (the cuBLAS order is unpinned).
__global__ void mean_per_channel_kernel(T4 mean, float* __restrict__ mean_pc) {
    const int c = blockIdx.x;
    const int area = mean.h * mean.w;
    double acc = 0;
    for (int i = threadIdx.x; i < area; i += blockDim.x)
        acc += (double)mean.p[mean.off(0, c, i / mean.w, i % mean.w)];
    __shared__ double sh[256];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) mean_pc[c] = (float)((1.0 / (double)area) * sh[0]);
}
// Mean subtraction :610-634.
__global__ void mean_subtract_kernel(T4 top, T4 mean, const float* __restrict__ mean_pc, int per_pixel) {
    const long long total = top.count();
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % top.w);
        long long r = idx / top.w;
        const int c = (int)(r % top.c); r /= top.c;
        const int y = (int)(r % top.h);
        const int n = (int)(r / top.h);
        const float m = per_pixel ? mean.p[mean.off(0, c, y, x)] : mean_pc[c];
        float* t = top.p + top.off(n, c, y, x);
        *t = *t - m;
    }
}

// ---------------------------------------------------------------------------------------------
// Glue.
// ---------------------------------------------------------------------------------------------
// ReLU: relu_layer.cu:9-14  x > 0 ? x : x*slope
__global__ void relu_kernel(T4 in, T4 out, float slope) {
    const long long total = out.count();
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % out.c);
        long long r = idx / out.c;
        const int x = (int)(r % out.w); r /= out.w;
        const int y = (int)(r % out.h);
        const int n = (int)(r / out.h);
        const float v = in.p[in.off(n, c, y, x)];
        out.p[out.off(n, c, y, x)] = v > 0 ? v : v * slope;
    }
}

struct EltArgs { T4 b[4]; float coeff[4]; int nb; };
// Eltwise SUM: eltwise_layer.cpp:59-65 (zero, then axpy in bottom order).
__global__ void eltwise_sum_kernel(EltArgs a, T4 out) {
    const long long total = out.count();
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % out.w);
        long long r = idx / out.w;
        const int c = (int)(r % out.c); r /= out.c;
        const int y = (int)(r % out.h);
        const int n = (int)(r / out.h);
        float acc = 0.f;
        for (int b = 0; b < a.nb; b++) acc = acc + a.coeff[b] * a.b[b].p[a.b[b].off(n, c, y, x)];
        out.p[out.off(n, c, y, x)] = acc;
    }
}

// ChannelNorm: channel_norm_layer.cu:17-30.
__global__ void channel_norm_kernel(T4 in, T4 out, int vec4) {
    const long long total = (long long)in.n * in.h * in.w;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % in.w);
        const int y = (int)((idx / in.w) % in.h);
        const int n = (int)(idx / ((long long)in.w * in.h));
        float norm = 0;
        if (vec4) {
            const float4 q = *reinterpret_cast<const float4*>(in.p + in.off(n, 0, y, x));
            const float v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int c = 0; c < 4; c++) if (c < in.c) norm = norm + v[c] * v[c];
        } else {
            for (int c = 0; c < in.c; c++) {
                const float v = in.p[in.off(n, c, y, x)];
                norm = norm + v * v;
            }
        }
        out.p[out.off(n, 0, y, x)] = sqrtf(norm);
    }
}

// Strided copy (Concat, layout conversion).  Two index decodings so that either side can be
// the coalesced one: SRC_CFAST decodes channel-fastest (good when src or dst is NHWC).
template <bool CFAST>
__global__ void copy_kernel(T4 src, T4 dst) {
    const long long total = dst.count();
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int n, c, y, x;
        if (CFAST) {
            c = (int)(idx % dst.c);
            long long r = idx / dst.c;
            x = (int)(r % dst.w); r /= dst.w;
            y = (int)(r % dst.h);
            n = (int)(r / dst.h);
        } else {
            x = (int)(idx % dst.w);
            long long r = idx / dst.w;
            y = (int)(r % dst.h); r /= dst.h;
            c = (int)(r % dst.c);
            n = (int)(r / dst.c);
        }
        dst.p[dst.off(n, c, y, x)] = src.p[src.off(n, c, y, x)];
    }
}
// Tiled transpose copy between a channel-planar side and a channel-fast side: 32 pixels x 32
// channels per tile through shared memory so that both global sides are coalesced.
__global__ void copy_transpose_kernel(T4 src, T4 dst, int src_cfast) {
    __shared__ float tile[32][33];          // [channel][pixel]
    const long long hw = (long long)dst.h * dst.w;
    const int n = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        // the fast thread index follows the source's contiguous dimension
        const int ci = src_cfast ? tx : j;
        const int pi = src_cfast ? j : tx;
        const long long p = p0 + pi;
        if (c0 + ci < dst.c && p < hw)
            tile[ci][pi] = src.p[src.off(n, c0 + ci, (int)(p / dst.w), (int)(p % dst.w))];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        // ... and the destination's contiguous dimension on the way out
        const int ci = src_cfast ? j : tx;
        const int pi = src_cfast ? tx : j;
        const long long p = p0 + pi;
        if (c0 + ci < dst.c && p < hw)
            dst.p[dst.off(n, c0 + ci, (int)(p / dst.w), (int)(p % dst.w))] = tile[ci][pi];
    }
}

__global__ void fill_kernel(T4 dst, float v) {
    const long long total = dst.count();
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % dst.c);
        long long r = idx / dst.c;
        const int x = (int)(r % dst.w); r /= dst.w;
        const int y = (int)(r % dst.h);
        const int n = (int)(r / dst.h);
        dst.p[dst.off(n, c, y, x)] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// Channel-fast ("NHWC") variants of the element-wise kernels: one thread per (pixel, group of 4 channels), 128-bit
// accesses where the view allows them.  Same per-element arithmetic as the generic kernels above (bit-identical).
// vr: the view may be READ as float4 groups (16-byte aligned pixels; lanes beyond C are padding or a neighbour's
// channels -- read, never used).  WRITES never touch a lane beyond C: a view cannot tell whether those lanes are its own
// padding or a sibling's channels inside a zero-copy Concat parent (a 2-channel child of a 2+2 parent has the same strides as
// a dense 2-channel blob), so the tail group is stored lane by lane; padding stays zero from the allocation.
// ---------------------------------------------------------------------------------------------
struct PxView { T4 t; int vr; };
static PxView px_view(const T4& t) {
    PxView v; v.t = t;
    v.vr = t.sc == 1 && !((uintptr_t)t.p & 15) && !(t.sw & 3) && !(t.sh & 3) && !(t.sn & 3) && t.sw >= (t.c + 3) / 4 * 4;
    return v;
}
__device__ __forceinline__ float4 px_load(const PxView& v, long long o, int c0) {
    if (v.vr) return *reinterpret_cast<const float4*>(v.t.p + o + c0);
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c0 + 0 < v.t.c) r.x = v.t.p[o + c0 + 0];
    if (c0 + 1 < v.t.c) r.y = v.t.p[o + c0 + 1];
    if (c0 + 2 < v.t.c) r.z = v.t.p[o + c0 + 2];
    if (c0 + 3 < v.t.c) r.w = v.t.p[o + c0 + 3];
    return r;
}
__device__ __forceinline__ void px_store(const PxView& v, long long o, int c0, float4 r) {
    if (v.vr && c0 + 4 <= v.t.c) { *reinterpret_cast<float4*>(v.t.p + o + c0) = r; return; }
    if (v.vr && c0 + 2 == v.t.c) { *reinterpret_cast<float2*>(v.t.p + o + c0) = make_float2(r.x, r.y); return; }   // 2-channel flows
    if (c0 + 0 < v.t.c) v.t.p[o + c0 + 0] = r.x;
    if (c0 + 1 < v.t.c) v.t.p[o + c0 + 1] = r.y;
    if (c0 + 2 < v.t.c) v.t.p[o + c0 + 2] = r.z;
    if (c0 + 3 < v.t.c) v.t.p[o + c0 + 3] = r.w;
}
#define FN2_PX_DECODE(T)                                                                       \
        const int g = (int)(idx % G);                                                          \
        long long r_ = idx / G;                                                                \
        const int x = (int)(r_ % (T).w); r_ /= (T).w;                                          \
        const int y = (int)(r_ % (T).h);                                                       \
        const int n = (int)(r_ / (T).h);                                                       \
        const int c0 = 4 * g;

struct EltPx { PxView b[4]; float coeff[4]; int nb; };
__global__ void eltwise_sum_px_kernel(EltPx a, PxView out) {
    const int G = (out.t.c + 3) / 4;
    const long long total = (long long)out.t.n * out.t.h * out.t.w * G;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        FN2_PX_DECODE(out.t)
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int b = 0; b < a.nb; b++) {
            const float4 v = px_load(a.b[b], a.b[b].t.off(n, 0, y, x), c0);
            acc.x = acc.x + a.coeff[b] * v.x; acc.y = acc.y + a.coeff[b] * v.y;
            acc.z = acc.z + a.coeff[b] * v.z; acc.w = acc.w + a.coeff[b] * v.w;
        }
        px_store(out, out.t.off(n, 0, y, x), c0, acc);
    }
}
__global__ void relu_px_kernel(PxView in, PxView out, float slope) {
    const int G = (out.t.c + 3) / 4;
    const long long total = (long long)out.t.n * out.t.h * out.t.w * G;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        FN2_PX_DECODE(out.t)
        float4 v = px_load(in, in.t.off(n, 0, y, x), c0);
        v.x = v.x > 0 ? v.x : v.x * slope; v.y = v.y > 0 ? v.y : v.y * slope;
        v.z = v.z > 0 ? v.z : v.z * slope; v.w = v.w > 0 ? v.w : v.w * slope;
        px_store(out, out.t.off(n, 0, y, x), c0, v);
    }
}
__global__ void copy_px_kernel(PxView src, PxView dst) {
    const int G = (dst.t.c + 3) / 4;
    const long long total = (long long)dst.t.n * dst.t.h * dst.t.w * G;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        FN2_PX_DECODE(dst.t)
        px_store(dst, dst.t.off(n, 0, y, x), c0, px_load(src, src.t.off(n, 0, y, x), c0));
    }
}
#undef FN2_PX_DECODE

// ---------------------------------------------------------------------------------------------
// Training-time augmentations of DataAugmentation (3-channel data).  Chromatic-eigen: ComputeChromaticEigenspace
// data_augmentation_layer.cu:148-185 + host finalisation :517-531, ChromaticEigenAugmentation :190-292; effects:
// ApplyEffects :295-318 and the Gaussian noise of :575-583.  space = 25 floats (tChromaticEigenSpace,
// augmentation_layer_base.hpp:117-129) followed by 3 doubles of scratch for the channel sums.
// The reference sums the mean with float atomics (order dependent); here the per-channel sums are double atomics, whose
// rounding is far below one float ulp.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_max_f(float* addr, float v) {       // v >= 0 or any sign: CAS loop like the reference
    float old = *addr;
    while (old < v) {
        const float assumed = old;
        old = __uint_as_float(atomicCAS(reinterpret_cast<unsigned int*>(addr), __float_as_uint(assumed), __float_as_uint(v)));
        if (old == assumed) break;
    }
}
__device__ __forceinline__ void atomic_min_f(float* addr, float v) {
    float old = *addr;
    while (old > v) {
        const float assumed = old;
        old = __uint_as_float(atomicCAS(reinterpret_cast<unsigned int*>(addr), __float_as_uint(assumed), __float_as_uint(v)));
        if (old == assumed) break;
    }
}
__global__ void eigenspace_init_kernel(float* space, const float* eigvec9) {
    const int i = threadIdx.x;
    if (i < 16) space[i] = (i >= 12 && i < 15) ? FLT_MAX : 0.f;
    if (i < 9) space[16 + i] = eigvec9[i];
    if (i < 3) reinterpret_cast<double*>(space + 26)[i] = 0.0;
}
__global__ void eigenspace_stats_kernel(T4 d, float* space) {
    const long long total = (long long)d.n * d.h * d.w;
    const float* ev = space + 16;
    float mx_eig[3] = {0, 0, 0}, mx[3] = {0, 0, 0}, mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    double sum[3] = {0, 0, 0};
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % d.w);
        const int y = (int)((idx / d.w) % d.h);
        const int n = (int)(idx / ((long long)d.w * d.h));
        float rgb[3];
        for (int c = 0; c < 3; c++) rgb[c] = d.p[d.off(n, c, y, x)];
        for (int c = 0; c < 3; c++) {
            const float eig = ev[3 * c] * rgb[0] + ev[3 * c + 1] * rgb[1] + ev[3 * c + 2] * rgb[2];
            mx_eig[c] = fmaxf(mx_eig[c], fabsf(eig));
            mx[c] = fmaxf(mx[c], rgb[c]);
            mn[c] = fminf(mn[c], rgb[c]);
            sum[c] += rgb[c];
        }
    }
    for (int c = 0; c < 3; c++) {
        for (int o = 16; o > 0; o >>= 1) {
            mx_eig[c] = fmaxf(mx_eig[c], __shfl_xor_sync(0xffffffffu, mx_eig[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
            sum[c] += __shfl_xor_sync(0xffffffffu, sum[c], o);
        }
        if ((threadIdx.x & 31) == 0) {
            atomic_max_f(space + 6 + c, mx_eig[c]);
            atomic_max_f(space + 9 + c, mx[c]);
            atomic_min_f(space + 12 + c, mn[c]);
            atomicAdd(reinterpret_cast<double*>(space + 26) + c, sum[c]);
        }
    }
}
__global__ void eigenspace_finish_kernel(float* space, int num, int height, int width) {
    if (threadIdx.x || blockIdx.x) return;
    float* mean_eig = space, *mean_rgb = space + 3, *max_abs_eig = space + 6;
    const float* ev = space + 16;
    const double* sum = reinterpret_cast<const double*>(space + 26);
    for (int c = 0; c < 3; c++) {
        mean_rgb[c] = (float)(sum[c] / (double)width / (double)height);
        mean_rgb[c] = mean_rgb[c] / num;
    }
    for (int c = 0; c < 3; c++) {
        mean_eig[c] = ev[3 * c] * mean_rgb[0] + ev[3 * c + 1] * mean_rgb[1] + ev[3 * c + 2] * mean_rgb[2];
        if (max_abs_eig[c] > 1e-2) mean_eig[c] = mean_eig[c] / max_abs_eig[c];
    }
    space[15] = sqrtf(max_abs_eig[0] * max_abs_eig[0] + max_abs_eig[1] * max_abs_eig[1] + max_abs_eig[2] * max_abs_eig[2]);
}
__global__ void chromatic_eigen_kernel(T4 d, const float* __restrict__ coeffs, const float* __restrict__ space, float max_multiplier) {
    const long long total = (long long)d.n * d.h * d.w;
    const float* mean_eig = space, *mean_rgb = space + 3, *max_abs_eig = space + 6, *eigvec = space + 16;
    const float max_l = space[15];
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % d.w);
        const int y = (int)((idx / d.w) % d.h);
        const int n = (int)(idx / ((long long)d.w * d.h));
        const float* ch = coeffs + 22 * n;
        float s, s1, l, l1 = 0.f;
        float rgb[3], eig[3];
        for (int c = 0; c < 3; c++) rgb[c] = d.p[d.off(n, c, y, x)] - mean_rgb[c];
        for (int c = 0; c < 3; c++) {
            eig[c] = eigvec[3 * c] * rgb[0] + eigvec[3 * c + 1] * rgb[1] + eigvec[3 * c + 2] * rgb[2];
            if (max_abs_eig[c] > 1e-2f) {
                eig[c] = eig[c] / max_abs_eig[c];
                eig[c] = copysignf(powf(fabsf(eig[c]), ch[c]), eig[c]);
                eig[c] = eig[c] + ch[3 + c];
                eig[c] = eig[c] * ch[6 + c];
            }
        }
        for (int c = 0; c < 3; c++) eig[c] = eig[c] + mean_eig[c];
        if (max_abs_eig[0] > 1e-2f) {
            eig[0] = copysignf(powf(fabsf(eig[0]), ch[9]), eig[0]);
            eig[0] = eig[0] + ch[12];
            eig[0] = eig[0] * ch[15];
        }
        s = sqrtf(eig[1] * eig[1] + eig[2] * eig[2]);
        s1 = s;
        if (s > 1e-2f) {
            s1 = powf(s1, ch[10]);
            s1 = fmaxf(s1 + ch[13], 0.f);
            s1 = s1 * ch[16];
        }
        if (ch[21] != 0) {
            const float t1 = cosf(ch[21]) * eig[1] - sinf(ch[21]) * eig[2];
            const float t2 = sinf(ch[21]) * eig[1] + cosf(ch[21]) * eig[2];
            eig[1] = t1; eig[2] = t2;
        }
        for (int c = 0; c < 3; c++) if (max_abs_eig[c] > 1e-2f) eig[c] = eig[c] * max_abs_eig[c];
        if (max_l > 1e-2f) {
            l1 = sqrtf(eig[0] * eig[0] + eig[1] * eig[1] + eig[2] * eig[2]);
            l1 = l1 / max_l;
        }
        if (s > 1e-2f) { eig[1] = eig[1] / s * s1; eig[2] = eig[2] / s * s1; }
        if (max_l > 1e-2f) {
            l = sqrtf(eig[0] * eig[0] + eig[1] * eig[1] + eig[2] * eig[2]);
            l1 = powf(l1, ch[18]);
            l1 = fmaxf(l1 + ch[19], 0.f);
            l1 = l1 * ch[20];
            l1 = l1 * max_l;
            if (l > 1e-2f)
                for (int c = 0; c < 3; c++) {
                    eig[c] = eig[c] / l * l1;
                    if (eig[c] > max_abs_eig[c]) eig[c] = max_abs_eig[c];
                }
        }
        for (int c = 0; c < 3; c++) {
            float v = eigvec[c] * eig[0] + eigvec[3 + c] * eig[1] + eigvec[6 + c] * eig[2];
            v = fminf(v, max_multiplier);
            v = fmaxf(v, 0.f);
            d.p[d.off(n, c, y, x)] = v;
        }
    }
}
__global__ void apply_effects_kernel(T4 d, const float* __restrict__ effects, float max_multiplier) {
    const long long total = (long long)d.n * d.h * d.w;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % d.w);
        const int y = (int)((idx / d.w) % d.h);
        const int n = (int)(idx / ((long long)d.w * d.h));
        const float* e = effects + 9 * n;
        const bool shadow = (x - d.w / 2) * e[4] + (y - d.h / 2) * e[5] - e[6] > 0;
        for (int c = 0; c < d.c; c++) {
            float sample = d.p[d.off(n, c, y, x)];
            if (shadow) sample -= e[7];
            d.p[d.off(n, c, y, x)] = clampf(sample, 0.f, max_multiplier);
        }
    }
}
// counter-based generator (SplitMix64 of (seed, element index)) + Box-Muller.  The reference draws from cuRAND's default
// generator with an unpinned seed/offset: the noise field is not reproducible there either, only its distribution.
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__global__ void gaussian_noise_kernel(T4 d, const float* __restrict__ effects, unsigned long long seed) {
    const long long total = d.count();
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % d.w);
        long long r = idx / d.w;
        const int y = (int)(r % d.h); r /= d.h;
        const int c = (int)(r % d.c);
        const int n = (int)(r / d.c);
        const float sigma = effects[9 * n + 8];
        if (!(sigma > 0)) continue;
        const unsigned long long h = splitmix64(seed ^ splitmix64((unsigned long long)idx));
        const float u1 = ((float)(h >> 40) + 1.0f) * (1.0f / 16777216.0f);                 // (0, 1]
        const float u2 = (float)((h >> 16) & 0xffffff) * (1.0f / 16777216.0f);
        const float g = sqrtf(-2.0f * logf(u1)) * cosf(6.28318530718f * u2);
        d.p[d.off(n, c, y, x)] = d.p[d.off(n, c, y, x)] + sigma * g;
    }
}

}  // namespace fn2

using namespace fn2;

extern "C" {

int fn2_flow_warp_forward(const fn2_tensor* image, const fn2_tensor* flow, const fn2_tensor* warped,
                          int fill_nan, void* stream) {
    FN2_CHECK_ARG(valid(image) && valid(flow) && valid(warped), "flow_warp: null/empty tensor");
    T4 img = view(image), fl = view(flow), out = view(warped);
    FN2_CHECK_ARG(fl.c == 2, "flow_warp: flow must have 2 channels (flow_warp_layer.cpp:46)");
    FN2_CHECK_ARG(fl.n == img.n && fl.h == img.h && fl.w == img.w,
                  "flow_warp: flow dims must match image (flow_warp_layer.cpp:45-48)");
    FN2_CHECK_ARG(same_dims(img, out), "flow_warp: top must have the image's shape");
    // 0xFFE00000 is the reference GPU NaN pattern (flow_warp_layer.cu:372-375)
    float fill = 0.f;
    if (fill_nan) { unsigned u = 0xFFE00000u; memcpy(&fill, &u, 4); }
    const long long total = (long long)out.n * out.h * out.w;
    const int vec4 = img.sc == 1 && img.c <= 4 && img.sw >= 4 && !((uintptr_t)img.p & 15) && !(img.sw & 3) && !(img.sh & 3) && !(img.sn & 3);
    flow_warp_fwd_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(img, fl, out, fill, vec4);
    FN2_LAUNCH_CHECK();
    return FN2_OK;
}

int fn2_flow_warp_backward(const fn2_tensor* image, const fn2_tensor* flow,
                           const fn2_tensor* warped_diff, const fn2_tensor* image_diff,
                           const fn2_tensor* flow_diff, void* stream) {
    FN2_CHECK_ARG(valid(image) && valid(flow) && valid(warped_diff) && valid(image_diff) && valid(flow_diff),
                  "flow_warp_backward: null/empty tensor");
    T4 img = view(image), fl = view(flow), wd = view(warped_diff), id = view(image_diff), fd = view(flow_diff);
    FN2_CHECK_ARG(fl.c == 2 && fd.c == 2 && same_dims(img, wd) && same_dims(img, id) && same_dims(fl, fd),
                  "flow_warp_backward: shape mismatch");
    fill_kernel<<<ew_grid(id.count(), 256), 256, 0, (cudaStream_t)stream>>>(id, 0.f);
    FN2_LAUNCH_CHECK();
    const long long total = (long long)img.n * img.h * img.w;
    flow_warp_bwd_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(img, fl, wd, id, fd);
    FN2_LAUNCH_CHECK();
    return FN2_OK;
}

int fn2_resample_forward(const fn2_tensor* bottom, const fn2_tensor* top, int type, int antialias,
                         void* stream) {
    FN2_CHECK_ARG(valid(bottom) && valid(top), "resample: null/empty tensor");
    T4 in = view(bottom), out = view(top);
    FN2_CHECK_ARG(in.n == out.n && in.c == out.c,
                  "resample: top channel count must match bottom (resample_layer.cu:143)");
    const float fx = (float)in.w / (float)out.w;
    const float fy = (float)in.h / (float)out.h;
    const int isDown = (fx > 1) || (fy > 1);
    const int aa = isDown && antialias;
    const int grid = ew_grid((long long)out.n * out.h * out.w, 256);
    cudaStream_t s = (cudaStream_t)stream;
    if (type == 1) resample_kernel<1><<<grid, 256, 0, s>>>(in, out, fx, fy, aa);
    else if (type == 2) resample_kernel<2><<<grid, 256, 0, s>>>(in, out, fx, fy, aa);
    else if (type == 3) resample_kernel<3><<<grid, 256, 0, s>>>(in, out, fx, fy, aa);
    else { set_error("resample: unsupported type %d (resample_layer.cu:204)", type); return FN2_ERR_INVALID; }
    FN2_LAUNCH_CHECK();
    return FN2_OK;
}

int fn2_warp_block_forward(const fn2_tensor* flow_in, const fn2_tensor* image0, const fn2_tensor* image1, const fn2_tensor* flow_full,
                           const fn2_tensor* warped, const fn2_tensor* err, const fn2_tensor* err_norm, const fn2_tensor* flow_scaled,
                           float scale_coeff, int fill_nan, void* stream) {
    FN2_CHECK_ARG(valid(flow_in) && valid(image0) && valid(image1) && valid(flow_full) && valid(warped) && valid(err) && valid(err_norm) &&
                  valid(flow_scaled), "warp_block: null/empty tensor");
    T4 fin = view(flow_in), i0 = view(image0), i1 = view(image1), ff = view(flow_full), wp = view(warped), er = view(err), en = view(err_norm),
       fs = view(flow_scaled);
    FN2_CHECK_ARG(fin.c == 2 && ff.c == 2 && fs.c == 2 && en.c == 1 && i1.c <= 4, "warp_block: channel counts");
    FN2_CHECK_ARG(same_dims(i0, i1) && same_dims(i0, wp) && same_dims(i0, er) && same_dims(ff, fs) && ff.n == i0.n && ff.h == i0.h && ff.w == i0.w &&
                  en.n == i0.n && en.h == i0.h && en.w == i0.w && fin.n == ff.n, "warp_block: shape mismatch");
    const float fx = (float)fin.w / (float)ff.w, fy = (float)fin.h / (float)ff.h;
    FN2_CHECK_ARG(fx <= 1.f && fy <= 1.f, "warp_block: the flow must be up-sampled (use the layer chain otherwise)");
    float fill = 0.f;
    if (fill_nan) { unsigned u = 0xFFE00000u; memcpy(&fill, &u, 4); }
    const int vec4 = i1.sc == 1 && i1.c <= 4 && i1.sw >= 4 && !((uintptr_t)i1.p & 15) && !(i1.sw & 3) && !(i1.sh & 3) && !(i1.sn & 3);
    const long long total = (long long)ff.n * ff.h * ff.w;
    warp_block_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(fin, i0, i1, ff, wp, er, en, fs, fx, fy, scale_coeff, fill, vec4);
    FN2_LAUNCH_CHECK();
    return FN2_OK;
}

int fn2_spatial_augmentation(const fn2_tensor* bottom, const fn2_tensor* top,
                             const float* trans_mats_dev, void* stream) {
    FN2_CHECK_ARG(valid(bottom) && valid(top) && trans_mats_dev, "spatial_augmentation: null argument");
    T4 in = view(bottom), out = view(top);
    FN2_CHECK_ARG(in.n == out.n && in.c == out.c, "spatial_augmentation: num/channels must match");
    FN2_CHECK_ARG(in.w >= 2 && in.h >= 2, "spatial_augmentation: bottom must be at least 2x2");
    const long long total = (long long)out.n * out.h * out.w;
    spatial_aug_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, trans_mats_dev);
    FN2_LAUNCH_CHECK();
    return FN2_OK;
}

int fn2_color_contrast_augmentation(const fn2_tensor* data, const float* chroma_dev,
                                    float max_multiplier, void* stream) {
    FN2_CHECK_ARG(valid(data) && chroma_dev, "color_contrast: null argument");
    T4 d = view(data);
    FN2_CHECK_ARG(d.c == 3, "Chromatic augmentations only work with 3-channel input "
                            "(data_augmentation_layer.cu:561)");
    const long long total = (long long)d.n * d.h * d.w;
    color_contrast_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(d, chroma_dev, max_multiplier);
    FN2_LAUNCH_CHECK();
    return FN2_OK;
}

int fn2_mean_update(const fn2_tensor* top, const fn2_tensor* mean_pp, float* mean_pc_dev,
                    float num_iter, void* stream) {
    FN2_CHECK_ARG(valid(top) && valid(mean_pp) && mean_pc_dev, "mean_update: null argument");
    T4 t = view(top), m = view(mean_pp);
    FN2_CHECK_ARG(m.n == 1 && m.c == t.c && m.h == t.h && m.w == t.w,
                  "mean_update: mean blob must be (1,C,H,W) (data_augmentation_layer.cu:603)");
    const long long total = (long long)t.c * t.h * t.w;
    mean_update_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(t, m, num_iter);
    FN2_LAUNCH_CHECK();
    mean_per_channel_kernel<<<t.c, 256, 0, (cudaStream_t)stream>>>(m, mean_pc_dev);
    FN2_LAUNCH_CHECK();
    return FN2_OK;
}

int fn2_mean_subtract(const fn2_tensor* top, const fn2_tensor* mean_pp, const float* mean_pc_dev,
                      int per_pixel, void* stream) {
    FN2_CHECK_ARG(valid(top), "mean_subtract: null top");
    T4 t = view(top), m = t;
    if (per_pixel) {
        FN2_CHECK_ARG(valid(mean_pp), "mean_subtract: per-pixel mean missing");
        m = view(mean_pp);
        FN2_CHECK_ARG(m.c == t.c && m.h == t.h && m.w == t.w, "mean_subtract: mean shape mismatch");
    } else {
        FN2_CHECK_ARG(mean_pc_dev, "mean_subtract: per-channel mean missing");
    }
    mean_subtract_kernel<<<ew_grid(t.count(), 256), 256, 0, (cudaStream_t)stream>>>(t, m, mean_pc_dev, per_pixel);
    FN2_LAUNCH_CHECK();
    return FN2_OK;
}

int fn2_relu_forward(const fn2_tensor* bottom, const fn2_tensor* top, float negative_slope, void* stream) {
    FN2_CHECK_ARG(valid(bottom) && valid(top), "relu: null/empty tensor");
    T4 in = view(bottom), out = view(top);
    FN2_CHECK_ARG(same_dims(in, out), "relu: shape mismatch");
    if (in.sc == 1 && out.sc == 1) {
        const long long groups = (long long)out.n * out.h * out.w * ((out.c + 3) / 4);
        relu_px_kernel<<<ew_grid(groups, 256), 256, 0, (cudaStream_t)stream>>>(px_view(in), px_view(out), negative_slope);
    } else
    relu_kernel<<<ew_grid(out.count(), 256), 256, 0, (cudaStream_t)stream>>>(in, out, negative_slope);
    FN2_LAUNCH_CHECK();
    return FN2_OK;
}

int fn2_eltwise_sum(const fn2_tensor* const* bottoms, const float* coeffs, int num_bottoms,
                    const fn2_tensor* top, void* stream) {
    FN2_CHECK_ARG(bottoms && valid(top) && num_bottoms >= 1 && num_bottoms <= 4,
                  "eltwise_sum: need 1..4 bottoms");
    EltArgs a;
    a.nb = num_bottoms;
    T4 out = view(top);
    for (int i = 0; i < num_bottoms; i++) {
        FN2_CHECK_ARG(valid(bottoms[i]), "eltwise_sum: null bottom");
        a.b[i] = view(bottoms[i]);
        FN2_CHECK_ARG(same_dims(a.b[i], out), "eltwise_sum: shape mismatch (eltwise_layer.cpp:33)");
        a.coeff[i] = coeffs ? coeffs[i] : 1.0f;
    }
    bool cfast = out.sc == 1;
    for (int i = 0; i < num_bottoms; i++) cfast = cfast && a.b[i].sc == 1;
    if (cfast) {
        EltPx e; e.nb = a.nb;
        for (int i = 0; i < a.nb; i++) { e.b[i] = px_view(a.b[i]); e.coeff[i] = a.coeff[i]; }
        const long long groups = (long long)out.n * out.h * out.w * ((out.c + 3) / 4);
        eltwise_sum_px_kernel<<<ew_grid(groups, 256), 256, 0, (cudaStream_t)stream>>>(e, px_view(out));
    } else
    eltwise_sum_kernel<<<ew_grid(out.count(), 256), 256, 0, (cudaStream_t)stream>>>(a, out);
    FN2_LAUNCH_CHECK();
    return FN2_OK;
}

int fn2_channel_norm_forward(const fn2_tensor* bottom, const fn2_tensor* top, void* stream) {
    FN2_CHECK_ARG(valid(bottom) && valid(top), "channel_norm: null/empty tensor");
    T4 in = view(bottom), out = view(top);
    FN2_CHECK_ARG(out.c == 1 && out.n == in.n && out.h == in.h && out.w == in.w,
                  "channel_norm: top must be (N,1,H,W) (channel_norm_layer.cpp:29)");
    const long long total = (long long)in.n * in.h * in.w;
    const int vec4 = in.sc == 1 && in.c <= 4 && in.sw >= 4 && !((uintptr_t)in.p & 15) && !(in.sw & 3) && !(in.sh & 3) && !(in.sn & 3);
    channel_norm_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, vec4);
    FN2_LAUNCH_CHECK();
    return FN2_OK;
}

int fn2_copy(const fn2_tensor* src, const fn2_tensor* dst, void* stream) {
    FN2_CHECK_ARG(valid(src) && valid(dst), "copy: null/empty tensor");
    T4 s = view(src), d = view(dst);
    FN2_CHECK_ARG(same_dims(s, d), "copy: shape mismatch");
    cudaStream_t st = (cudaStream_t)stream;
    const bool s_cf = (s.sc == 1), d_cf = (d.sc == 1);
    const bool s_pl = (s.sw == 1), d_pl = (d.sw == 1);
    if (d.c >= 8 && ((s_cf && d_pl && !d_cf) || (s_pl && d_cf && !s_cf))) {
        const long long hw = (long long)d.h * d.w;
        dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((d.c + 31) / 32), (unsigned)d.n);
        copy_transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(s, d, s_cf ? 1 : 0);
    } else if (s_cf && d_cf) {
        const long long groups = (long long)d.n * d.h * d.w * ((d.c + 3) / 4);
        copy_px_kernel<<<ew_grid(groups, 256), 256, 0, st>>>(px_view(s), px_view(d));
    } else if (d_cf || (s_cf && !d_pl)) {
        copy_kernel<true><<<ew_grid(d.count(), 256), 256, 0, st>>>(s, d);
    } else {
        copy_kernel<false><<<ew_grid(d.count(), 256), 256, 0, st>>>(s, d);
    }
    FN2_LAUNCH_CHECK();
    return FN2_OK;
}

int fn2_chromatic_eigenspace(const fn2_tensor* data, const float* eigvec9_dev, float* space_dev, void* stream) {
    FN2_CHECK_ARG(valid(data) && eigvec9_dev && space_dev, "chromatic_eigenspace: null argument");
    T4 d = view(data);
    FN2_CHECK_ARG(d.c == 3, "Chromatic-Eigen augmentations only work with 3-channel input (data_augmentation_layer.cu:488)");
    FN2_CHECK_ARG(!((uintptr_t)space_dev & 7), "chromatic_eigenspace: space must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    eigenspace_init_kernel<<<1, 32, 0, st>>>(space_dev, eigvec9_dev);
    FN2_LAUNCH_CHECK();
    const long long total = (long long)d.n * d.h * d.w;
    eigenspace_stats_kernel<<<ew_grid(total, 256), 256, 0, st>>>(d, space_dev);
    FN2_LAUNCH_CHECK();
    eigenspace_finish_kernel<<<1, 32, 0, st>>>(space_dev, d.n, d.h, d.w);
    FN2_LAUNCH_CHECK();
    return FN2_OK;
}

int fn2_chromatic_eigen_augmentation(const fn2_tensor* data, const float* coeffs_dev, const float* space_dev,
                                     float max_multiplier, void* stream) {
    FN2_CHECK_ARG(valid(data) && coeffs_dev && space_dev, "chromatic_eigen: null argument");
    T4 d = view(data);
    FN2_CHECK_ARG(d.c == 3, "Chromatic-Eigen augmentations only work with 3-channel input (data_augmentation_layer.cu:551)");
    const long long total = (long long)d.n * d.h * d.w;
    chromatic_eigen_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(d, coeffs_dev, space_dev, max_multiplier);
    FN2_LAUNCH_CHECK();
    return FN2_OK;
}

int fn2_apply_effects(const fn2_tensor* data, const float* effects_dev, float max_multiplier, unsigned long long noise_seed,
                      int add_noise, void* stream) {
    FN2_CHECK_ARG(valid(data) && effects_dev, "apply_effects: null argument");
    T4 d = view(data);
    FN2_CHECK_ARG(d.c == 3, "Effect augmentations only work with 3-channel input (data_augmentation_layer.cu:567)");
    const long long total = (long long)d.n * d.h * d.w;
    apply_effects_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(d, effects_dev, max_multiplier);
    FN2_LAUNCH_CHECK();
    if (add_noise) {
        gaussian_noise_kernel<<<ew_grid(d.count(), 256), 256, 0, (cudaStream_t)stream>>>(d, effects_dev, noise_seed);
        FN2_LAUNCH_CHECK();
    }
    return FN2_OK;
}

int fn2_fill(const fn2_tensor* dst, float value, void* stream) {
    FN2_CHECK_ARG(valid(dst), "fill: null/empty tensor");
    T4 d = view(dst);
    fill_kernel<<<ew_grid(d.count(), 256), 256, 0, (cudaStream_t)stream>>>(d, value);
    FN2_LAUNCH_CHECK();
    return FN2_OK;
}

}  // extern "C"
