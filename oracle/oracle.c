/*
 * oracle.c -- CPU restatement of the FlowNet2 hot-path layer arithmetic of
 * lmb-freiburg/flownet2 (reference checkout at /root/reference, commit b92e198).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product path (flownet2_b200/) never links or calls anything here.
 *
 * PARITY PINNING STATUS
 *   - conv / deconv / ReLU / Eltwise / Concat: pinned by the reference's own known-answer
 *     tests (deconv-of-ones 3.1/6.1/12.1, Sobel separability, naive caffe_conv), see
 *     tests/test_oracle_golden.py.
 *   - FlowWarp: the reference has a CPU implementation (flow_warp_layer.cpp:57-117); this
 *     file follows it statement by statement.  No golden vectors exist in the reference.
 *   - Correlation / Resample / DataAugmentation / ChannelNorm: the reference has NO CPU
 *     implementation and NO tests ("parity unpinned").  The functions below restate the
 *     arithmetic of the reference .cu kernels; analytic identities in the tests are the
 *     only additional pin.
 *
 * All tensors are fp32, NCHW, row-major (Caffe Blob::offset, blob.hpp:153-163).
 * Floating-point contraction is OFF (compiled with -ffp-contract=off); wherever the
 * reference GPU kernel is a plain `sum += a*b` we use an explicit separate multiply and add.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define FN2O_API __attribute__((visibility("default")))

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* ------------------------------------------------------------------------------------ */
/* Correlation                                                                          */
/* ------------------------------------------------------------------------------------ */

/* Shape arithmetic of CorrelationLayer::Reshape, correlation_layer.cpp:41-84.
 * out[0]=top_channels out[1]=top_height out[2]=top_width out[3]=grid_radius out[4]=grid_width */
FN2O_API int fn2o_correlation_shape(int H, int W, int pad, int kernel_size, int max_disp,
                                    int stride1, int stride2, int* out) {
    if (kernel_size % 2 == 0) return -1;                       /* correlation_layer.cpp:22 */
    int pH = H + 2 * pad, pW = W + 2 * pad;
    int kernel_radius = (kernel_size - 1) / 2;                 /* :56 */
    int border = max_disp + kernel_radius;                     /* :57 */
    int top_w = (int)ceilf((float)(pW - border * 2) / (float)stride1);   /* :59 */
    int top_h = (int)ceilf((float)(pH - border * 2) / (float)stride1);   /* :60 */
    if (top_w < 1 || top_h < 1) return -2;                     /* :62-63 */
    int gr = max_disp / stride2;                               /* :66 */
    int gw = gr * 2 + 1;                                       /* :67 */
    out[0] = gw * gw; out[1] = top_h; out[2] = top_w; out[3] = gr; out[4] = gw;
    return 0;
}

/* blob_rearrange_kernel2, correlation_layer.cu:24-42 (+ memset :447-448):
 * NCHW -> zero padded NHWC.  Caller frees. */
static float* rearrange_padded_nhwc(const float* in, int N, int C, int H, int W, int pad) {
    int pH = H + 2 * pad, pW = W + 2 * pad;
    float* out = (float*)calloc((size_t)N * pH * pW * C, sizeof(float));
    for (int n = 0; n < N; n++)
        for (int c = 0; c < C; c++)
            for (int y = 0; y < H; y++)
                for (int x = 0; x < W; x++)
                    out[(((size_t)n * pH + (y + pad)) * pW + (x + pad)) * C + c] =
                        in[(((size_t)n * C + c) * H + y) * W + x];
    return out;
}

/* CorrelateData (MULTIPLY, corr_type 0) correlation_layer.cu:46-114 and
 * CorrelateDataSubtract (corr_type 1) :253-293.
 * mode_exact_order=1 reproduces the reference summation order for MULTIPLY: 32 lane-strided
 * partial sums (:84-97) added serially by lane 0 (:101-105); 0 sums channels in plain order
 * with a double accumulator (used to size tolerances).  */
FN2O_API int fn2o_correlation_fwd(const float* bot0, const float* bot1, float* top,
                                  int N, int C, int H, int W, int pad, int kernel_size,
                                  int max_disp, int stride1, int stride2, int corr_type,
                                  int mode_exact_order) {
    int shp[5];
    int rc = fn2o_correlation_shape(H, W, pad, kernel_size, max_disp, stride1, stride2, shp);
    if (rc) return rc;
    const int topC = shp[0], topH = shp[1], topW = shp[2], gr = shp[3], gw = shp[4];
    const int pH = H + 2 * pad, pW = W + 2 * pad;
    const int kr = (kernel_size - 1) / 2;
    float* r0 = rearrange_padded_nhwc(bot0, N, C, H, W, pad);
    float* r1 = rearrange_padded_nhwc(bot1, N, C, H, W, pad);
    const int sumelems = kernel_size * kernel_size * C;        /* :106 / :289 */
    const size_t topcount = (size_t)topC * topH * topW;

#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; n++)
        for (int y = 0; y < topH; y++)
            for (int x = 0; x < topW; x++) {
                /* upper-left corner of the patch in padded coords, :56-57 (MULTIPLY).  The
                 * SUBTRACT kernel uses the patch centre x1 = x*s1 + kr + md and loops
                 * j,i in [-kr,kr] (:265-272), which addresses the same pixels. */
                int x1 = x * stride1 + max_disp;
                int y1 = y * stride1 + max_disp;
                for (int tc = 0; tc < topC; tc++) {
                    int s2o = (tc % gw - gr) * stride2;        /* :81 */
                    int s2p = (tc / gw - gr) * stride2;        /* :82 */
                    int x2 = x1 + s2o, y2 = y1 + s2p;
                    float result;
                    if (corr_type == 0 && mode_exact_order) {
                        float lane[32];
                        for (int t = 0; t < 32; t++) lane[t] = 0.f;
                        for (int j = 0; j < kernel_size; j++)
                            for (int i = 0; i < kernel_size; i++) {
                                const float* a = r0 + (((size_t)n * pH + y1 + j) * pW + x1 + i) * C;
                                const float* b = r1 + (((size_t)n * pH + y2 + j) * pW + x2 + i) * C;
                                for (int ch = 0; ch < C; ch++) {
                                    float prod = a[ch] * b[ch];
                                    lane[ch & 31] = lane[ch & 31] + prod;
                                }
                            }
                        float total = 0.f;
                        for (int t = 0; t < 32; t++) total += lane[t];
                        result = total / (float)sumelems;
                    } else if (corr_type == 0) {
                        double acc = 0.0;
                        for (int j = 0; j < kernel_size; j++)
                            for (int i = 0; i < kernel_size; i++) {
                                const float* a = r0 + (((size_t)n * pH + y1 + j) * pW + x1 + i) * C;
                                const float* b = r1 + (((size_t)n * pH + y2 + j) * pW + x2 + i) * C;
                                for (int ch = 0; ch < C; ch++) acc += (double)a[ch] * (double)b[ch];
                            }
                        result = (float)(acc / (double)sumelems);
                    } else {
                        /* :268-284, plain serial float sum, order j,i,l */
                        float sum = 0.f;
                        double dsum = 0.0;
                        for (int j = 0; j < kernel_size; j++)
                            for (int i = 0; i < kernel_size; i++) {
                                const float* a = r0 + (((size_t)n * pH + y1 + j) * pW + x1 + i) * C;
                                const float* b = r1 + (((size_t)n * pH + y2 + j) * pW + x2 + i) * C;
                                for (int ch = 0; ch < C; ch++) {
                                    float d = fabsf(a[ch] - b[ch]);
                                    sum += d; dsum += (double)d;
                                }
                            }
                        result = mode_exact_order ? sum / (float)sumelems
                                                  : (float)(dsum / (double)sumelems);
                    }
                    (void)kr;
                    /* :107-108: index = ((tc*topH + y)*topW)+x ; top[index + item*topcount] */
                    top[(size_t)n * topcount + ((size_t)tc * topH + y) * topW + x] = result;
                }
            }
    free(r0); free(r1);
    return 0;
}

/* CorrelateDataBackward0/1 (MULTIPLY), correlation_layer.cu:118-249, drivers :508-572.
 * Integer ranges via the ROUND_OFF trick are restated literally (bit-exact index path). */
#define FN2O_ROUND_OFF 50000
FN2O_API int fn2o_correlation_bwd(const float* bot0, const float* bot1, const float* topdiff,
                                  float* bot0diff, float* bot1diff,
                                  int N, int C, int H, int W, int pad, int kernel_size,
                                  int max_disp, int stride1, int stride2) {
    int shp[5];
    int rc = fn2o_correlation_shape(H, W, pad, kernel_size, max_disp, stride1, stride2, shp);
    if (rc) return rc;
    const int topC = shp[0], topH = shp[1], topW = shp[2], gr = shp[3], gw = shp[4];
    const int pH = H + 2 * pad, pW = W + 2 * pad;
    const int kr = (kernel_size - 1) / 2;
    float* r0 = rearrange_padded_nhwc(bot0, N, C, H, W, pad);
    float* r1 = rearrange_padded_nhwc(bot1, N, C, H, W, pad);
    const int sumelems = (kr * 2 + 1) * (kr * 2 + 1) * C;
    const int round_off = FN2O_ROUND_OFF;
    const int round_off_s1 = stride1 * round_off;
    const size_t bottomcount = (size_t)C * H * W;

#pragma omp parallel for collapse(2) schedule(static)
    for (int item = 0; item < N; item++)
        for (int my = 0; my < H; my++)
            for (int lx = 0; lx < W; lx++)
                for (int n = 0; n < C; n++) {
                    int l = lx + pad, m = my + pad;
                    /* ---- Backward0 :131-173 ---- */
                    {
                        int xmin = (l - 2 * kr - max_disp + round_off_s1 - 1) / stride1 + 1 - round_off;
                        int ymin = (m - 2 * kr - max_disp + round_off_s1 - 1) / stride1 + 1 - round_off;
                        int xmax = (l - max_disp + round_off_s1) / stride1 - round_off;
                        int ymax = (m - max_disp + round_off_s1) / stride1 - round_off;
                        float sum = 0.f;
                        if (xmax >= 0 && ymax >= 0 && (xmin <= topW - 1) && (ymin <= topH - 1)) {
                            xmin = imax(0, xmin); xmax = imin(topW - 1, xmax);
                            ymin = imax(0, ymin); ymax = imin(topH - 1, ymax);
                            for (int p = -gr; p <= gr; p++)
                                for (int o = -gr; o <= gr; o++) {
                                    int s2o = stride2 * o, s2p = stride2 * p;
                                    float bot1tmp = r1[(((size_t)item * pH + (m + s2p)) * pW + (l + s2o)) * C + n];
                                    int op = (p + gr) * gw + (o + gr);
                                    size_t idxop = (size_t)item * topC + op;
                                    for (int y = ymin; y <= ymax; y++)
                                        for (int x = xmin; x <= xmax; x++) {
                                            float prod = topdiff[(idxop * topH + y) * topW + x] * bot1tmp;
                                            sum = sum + prod;
                                        }
                                }
                        }
                        bot0diff[(size_t)item * bottomcount + ((size_t)n * H + my) * W + lx] = sum / (float)sumelems;
                    }
                    /* ---- Backward1 :200-246 ---- */
                    {
                        float sum = 0.f;
                        for (int p = -gr; p <= gr; p++)
                            for (int o = -gr; o <= gr; o++) {
                                int s2o = stride2 * o, s2p = stride2 * p;
                                int xmin = (l - 2 * kr - max_disp - s2o + round_off_s1 - 1) / stride1 + 1 - round_off;
                                int ymin = (m - 2 * kr - max_disp - s2p + round_off_s1 - 1) / stride1 + 1 - round_off;
                                int xmax = (l - max_disp - s2o + round_off_s1) / stride1 - round_off;
                                int ymax = (m - max_disp - s2p + round_off_s1) / stride1 - round_off;
                                if (xmax >= 0 && ymax >= 0 && (xmin <= topW - 1) && (ymin <= topH - 1)) {
                                    xmin = imax(0, xmin); xmax = imin(topW - 1, xmax);
                                    ymin = imax(0, ymin); ymax = imin(topH - 1, ymax);
                                    float bot0tmp = r0[(((size_t)item * pH + (m - s2p)) * pW + (l - s2o)) * C + n];
                                    int op = (p + gr) * gw + (o + gr);
                                    size_t idxop = (size_t)item * topC + op;
                                    for (int y = ymin; y <= ymax; y++)
                                        for (int x = xmin; x <= xmax; x++) {
                                            float prod = topdiff[(idxop * topH + y) * topW + x] * bot0tmp;
                                            sum = sum + prod;
                                        }
                                }
                            }
                        bot1diff[(size_t)item * bottomcount + ((size_t)n * H + my) * W + lx] = sum / (float)sumelems;
                    }
                }
    free(r0); free(r1);
    return 0;
}

/* Correlation1D and the SUBTRACT gradients: correlation_layer1d.{cpp,cu} and correlation_layer.cu:298-427.
 * The reference indexes zero padded NHWC copies; here out-of-image taps are the same zeros by coordinate test (for
 * single_direction = -1 the reference's x_shift = -grid_width (correlation_layer1d.cu Forward_gpu) reaches stride_2 columns
 * past its own padding into the neighbouring row's padding -- also zeros).
 * one_d: rows are neither padded nor displaced (ymin/ymax without max_displacement, :134,:137).
 * corr_type 1: sign = (bot0 >= bot1) ? +1 : -1 taken AT THE DISPLACED position for both gradients (:327-329, :397-399). */
static float tap0(const float* b, int C, int H, int W, int n, int c, int y, int x) {
    return (y >= 0 && y < H && x >= 0 && x < W) ? b[(((size_t)n * C + c) * H + y) * W + x] : 0.f;
}
FN2O_API int fn2o_correlation1d_shape(int H, int W, int pad, int kernel_size, int max_disp, int stride1, int stride2,
                                      int single_direction, int* out) {
    if (kernel_size % 2 == 0) return -1;
    int kr = (kernel_size - 1) / 2, border = max_disp + kr;
    int top_w = (int)ceilf((float)(W + 2 * pad - border * 2) / (float)stride1);      /* correlation_layer1d.cpp:55 */
    int top_h = (int)ceilf((float)(H - kr * 2) / (float)stride1);                   /* :56 */
    if (top_w < 1 || top_h < 1) return -2;
    int gr = max_disp / stride2;
    int gw = single_direction != 0 ? gr + 1 : gr * 2 + 1;                            /* :64-68 */
    int x_shift = -gr;
    if (single_direction == -1) x_shift = -gw; else if (single_direction == 1) x_shift = 0;
    out[0] = gw; out[1] = top_h; out[2] = top_w; out[3] = gr; out[4] = x_shift;
    return 0;
}
FN2O_API int fn2o_correlation1d_fwd(const float* bot0, const float* bot1, float* top, int N, int C, int H, int W, int pad,
                                    int kernel_size, int max_disp, int stride1, int stride2, int single_direction, int corr_type) {
    int shp[5];
    int rc = fn2o_correlation1d_shape(H, W, pad, kernel_size, max_disp, stride1, stride2, single_direction, shp);
    if (rc) return rc;
    const int topC = shp[0], topH = shp[1], topW = shp[2], x_shift = shp[4];
    const int sumelems = kernel_size * kernel_size * C;
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; n++)
        for (int y = 0; y < topH; y++)
            for (int x = 0; x < topW; x++) {
                int x1 = x * stride1 + max_disp - pad, y1 = y * stride1;               /* unpadded; correlation_layer1d.cu:56-57 */
                for (int tc = 0; tc < topC; tc++) {
                    int s2o = (tc + x_shift) * stride2;                                 /* :83 */
                    double acc = 0;
                    for (int j = 0; j < kernel_size; j++)
                        for (int i = 0; i < kernel_size; i++)
                            for (int ch = 0; ch < C; ch++) {
                                float a = tap0(bot0, C, H, W, n, ch, y1 + j, x1 + i);
                                float b = tap0(bot1, C, H, W, n, ch, y1 + j, x1 + s2o + i);
                                acc += corr_type == 0 ? (double)a * (double)b : (double)fabsf(a - b);
                            }
                    top[(((size_t)n * topC + tc) * topH + y) * topW + x] = (float)(acc / (double)sumelems);
                }
            }
    return 0;
}
FN2O_API int fn2o_correlation_bwd_ex(const float* bot0, const float* bot1, const float* topdiff, float* bot0diff, float* bot1diff,
                                     int N, int C, int H, int W, int pad, int kernel_size, int max_disp, int stride1, int stride2,
                                     int corr_type, int one_d, int single_direction) {
    int shp[5];
    int topC, topH, topW, gr, gw = 0, x_shift = 0;
    if (one_d) {
        int rc = fn2o_correlation1d_shape(H, W, pad, kernel_size, max_disp, stride1, stride2, single_direction, shp);
        if (rc) return rc;
        topC = shp[0]; topH = shp[1]; topW = shp[2]; gr = shp[3]; x_shift = shp[4];
    } else {
        int rc = fn2o_correlation_shape(H, W, pad, kernel_size, max_disp, stride1, stride2, shp);
        if (rc) return rc;
        topC = shp[0]; topH = shp[1]; topW = shp[2]; gr = shp[3]; gw = shp[4];
    }
    const int kr = (kernel_size - 1) / 2;
    const int sumelems = (kr * 2 + 1) * (kr * 2 + 1) * C;
    const int ro = FN2O_ROUND_OFF, ros1 = stride1 * FN2O_ROUND_OFF;
    const int ypad = one_d ? 0 : pad, ymd = one_d ? 0 : max_disp;
#pragma omp parallel for collapse(2) schedule(static)
    for (int item = 0; item < N; item++)
        for (int my = 0; my < H; my++)
            for (int lx = 0; lx < W; lx++)
                for (int n = 0; n < C; n++) {
                    const int l = lx + pad, m = my + ypad;
                    double s0 = 0, s1 = 0;
                    for (int tc = 0; tc < topC; tc++) {
                        const int s2o = one_d ? (tc + x_shift) * stride2 : (tc % gw - gr) * stride2;
                        const int s2p = one_d ? 0 : (tc / gw - gr) * stride2;
                        /* gradient w.r.t. bottom0: ranges without the displacement (:131-140), other map at +displacement */
                        {
                            int xmin = (l - 2 * kr - max_disp + ros1 - 1) / stride1 + 1 - ro, xmax = (l - max_disp + ros1) / stride1 - ro;
                            int ymin = (m - 2 * kr - ymd + ros1 - 1) / stride1 + 1 - ro, ymax = (m - ymd + ros1) / stride1 - ro;
                            if (xmax >= 0 && ymax >= 0 && xmin <= topW - 1 && ymin <= topH - 1) {
                                xmin = imax(0, xmin); xmax = imin(topW - 1, xmax); ymin = imax(0, ymin); ymax = imin(topH - 1, ymax);
                                float a = tap0(bot0, C, H, W, item, n, my + s2p, lx + s2o), b = tap0(bot1, C, H, W, item, n, my + s2p, lx + s2o);
                                float f = corr_type == 0 ? b : (a >= b ? 1.f : -1.f);
                                for (int y = ymin; y <= ymax; y++)
                                    for (int x = xmin; x <= xmax; x++)
                                        s0 += (double)topdiff[(((size_t)item * topC + tc) * topH + y) * topW + x] * (double)f;
                            }
                        }
                        /* gradient w.r.t. bottom1: ranges shifted by the displacement (:208-215), other map at -displacement */
                        {
                            int xmin = (l - 2 * kr - max_disp - s2o + ros1 - 1) / stride1 + 1 - ro, xmax = (l - max_disp - s2o + ros1) / stride1 - ro;
                            int ymin = (m - 2 * kr - ymd - s2p + ros1 - 1) / stride1 + 1 - ro, ymax = (m - ymd - s2p + ros1) / stride1 - ro;
                            if (xmax >= 0 && ymax >= 0 && xmin <= topW - 1 && ymin <= topH - 1) {
                                xmin = imax(0, xmin); xmax = imin(topW - 1, xmax); ymin = imax(0, ymin); ymax = imin(topH - 1, ymax);
                                float a = tap0(bot0, C, H, W, item, n, my - s2p, lx - s2o), b = tap0(bot1, C, H, W, item, n, my - s2p, lx - s2o);
                                float f = corr_type == 0 ? a : (a >= b ? -1.f : 1.f);
                                for (int y = ymin; y <= ymax; y++)
                                    for (int x = xmin; x <= xmax; x++)
                                        s1 += (double)topdiff[(((size_t)item * topC + tc) * topH + y) * topW + x] * (double)f;
                            }
                        }
                    }
                    const size_t o = (((size_t)item * C + n) * H + my) * W + lx;
                    bot0diff[o] = (float)(s0 / (double)sumelems);
                    bot1diff[o] = (float)(s1 / (double)sumelems);
                }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* FlowWarp                                                                             */
/* ------------------------------------------------------------------------------------ */

/* FlowWarpLayer::Forward_cpu, flow_warp_layer.cpp:57-117.  fill_nan: 0 -> ZERO, 1 -> NaN. */
FN2O_API int fn2o_flow_warp_fwd(const float* image, const float* flow, float* warped,
                                int num, int channels, int height, int width, int fill_nan) {
    const int wh_size = width * height;
    const int whc_size = width * height * channels;
    const float fillValue = fill_nan ? NAN : 0.f;              /* :72 */
#pragma omp parallel for schedule(static)
    for (int n = 0; n < num; n++) {
        size_t off = (size_t)whc_size * n;
        for (int x = 0; x < width; x++)
            for (int y = 0; y < height; y++) {
                float fx = flow[(size_t)2 * wh_size * n + y * width + x];            /* :80 */
                float fy = flow[(size_t)2 * wh_size * n + wh_size + y * width + x];  /* :81 */
                float x2 = (float)x + fx;
                float y2 = (float)y + fy;
                if (x2 >= 0 && y2 >= 0 && x2 < width && y2 < height) {               /* :86 */
                    int ix2_L = (int)x2;
                    int iy2_T = (int)y2;
                    int ix2_R = imin(ix2_L + 1, width - 1);
                    int iy2_B = imin(iy2_T + 1, height - 1);
                    float alpha = x2 - ix2_L;
                    float beta = y2 - iy2_T;
                    for (int c = 0; c < channels; c++) {
                        float TL = image[off + (size_t)c * wh_size + iy2_T * width + ix2_L];
                        float TR = image[off + (size_t)c * wh_size + iy2_T * width + ix2_R];
                        float BL = image[off + (size_t)c * wh_size + iy2_B * width + ix2_L];
                        float BR = image[off + (size_t)c * wh_size + iy2_B * width + ix2_R];
                        /* :103-107, evaluated left to right without contraction */
                        float t0 = ((1 - alpha) * (1 - beta)) * TL;
                        float t1 = (alpha * (1 - beta)) * TR;
                        float t2 = ((1 - alpha) * beta) * BL;
                        float t3 = (alpha * beta) * BR;
                        warped[off + (size_t)c * wh_size + y * width + x] = ((t0 + t1) + t2) + t3;
                    }
                } else {
                    for (int c = 0; c < channels; c++)
                        warped[off + (size_t)c * wh_size + y * width + x] = fillValue;
                }
            }
    }
    return 0;
}

/* FlowWarpLayer::Backward_cpu, flow_warp_layer.cpp:120-198 (GPU twin .cu:170-229).
 * image_diff and flow_diff are overwritten. */
FN2O_API int fn2o_flow_warp_bwd(const float* image, const float* flow, const float* warped_diff,
                                float* image_diff, float* flow_diff,
                                int num, int channels, int height, int width) {
    const int wh_size = width * height;
    const int whc_size = width * height * channels;
    memset(image_diff, 0, sizeof(float) * (size_t)whc_size * num);
    memset(flow_diff, 0, sizeof(float) * (size_t)wh_size * 2 * num);
    for (int n = 0; n < num; n++) {
        size_t off = (size_t)whc_size * n;
        for (int x = 0; x < width; x++)
            for (int y = 0; y < height; y++) {
                float fx = flow[(size_t)2 * wh_size * n + y * width + x];
                float fy = flow[(size_t)2 * wh_size * n + wh_size + y * width + x];
                float x2 = (float)x + fx;
                float y2 = (float)y + fy;
                if (x2 >= 0 && y2 >= 0 && x2 < width && y2 < height) {
                    int ix2_L = (int)x2;
                    int iy2_T = (int)y2;
                    int ix2_R = imin(ix2_L + 1, width - 1);
                    int iy2_B = imin(iy2_T + 1, height - 1);
                    float alpha = x2 - ix2_L;
                    float beta = y2 - iy2_T;
                    for (int c = 0; c < channels; c++) {
                        float wd = warped_diff[off + (size_t)c * wh_size + y * width + x];
                        image_diff[off + (size_t)c * wh_size + iy2_T * width + ix2_L] += wd * (1 - alpha) * (1 - beta);
                        image_diff[off + (size_t)c * wh_size + iy2_T * width + ix2_R] += wd * alpha * (1 - beta);
                        image_diff[off + (size_t)c * wh_size + iy2_B * width + ix2_L] += wd * (1 - alpha) * beta;
                        image_diff[off + (size_t)c * wh_size + iy2_B * width + ix2_R] += wd * alpha * beta;
                    }
                    float gamma = iy2_B - y2;
                    float bot_diff = 0;
                    for (int c = 0; c < channels; c++) {
                        float temp = 0;
                        temp += gamma * (image[off + (size_t)c * wh_size + iy2_T * width + ix2_R] -
                                         image[off + (size_t)c * wh_size + iy2_T * width + ix2_L]);
                        temp += (1 - gamma) * (image[off + (size_t)c * wh_size + iy2_B * width + ix2_R] -
                                               image[off + (size_t)c * wh_size + iy2_B * width + ix2_L]);
                        bot_diff += warped_diff[off + (size_t)c * wh_size + y * width + x] * temp;
                    }
                    flow_diff[(size_t)2 * wh_size * n + y * width + x] = bot_diff;
                    gamma = ix2_R - x2;
                    bot_diff = 0;
                    for (int c = 0; c < channels; c++) {
                        float temp = 0;
                        temp += gamma * (image[off + (size_t)c * wh_size + iy2_B * width + ix2_L] -
                                         image[off + (size_t)c * wh_size + iy2_T * width + ix2_L]);
                        temp += (1 - gamma) * (image[off + (size_t)c * wh_size + iy2_B * width + ix2_R] -
                                               image[off + (size_t)c * wh_size + iy2_T * width + ix2_R]);
                        bot_diff += warped_diff[off + (size_t)c * wh_size + y * width + x] * temp;
                    }
                    flow_diff[(size_t)2 * wh_size * n + wh_size + y * width + x] = bot_diff;
                }
            }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* Resample                                                                             */
/* ------------------------------------------------------------------------------------ */

static inline float bicubicCoeff(float x_) {                   /* resample_layer.cu:14-20 */
    float x = fabsf(x_);
    if (x <= 1.0f) return x * x * (1.5f * x - 2.5f) + 1.0f;
    else if (x < 2.0f) return x * (x * (-0.5f * x + 2.5f) - 4.0f) + 2.0f;
    else return 0.0f;
}
static inline float triangleCoeff(float x) {                   /* resample_layer.cu:28-33 */
    if (-1 <= x && x < 0) return x + 1;
    if (0 <= x && x <= 1) return 1 - x;
    return 0;
}

/* ResampleLayer::Forward_gpu resample_layer.cu:128-206 with InterpolationKernel :40-95 and
 * NearestNeighborKernel :98-125.  type: 1 NEAREST, 2 LINEAR, 3 CUBIC (caffe.proto:666-671).
 * `planes` = N*C (c = index / out_channelsize spans N*C, :58).
 * NEAREST: the reference does not clamp (:120-123); out-of-range source indices are clamped
 * here (documented deviation -- the reference would read out of bounds). */
FN2O_API int fn2o_resample_fwd(const float* in, float* out, int planes, int in_h, int in_w,
                               int out_h, int out_w, int type, int antialias_param) {
    const float fx = (float)in_w / (float)out_w;                /* :146 */
    const float fy = (float)in_h / (float)out_h;                /* :147 */
    const int in_cs = in_w * in_h, out_cs = out_w * out_h;
    if (type == 1) {
#pragma omp parallel for schedule(static)
        for (int c = 0; c < planes; c++)
            for (int y_out = 0; y_out < out_h; y_out++)
                for (int x_out = 0; x_out < out_w; x_out++) {
                    float x_in = (x_out * fx + fy / 2.0f) - 0.5f;   /* :116 (sic: fy) */
                    float y_in = (y_out * fy + fx / 2.0f) - 0.5f;   /* :117 (sic: fx) */
                    int xr = (int)roundf(x_in), yr = (int)roundf(y_in);
                    xr = imin(imax(xr, 0), in_w - 1); yr = imin(imax(yr, 0), in_h - 1);
                    out[(size_t)c * out_cs + y_out * out_w + x_out] = in[(size_t)c * in_cs + yr * in_w + xr];
                }
        return 0;
    }
    if (type != 2 && type != 3) return -1;
    const int bicubic = (type == 3);
    const int kernel_width = bicubic ? 4 : 2;                   /* :182-185 */
    const int isDownsample = (fx > 1) || (fy > 1);              /* :179 */
    const int antialias = isDownsample && antialias_param;      /* :180 */
#pragma omp parallel for schedule(static)
    for (int c = 0; c < planes; c++)
        for (int y_out = 0; y_out < out_h; y_out++)
            for (int x_out = 0; x_out < out_w; x_out++) {
                float x_in = (x_out * fx + fy / 2.0f) - 0.5f;   /* :62 */
                float y_in = (y_out * fy + fx / 2.0f) - 0.5f;   /* :63 */
                int x_in_round = (int)roundf(x_in);             /* :65 */
                int y_in_round = (int)roundf(y_in);             /* :66 */
                float sum = 0, wsum = 0;
                float ax = 1.0f / (antialias ? fx : 1.0f);      /* :71 */
                float ay = 1.0f / (antialias ? fy : 1.0f);
                int rx = (fx < 1.0f) ? 2 : (int)ceilf((float)kernel_width / ax);  /* :73 */
                int ry = (fy < 1.0f) ? 2 : (int)ceilf((float)kernel_width / ay);
                for (int y = y_in_round - ry; y <= y_in_round + ry; y++)
                    for (int x = x_in_round - rx; x <= x_in_round + rx; x++) {
                        if (y < 0 || x < 0) continue;
                        if (y >= in_h || x >= in_w) continue;
                        float dx = x_in - x;
                        float dy = y_in - y;
                        float w;
                        if (bicubic) w = (ax * bicubicCoeff(ax * dx)) * ay * bicubicCoeff(ay * dy);
                        else         w = (ax * triangleCoeff(ax * dx)) * ay * triangleCoeff(ay * dy);
                        float prod = w * in[(size_t)c * in_cs + y * in_w + x];
                        sum = sum + prod;
                        wsum += w;
                    }
                out[(size_t)c * out_cs + y_out * out_w + x_out] = (!wsum) ? 0 : (sum / wsum);
            }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* DataAugmentation                                                                     */
/* ------------------------------------------------------------------------------------ */

/* tTransMat memory order is {t0,t2,t4,t1,t3,t5} in the reference struct; here the caller
 * passes 6 floats per sample in NAME order t0..t5.  Helpers restate
 * augmentation_layer_base.cpp:15-48. */
typedef struct { float t0, t1, t2, t3, t4, t5; } tmat;
static void tm_identity(tmat* t) { t->t0 = 1; t->t2 = 0; t->t4 = 0; t->t1 = 0; t->t3 = 1; t->t5 = 0; }
static void tm_left(tmat* t, float u0, float u1, float u2, float u3, float u4, float u5) {
    float t0 = t->t0, t2 = t->t2, t4 = t->t4, t1 = t->t1, t3 = t->t3, t5 = t->t5;
    t->t0 = t0 * u0 + t1 * u2;  t->t1 = t0 * u1 + t1 * u3;
    t->t2 = t2 * u0 + t3 * u2;  t->t3 = t2 * u1 + t3 * u3;
    t->t4 = (t4 * u0 + t5 * u2) + u4;  t->t5 = (t4 * u1 + t5 * u3) + u5;
}

/* tTransMat::fromCoeff, augmentation_layer_base.cpp:38-48.  has_* flags mirror protobuf
 * presence after clear_defaults (:339-349: fields within 1e-3 of their default are cleared).
 * out6 receives t0..t5. */
FN2O_API void fn2o_transmat_from_coeff(float mirror, float angle, float dx, float dy,
                                       float zoom_x, float zoom_y, int width, int height,
                                       int bottomwidth, int bottomheight, float* out6) {
    tmat t; tm_identity(&t);
    /* clear_defaults: defaults are mirror 0, angle 0, dx 0, dy 0, zoom 1 (caffe.proto:436-447) */
    int has_mirror = !(fabs(0.0 - mirror) < 1e-3);
    int has_angle = !(fabs(0.0 - angle) < 1e-3);
    int has_dx = !(fabs(0.0 - dx) < 1e-3), has_dy = !(fabs(0.0 - dy) < 1e-3);
    int has_zx = !(fabs(1.0 - zoom_x) < 1e-3), has_zy = !(fabs(1.0 - zoom_y) < 1e-3);
    if (!has_mirror) mirror = 0; if (!has_angle) angle = 0;
    if (!has_dx) dx = 0; if (!has_dy) dy = 0;
    if (!has_zx) zoom_x = 1; if (!has_zy) zoom_y = 1;
    if (mirror) tm_left(&t, -1, 0, 0, 1, (float)(.5 * (float)width), (float)(-.5 * (float)height));
    else        tm_left(&t, 1, 0, 0, 1, (float)(-.5 * (float)width), (float)(-.5 * (float)height));
    if (has_angle) tm_left(&t, (float)cos(angle), (float)sin(angle), (float)-sin(angle), (float)cos(angle), 0, 0);
    if (has_dx || has_dy) tm_left(&t, 1, 0, 0, 1, dx * (float)width, dy * (float)height);
    if (has_zx || has_zy) tm_left(&t, (float)(1.0 / zoom_x), 0, 0, (float)(1.0 / zoom_y), 0, 0);
    tm_left(&t, 1, 0, 0, 1, (float)(.5 * (float)bottomwidth), (float)(.5 * (float)bottomheight));
    out6[0] = t.t0; out6[1] = t.t1; out6[2] = t.t2; out6[3] = t.t3; out6[4] = t.t4; out6[5] = t.t5;
}

static inline float clampf(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }

/* SpatialAugmentation, data_augmentation_layer.cu:25-70.  mats: 6 floats per sample,
 * t0..t5 by name.  The index clamps min(idx, src_count) of :53-55 are kept. */
FN2O_API int fn2o_spatial_augmentation(const float* src, float* dst, const float* mats,
                                       int num, int channels, int height, int width,
                                       int dest_height, int dest_width) {
    const int src_count = num * channels * height * width;
#pragma omp parallel for schedule(static)
    for (int cn = 0; cn < num * channels; cn++) {
        int n = cn / channels;
        const float* m = mats + 6 * n;
        for (int y = 0; y < dest_height; y++)
            for (int x = 0; x < dest_width; x++) {
                float xpos = (x * m[0] + y * m[2]) + m[4];      /* :42 */
                float ypos = (x * m[1] + y * m[3]) + m[5];      /* :43 */
                xpos = clampf(xpos, 0.0f, (float)(width) - 1.05f);   /* :45 */
                ypos = clampf(ypos, 0.0f, (float)(height) - 1.05f);  /* :46 */
                float tlx = floorf(xpos);
                float tly = floorf(ypos);
                int srcIdxOff = (int)(width * (height * cn + tly) + tlx);   /* :52 float arithmetic */
                float sampleTL = src[srcIdxOff];
                float sampleTR = src[imin(srcIdxOff + 1, src_count)];
                float sampleBL = src[imin(srcIdxOff + width, src_count)];
                float sampleBR = src[imin(srcIdxOff + 1 + width, src_count)];
                float xdist = xpos - tlx;
                float ydist = ypos - tly;
                float s0 = ((1 - xdist) * (1 - ydist)) * sampleTL;   /* :62-65 term order TL,BR,BL,TR */
                float s1 = ((xdist) * (ydist)) * sampleBR;
                float s2 = ((1 - xdist) * (ydist)) * sampleBL;
                float s3 = ((xdist) * (1 - ydist)) * sampleTR;
                dst[((size_t)cn * dest_height + y) * dest_width + x] = ((s0 + s1) + s2) + s3;
            }
    }
    return 0;
}

/* ColorContrastAugmentation, data_augmentation_layer.cu:73-117.  chroma: per sample
 * {gamma, brightness, contrast, color0, color1, color2}.  In place on data (N,3,H,W). */
FN2O_API int fn2o_color_contrast_augmentation(float* data, const float* chroma, int num,
                                              int height, int width, float max_multiplier) {
#pragma omp parallel for schedule(static)
    for (int n = 0; n < num; n++) {
        const float* ch = chroma + 6 * n;
        for (int y = 0; y < height; y++)
            for (int x = 0; x < width; x++) {
                size_t idx[3]; float rgb[3]; float mean_in = 0, mean_out = 0;
                for (int c = 0; c < 3; c++) {
                    idx[c] = (size_t)width * ((size_t)height * (3 * n + c) + y) + x;
                    rgb[c] = data[idx[c]];
                    mean_in += rgb[c];
                    rgb[c] *= ch[3 + c];
                    mean_out += rgb[c];
                }
                float brightness_coeff = mean_in / (mean_out + 0.01f);
                for (int c = 0; c < 3; c++) {
                    rgb[c] = clampf(rgb[c] * brightness_coeff, 0.f, 1.f);
                    rgb[c] = powf(rgb[c], ch[0]);
                    rgb[c] = rgb[c] + ch[1];
                    rgb[c] = 0.5f + (rgb[c] - 0.5f) * ch[2];
                    data[idx[c]] = clampf(rgb[c], 0.f, max_multiplier);
                }
            }
    }
    return 0;
}

/* Chromatic-eigen augmentation (training use of DataAugmentation).  PARITY UNPINNED: the reference has no test or CPU path
 * for it; this restates data_augmentation_layer.cu:148-185 (ComputeChromaticEigenspace), :490-545 (host finalisation in
 * Forward_gpu) and :190-292 (ChromaticEigenAugmentation).  The reference kernels address the plane as [x*height + y]; since
 * every (x, y) is visited exactly once and the per-pixel arithmetic does not depend on the position, that is the plain
 * per-pixel transform restated here.  space[25] = mean_eig[3], mean_rgb[3], max_abs_eig[3], max_rgb[3], min_rgb[3],
 * max_l, eigvec[9] (tChromaticEigenSpace, augmentation_layer_base.hpp:117-129).  The mean is summed in double (the
 * reference's float atomicAdd order is not reproducible). */
FN2O_API int fn2o_chromatic_eigenspace(const float* data, int num, int height, int width, const float* eigvec9, float* space) {
    double sum[3] = {0, 0, 0};
    float max_abs_eig[3] = {0, 0, 0}, max_rgb[3] = {0, 0, 0}, min_rgb[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    const size_t plane = (size_t)height * width;
    for (int n = 0; n < num; n++)
        for (size_t i = 0; i < plane; i++) {
            float rgb[3];
            for (int c = 0; c < 3; c++) rgb[c] = data[((size_t)n * 3 + c) * plane + i];
            for (int c = 0; c < 3; c++) {
                float eig = eigvec9[3 * c] * rgb[0] + eigvec9[3 * c + 1] * rgb[1] + eigvec9[3 * c + 2] * rgb[2];
                if (fabsf(eig) > max_abs_eig[c]) max_abs_eig[c] = fabsf(eig);
                if (rgb[c] > max_rgb[c]) max_rgb[c] = rgb[c];
                if (rgb[c] < min_rgb[c]) min_rgb[c] = rgb[c];
                sum[c] += rgb[c];
            }
        }
    float* mean_eig = space, *mean_rgb = space + 3;
    for (int c = 0; c < 3; c++) {
        mean_rgb[c] = (float)(sum[c] / (double)width / (double)height);
        mean_rgb[c] = mean_rgb[c] / num;                                      /* :517-518 */
        space[6 + c] = max_abs_eig[c]; space[9 + c] = max_rgb[c]; space[12 + c] = min_rgb[c];
    }
    for (int c = 0; c < 3; c++) {                                            /* :520-526 */
        mean_eig[c] = eigvec9[3 * c] * mean_rgb[0] + eigvec9[3 * c + 1] * mean_rgb[1] + eigvec9[3 * c + 2] * mean_rgb[2];
        if (max_abs_eig[c] > 1e-2) mean_eig[c] = mean_eig[c] / max_abs_eig[c];
    }
    space[15] = sqrtf(max_abs_eig[0] * max_abs_eig[0] + max_abs_eig[1] * max_abs_eig[1] + max_abs_eig[2] * max_abs_eig[2]);
    for (int i = 0; i < 9; i++) space[16 + i] = eigvec9[i];
    return 0;
}

/* coeffs: per sample 22 floats in tChromaticEigenCoeffs order (augmentation_layer_base.hpp:52-75). In place, (N,3,H,W). */
FN2O_API int fn2o_chromatic_eigen_augmentation(float* data, const float* coeffs, const float* space, int num, int height,
                                               int width, float max_multiplier) {
    const float* mean_eig = space, *mean_rgb = space + 3, *max_abs_eig = space + 6, *eigvec = space + 16;
    const float max_l = space[15];
    const size_t plane = (size_t)height * width;
#pragma omp parallel for schedule(static)
    for (int n = 0; n < num; n++) {
        const float* ch = coeffs + 22 * n;
        const float pow_nomean[3] = {ch[0], ch[1], ch[2]}, add_nomean[3] = {ch[3], ch[4], ch[5]}, mult_nomean[3] = {ch[6], ch[7], ch[8]};
        const float pow_withmean0 = ch[9], pow_withmean1 = ch[10], add_withmean0 = ch[12], add_withmean1 = ch[13];
        const float mult_withmean0 = ch[15], mult_withmean1 = ch[16];
        const float lmult_pow = ch[18], lmult_add = ch[19], lmult_mult = ch[20], col_angle = ch[21];
        for (size_t i = 0; i < plane; i++) {
            float s, s1, l, l1 = 0.f;
            float rgb[3], eig[3];
            for (int c = 0; c < 3; c++) rgb[c] = data[((size_t)n * 3 + c) * plane + i] - mean_rgb[c];
            for (int c = 0; c < 3; c++) {
                eig[c] = eigvec[3 * c] * rgb[0] + eigvec[3 * c + 1] * rgb[1] + eigvec[3 * c + 2] * rgb[2];
                if (max_abs_eig[c] > 1e-2f) {
                    eig[c] = eig[c] / max_abs_eig[c];
                    eig[c] = copysignf(powf(fabsf(eig[c]), pow_nomean[c]), eig[c]);
                    eig[c] = eig[c] + add_nomean[c];
                    eig[c] = eig[c] * mult_nomean[c];
                }
            }
            for (int c = 0; c < 3; c++) eig[c] = eig[c] + mean_eig[c];
            if (max_abs_eig[0] > 1e-2f) {
                eig[0] = copysignf(powf(fabsf(eig[0]), pow_withmean0), eig[0]);
                eig[0] = eig[0] + add_withmean0;
                eig[0] = eig[0] * mult_withmean0;
            }
            s = sqrtf(eig[1] * eig[1] + eig[2] * eig[2]);
            s1 = s;
            if (s > 1e-2f) {
                s1 = powf(s1, pow_withmean1);
                s1 = fmaxf(s1 + add_withmean1, 0.f);
                s1 = s1 * mult_withmean1;
            }
            if (col_angle != 0) {
                float t1 = cosf(col_angle) * eig[1] - sinf(col_angle) * eig[2];
                float t2 = sinf(col_angle) * eig[1] + cosf(col_angle) * eig[2];
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
                l1 = powf(l1, lmult_pow);
                l1 = fmaxf(l1 + lmult_add, 0.f);
                l1 = l1 * lmult_mult;
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
                data[((size_t)n * 3 + c) * plane + i] = v;
            }
        }
    }
    return 0;
}

/* ApplyEffects, data_augmentation_layer.cu:295-318 (only the shadow and the clamp are implemented by the reference; fog and
 * motion blur coefficients are carried but unused; noise is added afterwards by cuRAND, :575-583 -- not restated, RNG
 * unpinned).  effects: per sample 9 floats in tEffectCoeffs order (fog_amount, fog_size, motion_blur_angle,
 * motion_blur_size, shadow_nx, shadow_ny, shadow_distance, shadow_strength, noise).  In place, (N,C,H,W). */
FN2O_API int fn2o_apply_effects(float* data, const float* effects, int num, int channels, int height, int width,
                                float max_multiplier) {
    for (int n = 0; n < num; n++) {
        const float* e = effects + 9 * n;
        for (int c = 0; c < channels; c++)
            for (int y = 0; y < height; y++)
                for (int x = 0; x < width; x++) {
                    float* p = data + (((size_t)n * channels + c) * height + y) * width + x;
                    float sample = *p;
                    if ((x - width / 2) * e[4] + (y - height / 2) * e[5] - e[6] > 0) sample -= e[7];
                    *p = clampf(sample, 0.f, max_multiplier);
                }
    }
    return 0;
}

/* Mean handling of DataAugmentationLayer::Forward_gpu, data_augmentation_layer.cu:594-634.
 * mode 0: recompute_mean>0 path.  state = {num_iter (already incremented, :353-354)},
 *         mean_pp[C*H*W], mean_pc[C] are layer blobs_[1], blobs_[2] and are UPDATED when
 *         num_iter <= recompute_mean (:600-608).  mean_per_pixel selects :610-613 vs :614-621.
 * mode 1: fixed per-channel `mean:` triple (:624-634), mean_pc holds the 3 values.
 * top is modified in place.  */
FN2O_API int fn2o_mean_subtract(float* top, int num, int channels, int height, int width,
                                int mode, float num_iter, int recompute_mean, int mean_per_pixel,
                                float* mean_pp, float* mean_pc) {
    const int area = height * width;
    const int count = area * channels;
    if (mode == 0) {
        if (num_iter <= (float)recompute_mean) {
            /* scal(count, num_iter-1), axpy(1/num) per sample in order, scal(1/num_iter) */
            for (int i = 0; i < count; i++) mean_pp[i] = mean_pp[i] * (num_iter - 1.0f);
            for (int n = 0; n < num; n++)
                for (int i = 0; i < count; i++) {
                    float p = (1.0f / (float)num) * top[(size_t)n * count + i];
                    mean_pp[i] = mean_pp[i] + p;
                }
            for (int i = 0; i < count; i++) mean_pp[i] = mean_pp[i] * (1.0f / num_iter);
            /* gemv: per channel (1/area) * sum over area; cuBLAS order unpinned -> double acc */
            for (int c = 0; c < channels; c++) {
                double acc = 0;
                for (int i = 0; i < area; i++) acc += mean_pp[(size_t)c * area + i];
                mean_pc[c] = (float)((1.0 / (double)area) * acc);
            }
        }
        if (mean_per_pixel) {
            for (int n = 0; n < num; n++)
                for (int i = 0; i < count; i++)
                    top[(size_t)n * count + i] = top[(size_t)n * count + i] - mean_pp[i];
        } else {
            for (int n = 0; n < num; n++)
                for (int c = 0; c < channels; c++)
                    for (int i = 0; i < area; i++) {
                        size_t k = (size_t)n * count + (size_t)c * area + i;
                        top[k] = top[k] - mean_pc[c];      /* gemm alpha=-1,K=1: top += -(m*1) */
                    }
        }
    } else {
        for (int n = 0; n < num; n++)
            for (int c = 0; c < channels; c++)
                for (int i = 0; i < area; i++) {
                    size_t k = (size_t)n * count + (size_t)c * area + i;
                    top[k] = top[k] - mean_pc[c];
                }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* Convolution / Deconvolution / glue                                                   */
/* ------------------------------------------------------------------------------------ */

/* im2col_cpu, util/im2col.cpp:19-55 (zero fill of the padding).  col is [C*kh*kw][Ho*Wo]. */
static void im2col(const float* im, int C, int H, int W, int kh, int kw, int ph, int pw,
                   int sh, int sw, int dh, int dw, float* col) {
    const int Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) / sh + 1;
    const int Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) / sw + 1;
    for (int c = 0; c < C; c++)
        for (int r = 0; r < kh; r++)
            for (int s = 0; s < kw; s++) {
                float* dst = col + (((size_t)c * kh + r) * kw + s) * Ho * Wo;
                for (int y = 0; y < Ho; y++) {
                    int iy = -ph + r * dh + y * sh;
                    for (int x = 0; x < Wo; x++) {
                        int ix = -pw + s * dw + x * sw;
                        dst[y * Wo + x] = (iy >= 0 && iy < H && ix >= 0 && ix < W)
                                              ? im[((size_t)c * H + iy) * W + ix] : 0.f;
                    }
                }
            }
}

/* ConvolutionLayer::Forward_cpu conv_layer.cpp:25-40 -> forward_cpu_gemm
 * base_conv_layer.cpp:257-273 + forward_cpu_bias :275-280.  Per sample im2col then
 * out[co][p] = sum_k W[co][k] col[k][p] accumulated in k-ascending order (the reference's
 * CBLAS order is unpinned, Makefile.config.example:46), then + bias.  f64acc=1 accumulates
 * in double (used to size tolerances).  weights [Co][Ci/g][kh][kw]. */
FN2O_API int fn2o_conv_fwd(const float* in, const float* weight, const float* bias, float* out,
                           int N, int Ci, int H, int W, int Co, int kh, int kw, int ph, int pw,
                           int sh, int sw, int dh, int dw, int group, int f64acc) {
    const int Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) / sh + 1;     /* conv_layer.cpp:8-22 */
    const int Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) / sw + 1;
    if (Ho < 1 || Wo < 1 || Ci % group || Co % group) return -1;
    const int Cig = Ci / group, Cog = Co / group;
    const size_t K = (size_t)Cig * kh * kw, P = (size_t)Ho * Wo;
    float* col = (float*)malloc(sizeof(float) * (size_t)Ci * kh * kw * P);
    if (!col) return -2;
    for (int n = 0; n < N; n++) {
        im2col(in + (size_t)n * Ci * H * W, Ci, H, W, kh, kw, ph, pw, sh, sw, dh, dw, col);
        if (!f64acc && group == 1) {
            /* Same arithmetic as the plain loop below (each output accumulates its products in
             * k-ascending order, then + bias), blocked 8 output channels x 1024 pixels so that a
             * col row is read once per 8 outputs; only there to make the CPU baseline less slow. */
            const int CB = 8;
            const size_t PB = 1024;
            const int ncb = (Co + CB - 1) / CB;
            const int npb = (int)((P + PB - 1) / PB);
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
            for (int cb = 0; cb < ncb; cb++)
                for (int pb = 0; pb < npb; pb++) {
                    const int co0 = cb * CB, nco = imin(CB, Co - co0);
                    const size_t p0 = (size_t)pb * PB, np_ = (P - p0 < PB) ? (P - p0) : PB;
                    float acc[8][1024];
                    for (int j = 0; j < nco; j++) for (size_t p = 0; p < np_; p++) acc[j][p] = 0.f;
                    for (size_t k = 0; k < K; k++) {
                        const float* cr = col + k * P + p0;
                        for (int j = 0; j < nco; j++) {
                            const float wv = weight[(size_t)(co0 + j) * K + k];
                            float* a = acc[j];
                            for (size_t p = 0; p < np_; p++) { float pr = wv * cr[p]; a[p] = a[p] + pr; }
                        }
                    }
                    for (int j = 0; j < nco; j++) {
                        float* o = out + ((size_t)n * Co + co0 + j) * P + p0;
                        if (bias) for (size_t p = 0; p < np_; p++) o[p] = acc[j][p] + bias[co0 + j];
                        else      for (size_t p = 0; p < np_; p++) o[p] = acc[j][p];
                    }
                }
            continue;
        }
#pragma omp parallel for schedule(dynamic, 1)
        for (int co = 0; co < Co; co++) {
            int g = co / Cog;
            const float* wrow = weight + (size_t)co * K;
            const float* colg = col + (size_t)g * K * P;
            float* o = out + ((size_t)n * Co + co) * P;
            if (f64acc) {
                double* acc = (double*)calloc(P, sizeof(double));
                for (size_t k = 0; k < K; k++) {
                    double wv = wrow[k]; const float* cr = colg + k * P;
                    for (size_t p = 0; p < P; p++) acc[p] += wv * (double)cr[p];
                }
                for (size_t p = 0; p < P; p++) o[p] = (float)(acc[p] + (bias ? (double)bias[co] : 0.0));
                free(acc);
            } else {
                for (size_t p = 0; p < P; p++) o[p] = 0.f;
                for (size_t k = 0; k < K; k++) {
                    float wv = wrow[k]; const float* cr = colg + k * P;
                    for (size_t p = 0; p < P; p++) { float pr = wv * cr[p]; o[p] = o[p] + pr; }
                }
                if (bias) for (size_t p = 0; p < P; p++) o[p] = o[p] + bias[co];
            }
        }
    }
    free(col);
    return 0;
}

/* DeconvolutionLayer::Forward_cpu deconv_layer.cpp:25-40 -> backward_cpu_gemm
 * base_conv_layer.cpp:283-298 (col = W^T * in, then col2im accumulate, im2col.cpp:158-190)
 * + bias.  weights [Ci][Co/g][kh][kw]; out size s*(in-1)+k_ext-2p (deconv_layer.cpp:18-19).
 * Gather form below visits contributions for an output pixel in (ci, r, s)-ascending order. */
FN2O_API int fn2o_deconv_fwd(const float* in, const float* weight, const float* bias, float* out,
                             int N, int Ci, int H, int W, int Co, int kh, int kw, int ph, int pw,
                             int sh, int sw, int group, int f64acc) {
    const int Ho = sh * (H - 1) + kh - 2 * ph;
    const int Wo = sw * (W - 1) + kw - 2 * pw;
    if (Ho < 1 || Wo < 1 || Ci % group || Co % group) return -1;
    const int Cig = Ci / group, Cog = Co / group;
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; n++)
        for (int co = 0; co < Co; co++) {
            int g = co / Cog, cog = co % Cog;
            for (int oy = 0; oy < Ho; oy++)
                for (int ox = 0; ox < Wo; ox++) {
                    float facc = 0.f; double dacc = 0.0;
                    for (int cig = 0; cig < Cig; cig++) {
                        int ci = g * Cig + cig;
                        for (int r = 0; r < kh; r++) {
                            int ty = oy + ph - r;
                            if (ty < 0 || ty % sh) continue;
                            int iy = ty / sh; if (iy >= H) continue;
                            for (int s = 0; s < kw; s++) {
                                int tx = ox + pw - s;
                                if (tx < 0 || tx % sw) continue;
                                int ix = tx / sw; if (ix >= W) continue;
                                float a = in[(((size_t)n * Ci + ci) * H + iy) * W + ix];
                                float wv = weight[(((size_t)ci * Cog + cog) * kh + r) * kw + s];
                                if (f64acc) dacc += (double)a * (double)wv;
                                else { float pr = a * wv; facc = facc + pr; }
                            }
                        }
                    }
                    float v = f64acc ? (float)(dacc + (bias ? (double)bias[co] : 0.0))
                                     : (bias ? facc + bias[co] : facc);
                    out[(((size_t)n * Co + co) * Ho + oy) * Wo + ox] = v;
                }
        }
    return 0;
}

/* ReLULayer::Forward_cpu relu_layer.cpp:9-19: max(x,0) + slope*min(x,0). */
FN2O_API void fn2o_relu(const float* in, float* out, size_t count, float negative_slope) {
    for (size_t i = 0; i < count; i++) {
        float x = in[i];
        out[i] = fmaxf(x, 0.f) + negative_slope * fminf(x, 0.f);
    }
}

/* EltwiseLayer SUM, eltwise_layer.cpp:59-65: top = 0; top += coeff_i * bottom_i in order. */
FN2O_API void fn2o_eltwise_sum(const float* const* bottoms, const float* coeffs, int nb,
                               float* top, size_t count) {
    for (size_t i = 0; i < count; i++) top[i] = 0.f;
    for (int b = 0; b < nb; b++)
        for (size_t i = 0; i < count; i++) { float p = coeffs[b] * bottoms[b][i]; top[i] = top[i] + p; }
}

/* ChannelNormLayer::Forward_cpu channel_norm_layer.cpp:43-69 / NormForward .cu:17-30:
 * top[n,0,y,x] = sqrt(sum_c x^2), channels in ascending order. */
FN2O_API void fn2o_channel_norm(const float* in, float* out, int N, int C, int H, int W) {
    const size_t area = (size_t)H * W;
    for (int n = 0; n < N; n++)
        for (size_t i = 0; i < area; i++) {
            float s = 0.f;
            for (int c = 0; c < C; c++) { float v = in[((size_t)n * C + c) * area + i]; float p = v * v; s = s + p; }
            out[(size_t)n * area + i] = sqrtf(s);
        }
}
