// Training-side neighbours of the FlowNet2-C training step (SURVEY.md 8 row "next" 2): L1Loss, Downsample, FlowAugmentation.
//
// L1Loss (l1loss_layer.cpp:11-90, l1loss_layer.cu:67-192).  The reference chains an Eltwise difference, a NaN mask, optional
// Power(2) -> 1x1 sum convolution -> Power(0.5, shift = epsilon) ("l2_per_location", the end-point error), a plateau mask and two
// cuBLAS dot products through six intermediate blobs.  Here ONE pass over the two bottoms computes the masked loss sum and the
// not-NaN count per block, a second tiny kernel reduces the block partials in a fixed order (deterministic; cublasSdot's order is
// not) and writes loss and normaliser; the backward recomputes difference, masks and signs from the bottoms instead of reading
// stored intermediates: 2 reads forward, 2 reads + 2 writes backward, nothing else touches HBM.
// state (device, 4 floats): [0] loss, [1] normalize_coeff, [2] masked sum, [3] not-NaN count.
//
// Downsample (downsample_layer.cu:15-80): weighted box average around the rounded source position, NaN aware.
// FlowAugmentation (flow_augmentation_layer.cu:24-66): flow field of the augmented pair from the two spatial transforms.
#include "fn2_common.cuh"

namespace fn2 {

namespace {

struct L1P {
    int l2_per_location, two, plateau_on;
    float wsum, eps, plateau;
};

// per-location value / counts; shared by forward and backward so that both see the same masks
struct L1Loc {
    float s;            // l2: sum_c w * d^2 after the plateau mask
    bool plateau_kill;  // l2: location masked by the plateau
};

template <bool L2>
__global__ void l1loss_fwd_kernel(T4 a, T4 b, L1P p, float* __restrict__ part /* [grid][2] */) {
    const long long P = (long long)a.n * a.h * a.w;
    float loss = 0.f, cnt = 0.f;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < P; idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % a.w);
        const long long r = idx / a.w;
        const int y = (int)(r % a.h), n = (int)(r / a.h);
        float s = 0.f;
        for (int c = 0; c < a.c; c++) {
            float d = a.p[a.off(n, c, y, x)];
            if (p.two) d -= b.p[b.off(n, c, y, x)];
            const bool ok = d == d;                                     // FindNotNaNs
            cnt += ok ? 1.f : 0.f;
            if (L2) {
                const float dz = ok ? d : 0.f;                          // KillMasked
                s += p.wsum * (dz * dz);                                // Power(2) then the 1x1 sum convolution
            } else {
                const bool keep = ok && !(p.plateau_on && fabsf(d) < p.plateau);   // MaskPlateauValues
                loss += keep ? fabsf(d) : 0.f;                          // d * sign(d)
            }
        }
        if (L2) {
            if (p.plateau_on && fabsf(s) < p.plateau * p.plateau) s = 0.f;        // MaskPlateauValuesInitial + KillMasked
            loss += sqrtf(s + p.eps);                                   // Power(0.5, shift = epsilon)
        }
    }
    __shared__ float sl[32], sc[32];
    for (int o = 16; o; o >>= 1) { loss += __shfl_xor_sync(0xffffffffu, loss, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
    if ((threadIdx.x & 31) == 0) { sl[threadIdx.x >> 5] = loss; sc[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int nw = blockDim.x >> 5;
        loss = threadIdx.x < nw ? sl[threadIdx.x] : 0.f;
        cnt = threadIdx.x < nw ? sc[threadIdx.x] : 0.f;
        for (int o = 16; o; o >>= 1) { loss += __shfl_xor_sync(0xffffffffu, loss, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
        if (threadIdx.x == 0) { part[2 * blockIdx.x] = loss; part[2 * blockIdx.x + 1] = cnt; }
    }
}

__global__ void l1loss_final_kernel(const float* __restrict__ part, int blocks, int num, int channels, int by_entries,
                                    float* __restrict__ state, float* __restrict__ loss_out) {
    if (threadIdx.x || blockIdx.x) return;
    double loss = 0.0, cnt = 0.0;
    for (int i = 0; i < blocks; i++) { loss += (double)part[2 * i]; cnt += (double)part[2 * i + 1]; }
    const float norm = by_entries ? (float)cnt / (float)channels : (float)num;      // l1loss_layer.cu:83-88
    const float l = (float)loss / norm;
    state[0] = l; state[1] = norm; state[2] = (float)loss; state[3] = (float)cnt;
    if (loss_out) *loss_out = l;
}

template <bool L2>
__global__ void l1loss_bwd_kernel(T4 a, T4 b, L1P p, const float* __restrict__ state, const float* __restrict__ top_diff,
                                  T4 da, T4 db, int prop_a, int prop_b, int acc_a, int acc_b) {
    const long long P = (long long)a.n * a.h * a.w;
    const float alpha = top_diff[0] / state[1];                         // l1loss_layer.cu:152
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < P; idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % a.w);
        const long long r = idx / a.w;
        const int y = (int)(r % a.h), n = (int)(r / a.h);
        float gs = 0.f;
        if (L2) {
            float s = 0.f;
            for (int c = 0; c < a.c; c++) {
                float d = a.p[a.off(n, c, y, x)];
                if (p.two) d -= b.p[b.off(n, c, y, x)];
                const float dz = d == d ? d : 0.f;
                s += p.wsum * (dz * dz);
            }
            const bool kill = p.plateau_on && fabsf(s) < p.plateau * p.plateau;
            if (kill) s = 0.f;
            const float root = sqrtf(s + p.eps);
            // PowerLayer::Backward (power_layer.cu:49-84): shift != 0: top_data / (x + shift) * power; shift == 0: top_data / x * power
            gs = (root / (s + p.eps)) * 0.5f * alpha;
            if (kill) gs = 0.f;                                         // KillMasked(plateau_l2_) on the sum's diff
        }
        for (int c = 0; c < a.c; c++) {
            float d = a.p[a.off(n, c, y, x)];
            if (p.two) d -= b.p[b.off(n, c, y, x)];
            const bool ok = d == d;
            float g;
            if (L2) {
                const float dz = ok ? d : 0.f;
                g = (2.f * dz) * (p.wsum * gs);                         // Power(2) backward of the sum convolution's backward
                if (!ok) g = 0.f;                                       // KillMasked(mask_)
            } else {
                const bool keep = ok && !(p.plateau_on && fabsf(d) < p.plateau);
                const float dz = keep ? d : 0.f;
                g = keep ? alpha * (dz > 0.f ? 1.f : -1.f) : 0.f;       // ComputeSign: 0 -> -1
            }
            if (prop_a) { float* o = da.p + da.off(n, c, y, x); *o = acc_a ? *o + g : g; }
            if (prop_b && p.two) { float* o = db.p + db.off(n, c, y, x); *o = acc_b ? *o - g : -g; }
        }
    }
}

// ---- Downsample ------------------------------------------------------------------------------------------------------------
__global__ void downsample_kernel(T4 src, T4 dst, float wscale, float hscale, int wrad, int hrad) {
    const long long total = dst.count();
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int dx = (int)(idx % dst.w);
        long long r = idx / dst.w;
        const int dy = (int)(r % dst.h); r /= dst.h;
        const int c = (int)(r % dst.c), n = (int)(r / dst.c);
        const float bx = ((float)dx / (float)(dst.w - 1)) * (float)(src.w - 1);
        const float by = ((float)dy / (float)(dst.h - 1)) * (float)(src.h - 1);
        const int ix = (int)roundf(bx), iy = (int)roundf(by);
        float val = 0.f, wsum = 0.f, wnan = 0.f;
        for (int yo = -hrad; yo <= hrad; yo++) {
            const int sy = iy + yo;
            for (int xo = -wrad; xo <= wrad; xo++) {
                const int sx = ix + xo;
                if (sx < 0 || sy < 0 || sx >= src.w || sy >= src.h) continue;
                float sample = src.p[src.off(n, c, sy, sx)];
                float wgt = fmaxf(0.f, 1.f - fabsf((float)sx - bx) / wscale) * fmaxf(0.f, 1.f - fabsf((float)sy - by) / hscale);
                if (sample != sample) { wnan += wgt; sample = 0.f; wgt = 0.f; }
                val += sample * wgt;
                wsum += wgt;
            }
        }
        dst.p[dst.off(n, c, dy, dx)] = (wnan / wsum > 0.5f) ? __int_as_float(0x7fffffff) : val / wsum;
    }
}

// Wide windows (the 1/64 level averages 161 x 151 source pixels per output): one warp-group per output element, the window
// spread over the threads, fixed-order tree reduction.
__global__ void downsample_wide_kernel(T4 src, T4 dst, float wscale, float hscale, int wrad, int hrad) {
    const long long idx = blockIdx.x;
    const int dx = (int)(idx % dst.w);
    long long r = idx / dst.w;
    const int dy = (int)(r % dst.h); r /= dst.h;
    const int c = (int)(r % dst.c), n = (int)(r / dst.c);
    const float bx = ((float)dx / (float)(dst.w - 1)) * (float)(src.w - 1);
    const float by = ((float)dy / (float)(dst.h - 1)) * (float)(src.h - 1);
    const int ix = (int)roundf(bx), iy = (int)roundf(by);
    const int ww = 2 * wrad + 1, wh = 2 * hrad + 1;
    float val = 0.f, wsum = 0.f, wnan = 0.f;
    for (int t = threadIdx.x; t < ww * wh; t += blockDim.x) {
        const int sy = iy - hrad + t / ww, sx = ix - wrad + t % ww;
        if (sx < 0 || sy < 0 || sx >= src.w || sy >= src.h) continue;
        float sample = src.p[src.off(n, c, sy, sx)];
        float wgt = fmaxf(0.f, 1.f - fabsf((float)sx - bx) / wscale) * fmaxf(0.f, 1.f - fabsf((float)sy - by) / hscale);
        if (sample != sample) { wnan += wgt; sample = 0.f; wgt = 0.f; }
        val += sample * wgt;
        wsum += wgt;
    }
    __shared__ float s[3][8];
    for (int o = 16; o; o >>= 1) {
        val += __shfl_xor_sync(0xffffffffu, val, o); wsum += __shfl_xor_sync(0xffffffffu, wsum, o); wnan += __shfl_xor_sync(0xffffffffu, wnan, o);
    }
    if ((threadIdx.x & 31) == 0) { s[0][threadIdx.x >> 5] = val; s[1][threadIdx.x >> 5] = wsum; s[2][threadIdx.x >> 5] = wnan; }
    __syncthreads();
    if (threadIdx.x == 0) {
        val = wsum = wnan = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) { val += s[0][i]; wsum += s[1][i]; wnan += s[2][i]; }
        dst.p[dst.off(n, c, dy, dx)] = (wnan / wsum > 0.5f) ? __int_as_float(0x7fffffff) : val / wsum;
    }
}

// ---- FlowAugmentation ------------------------------------------------------------------------------------------------------
// m1 / m2: 6 floats per sample in NAME order t0..t5 (x' = x*t0 + y*t2 + t4, y' = x*t1 + y*t3 + t5); m2 is already inverted.
__global__ void flow_aug_kernel(T4 src, T4 dst, const float* __restrict__ m1, const float* __restrict__ m2) {
    const long long total = (long long)dst.n * dst.h * dst.w;
    const long long src_count = src.count();
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int xi = (int)(idx % dst.w);
        const long long r = idx / dst.w;
        const int yi = (int)(r % dst.h), n = (int)(r / dst.h);
        const float x = (float)xi, y = (float)yi;
        const float* a = m1 + 6 * n;
        const float* b = m2 + 6 * n;
        const float x1 = x * a[0] + y * a[2] + a[4];
        const float y1 = x * a[1] + y * a[3] + a[5];
        // nearest flow sample; the reference indexes the dense NCHW array with an UNCHECKED position (only min(idx, count)), so
        // a position outside the image reads a neighbouring row / sample or out of bounds.  Same linear index here, clamped into
        // the array (identical whenever the reference's read is in bounds).
        const int sy = (int)(y1 + 0.5f), sx = (int)(x1 + 0.5f);
        long long ou = (long long)src.w * ((long long)src.h * (2 * n + 0) + sy) + sx;
        long long ov = (long long)src.w * ((long long)src.h * (2 * n + 1) + sy) + sx;
        ou = max(0LL, min(ou, src_count - 1));
        ov = max(0LL, min(ov, src_count - 1));
        auto at = [&](long long lin) {
            const int w = (int)(lin % src.w); long long q = lin / src.w;
            const int h = (int)(q % src.h); q /= src.h;
            const int c = (int)(q % src.c), nn = (int)(q / src.c);
            return src.p[src.off(nn, c, h, w)];
        };
        const float x2 = x1 + at(ou), y2 = y1 + at(ov);
        const float x3 = x2 * b[0] + y2 * b[2] + b[4];
        const float y3 = x2 * b[1] + y2 * b[3] + b[5];
        dst.p[dst.off(n, 0, yi, xi)] = x3 - x;
        dst.p[dst.off(n, 1, yi, xi)] = y3 - y;
    }
}

int loss_grid(long long P) { return (int)max(1LL, min((long long)num_sms() * 4, (P + 255) / 256)); }

}  // namespace
}  // namespace fn2

using namespace fn2;

extern "C" {

int fn2_l1loss_workspace_bytes(int N, int H, int W, size_t* bytes) {
    FN2_CHECK_ARG(bytes && N > 0 && H > 0 && W > 0, "l1loss_workspace_bytes: bad argument");
    *bytes = (size_t)loss_grid((long long)N * H * W) * 2 * sizeof(float);
    return FN2_OK;
}

int fn2_l1loss_forward(const fn2_tensor* bottom0, const fn2_tensor* bottom1, const fn2_l1loss_desc* d, float* state_dev,
                       float* loss_dev, void* workspace, size_t workspace_bytes, void* stream) {
    FN2_CHECK_ARG(valid(bottom0) && d && state_dev && workspace, "l1loss_forward: null argument");
    T4 a = view(bottom0), b = a;
    if (bottom1) { FN2_CHECK_ARG(valid(bottom1), "l1loss_forward: bad second bottom"); b = view(bottom1); FN2_CHECK_ARG(same_dims(a, b), "l1loss_forward: bottoms differ in shape"); }
    const long long P = (long long)a.n * a.h * a.w;
    const int grid = loss_grid(P);
    FN2_CHECK_ARG(workspace_bytes >= (size_t)grid * 2 * sizeof(float), "l1loss_forward: workspace too small");
    L1P p;
    p.l2_per_location = d->l2_per_location; p.two = bottom1 ? 1 : 0; p.plateau_on = d->plateau > 0.f; p.plateau = d->plateau;
    p.eps = d->epsilon; p.wsum = d->l2_prescale_by_channels ? 1.f / (float)a.c : 1.f;
    cudaStream_t st = (cudaStream_t)stream;
    if (p.l2_per_location) l1loss_fwd_kernel<true><<<grid, 256, 0, st>>>(a, b, p, (float*)workspace);
    else l1loss_fwd_kernel<false><<<grid, 256, 0, st>>>(a, b, p, (float*)workspace);
    FN2_LAUNCH_CHECK();
    l1loss_final_kernel<<<1, 32, 0, st>>>((const float*)workspace, grid, a.n, a.c, d->normalize_by_num_entries, state_dev, loss_dev);
    FN2_LAUNCH_CHECK();
    return FN2_OK;
}

int fn2_l1loss_backward(const fn2_tensor* bottom0, const fn2_tensor* bottom1, const fn2_l1loss_desc* d, const float* state_dev,
                        const float* top_diff_dev, const fn2_tensor* bottom0_diff, const fn2_tensor* bottom1_diff,
                        int accumulate0, int accumulate1, void* stream) {
    FN2_CHECK_ARG(valid(bottom0) && d && state_dev && top_diff_dev, "l1loss_backward: null argument");
    FN2_CHECK_ARG(bottom0_diff || bottom1_diff, "l1loss_backward: no gradient requested");
    T4 a = view(bottom0), b = a, da = a, db = a;
    if (bottom1) { b = view(bottom1); FN2_CHECK_ARG(same_dims(a, b), "l1loss_backward: bottoms differ in shape"); }
    if (bottom0_diff) { da = view(bottom0_diff); FN2_CHECK_ARG(same_dims(a, da), "l1loss_backward: bottom0 diff shape"); }
    if (bottom1_diff) { FN2_CHECK_ARG(bottom1, "l1loss_backward: diff for a missing bottom"); db = view(bottom1_diff); FN2_CHECK_ARG(same_dims(a, db), "l1loss_backward: bottom1 diff shape"); }
    L1P p;
    p.l2_per_location = d->l2_per_location; p.two = bottom1 ? 1 : 0; p.plateau_on = d->plateau > 0.f; p.plateau = d->plateau;
    p.eps = d->epsilon; p.wsum = d->l2_prescale_by_channels ? 1.f / (float)a.c : 1.f;
    const int grid = loss_grid((long long)a.n * a.h * a.w);
    cudaStream_t st = (cudaStream_t)stream;
    if (p.l2_per_location)
        l1loss_bwd_kernel<true><<<grid, 256, 0, st>>>(a, b, p, state_dev, top_diff_dev, da, db, bottom0_diff ? 1 : 0, bottom1_diff ? 1 : 0, accumulate0, accumulate1);
    else
        l1loss_bwd_kernel<false><<<grid, 256, 0, st>>>(a, b, p, state_dev, top_diff_dev, da, db, bottom0_diff ? 1 : 0, bottom1_diff ? 1 : 0, accumulate0, accumulate1);
    FN2_LAUNCH_CHECK();
    return FN2_OK;
}

int fn2_downsample_forward(const fn2_tensor* bottom, const fn2_tensor* top, void* stream) {
    FN2_CHECK_ARG(valid(bottom) && valid(top), "downsample: null argument");
    T4 s = view(bottom), t = view(top);
    FN2_CHECK_ARG(s.n == t.n && s.c == t.c, "downsample: num / channels differ");
    if (s.h == t.h && s.w == t.w) return fn2_copy(bottom, top, stream);            // downsample_layer.cpp:55-58 (shared data)
    const float wscale = (float)(s.w - 1) / (float)(t.w - 1), hscale = (float)(s.h - 1) / (float)(t.h - 1);
    const int wrad = (int)ceilf(wscale), hrad = (int)ceilf(hscale);
    if ((2 * wrad + 1) * (2 * hrad + 1) >= 256 && t.count() <= (1 << 20))
        downsample_wide_kernel<<<(unsigned)t.count(), 256, 0, (cudaStream_t)stream>>>(s, t, wscale, hscale, wrad, hrad);
    else
        downsample_kernel<<<ew_grid(t.count(), 256), 256, 0, (cudaStream_t)stream>>>(s, t, wscale, hscale, wrad, hrad);
    FN2_LAUNCH_CHECK();
    return FN2_OK;
}

int fn2_flow_augmentation(const fn2_tensor* flow, const fn2_tensor* top, const float* mats1_dev, const float* mats2_inverse_dev, void* stream) {
    FN2_CHECK_ARG(valid(flow) && valid(top) && mats1_dev && mats2_inverse_dev, "flow_augmentation: null argument");
    T4 s = view(flow), t = view(top);
    FN2_CHECK_ARG(s.c == 2 && t.c == 2 && s.n == t.n, "flow_augmentation: flow blobs need 2 channels and equal num");
    flow_aug_kernel<<<ew_grid((long long)t.n * t.h * t.w, 256), 256, 0, (cudaStream_t)stream>>>(s, t, mats1_dev, mats2_inverse_dev);
    FN2_LAUNCH_CHECK();
    return FN2_OK;
}

}  // extern "C"
