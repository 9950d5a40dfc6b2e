// tcgen05 bring-up probe (sm_100a): validates the PTX forms, descriptors and TMEM layouts the conv engine uses,
// and measures the numbers its design depends on.
//   test 1: D[128x128] = A[128x32] * B[128x32]^T  (kind::tf32, A from TMEM, B from smem K-major SWIZZLE_128B)
//   test 2: cycles per tcgen05.mma 128x128x8 (TS), per tcgen05.ld 32x32b.x32, per tcgen05.st 32x32b.x32
//   test 3: round-toward-zero bias of a chain of n accumulating MMAs (random tf32 data) vs float64
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(su32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 columns, one row per thread
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
          "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
          "r"(v[30]), "r"(v[31]) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem_c] (+)= A[tmem_a] * B[desc_b]   (kind::tf32, cta_group::1)
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_c, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_c), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(su32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0; int spins = 0;
    while (!done && spins++ < (1 << 22))
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(su32(bar)), "r"(parity) : "memory");
    return done != 0;
}

// K-major SWIZZLE_128B tile of 32-bit elements: rows of 32 elements (128 B), 8-row atoms of 1024 B
__host__ __device__ inline int sw128_offset_floats(int row, int k) {
    return (row / 8) * 256 + (row % 8) * 32 + (((k / 4) ^ (row % 8)) * 4) + (k % 4);
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // start address
    d |= (uint64_t)1 << 16;                              // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset: 8-row atom pitch
    d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
    return d;
}
__host__ __device__ inline uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int N_TILE = 128;

// mode 0: correctness (nsteps MMAs over distinct K slices: K = 8*nsteps <= 32*kblocks staged in smem/TMEM)
// A: [128][K] row-major global, B: [128][K] row-major global, D: [128][128]
__global__ void __launch_bounds__(128, 1) probe_kernel(const float* A, const float* B, float* D, int K, int mode, long long* cycles, int iters) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* bs = reinterpret_cast<float*>(smem);                       // up to 4 k-blocks: 4 * 16 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * 16384);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 4 * 16384 + 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kblocks = K / 32;
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    // stage B (swizzled) with plain stores
    for (int i = tid; i < 128 * K; i += 128) {
        const int n = i / K, k = i % K;
        bs[(k / 32) * 4096 + sw128_offset_floats(n, k % 32)] = B[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> async proxy (UMMA reads)
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    // A rows into TMEM columns [256, 256 + K)
    const uint32_t colA = 256;
    for (int kb = 0; kb < kblocks; kb++) {
        uint32_t v[32];
        for (int j = 0; j < 32; j++) v[j] = __float_as_uint(A[(warp * 32 + lane) * K + kb * 32 + j]);
        tmem_st32(lane_base + colA + kb * 32, v);
    }
    tmem_wait_st();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t idesc = make_idesc_tf32(128, N_TILE);
    long long t0 = 0, t1 = 0;
    if (mode == 0 || mode == 3) {
        if (tid == 0) {
            for (int ks = 0; ks < K / 8; ks++) {
                const int kb = ks / 4, kk = ks % 4;
                const uint64_t db = make_desc_sw128(su32(bs + kb * 4096) + kk * 32);
                mma_tf32_ts(tmem, tmem + colA + ks * 8, db, idesc, ks > 0);
            }
            mma_commit(bar);
        }
        bool ok = mbar_wait(bar, 0);
        fence_after();
        if (!ok && tid == 0) printf("TIMEOUT waiting for MMA commit\n");
        for (int c = 0; c < N_TILE; c += 32) {
            uint32_t v[32];
            tmem_ld32(lane_base + c, v);
            tmem_wait_ld();
            for (int j = 0; j < 32; j++) D[(warp * 32 + lane) * N_TILE + c + j] = __uint_as_float(v[j]);
        }
    } else if (mode == 1) {
        // MMA issue rate: iters back-to-back MMAs on the same operands
        if (tid == 0) {
            const uint64_t db = make_desc_sw128(su32(bs));
            t0 = clock64();
            for (int i = 0; i < iters; i++) mma_tf32_ts(tmem + (i & 1) * 128, tmem + colA, db + (uint64_t)((i & 3) * 2), idesc, 1);
            mma_commit(bar);
        }
        bool ok = mbar_wait(bar, 0);
        if (tid == 0) { t1 = clock64(); cycles[0] = t1 - t0; if (!ok) printf("TIMEOUT\n"); }
    } else if (mode == 2) {
        // TMEM load / store rate: every warp moves 128 columns per iteration
        uint32_t v[32];
        uint32_t sink = 0;
        __syncthreads();
        t0 = clock64();
        for (int i = 0; i < iters; i++) {
            for (int c = 0; c < 128; c += 32) { tmem_ld32(lane_base + c, v); tmem_wait_ld(); sink += v[i & 31]; }
        }
        t1 = clock64();
        if (tid == 0) cycles[0] = t1 - t0;
        __syncthreads();
        t0 = clock64();
        for (int i = 0; i < iters; i++) {
            for (int c = 0; c < 128; c += 32) tmem_st32(lane_base + c, v);
            tmem_wait_st();
        }
        t1 = clock64();
        if (tid == 0) cycles[1] = t1 - t0;
        // loads without waiting in between (pipelined)
        uint32_t w[32];
        __syncthreads();
        t0 = clock64();
        for (int i = 0; i < iters; i++) {
            tmem_ld32(lane_base + 0, v); tmem_ld32(lane_base + 32, w);
            tmem_wait_ld(); sink += v[i & 31] + w[i & 31];
            tmem_ld32(lane_base + 64, v); tmem_ld32(lane_base + 96, w);
            tmem_wait_ld(); sink += v[i & 31] + w[i & 31];
        }
        t1 = clock64();
        if (tid == 0) cycles[2] = t1 - t0;
        if (sink == 0x12345678) D[0] = 1.f;
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= ~0x1FFFu; memcpy(&x, &u, 4); return x; }

int main(int argc, char** argv) {
    const int smem = 4 * 16384 + 256;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long* dcyc; cudaMalloc(&dcyc, 64);
    float *dA, *dB, *dD;
    const int KMAX = 128;
    cudaMalloc(&dA, 128 * KMAX * 4); cudaMalloc(&dB, 128 * KMAX * 4); cudaMalloc(&dD, 128 * 128 * 4);
    srand(1);
    // ---- test 1: correctness for K = 8 .. 128
    for (int K : {32, 64, 128}) {
        std::vector<float> A(128 * K), B(128 * K), D(128 * 128);
        for (auto& v : A) v = tf32_trunc((float)rand() / RAND_MAX - 0.5f);
        for (auto& v : B) v = tf32_trunc((float)rand() / RAND_MAX - 0.5f);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(dD, 0, 128 * 128 * 4);
        probe_kernel<<<1, 128, smem>>>(dA, dB, dD, K, 0, dcyc, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("test1 K=%d: CUDA error %s\n", K, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (int m = 0; m < 128; m++) for (int n = 0; n < 128; n++) {
            double r = 0;
            for (int k = 0; k < K; k++) r += (double)A[m * K + k] * (double)B[n * K + k];
            maxerr = fmax(maxerr, fabs(r - D[m * 128 + n])); maxref = fmax(maxref, fabs(r));
        }
        printf("test1 K=%3d: max |D - ref| = %.3e (max |ref| %.3f)  %s\n", K, maxerr, maxref, maxerr < 1e-5 ? "OK" : "MISMATCH");
    }
    // ---- test 2: rates
    {
        long long c[8];
        probe_kernel<<<1, 128, smem>>>(dA, dB, dD, 32, 1, dcyc, 2000);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(c, dcyc, 64, cudaMemcpyDeviceToHost);
        printf("test2 MMA 128x128x8 tf32 (A in TMEM): %.1f cycles/MMA over 2000 (%s)\n", c[0] / 2000.0, cudaGetErrorString(e));
        probe_kernel<<<1, 128, smem>>>(dA, dB, dD, 32, 2, dcyc, 500);
        e = cudaDeviceSynchronize();
        cudaMemcpy(c, dcyc, 64, cudaMemcpyDeviceToHost);
        printf("test2 TMEM: ld 64 KB/iter: %.1f cyc/iter (%.1f B/clk/SM); st: %.1f cyc/iter (%.1f B/clk/SM); ld pipelined x2: %.1f cyc/iter (%.1f B/clk/SM) (%s)\n",
               c[0] / 500.0, 65536.0 / (c[0] / 500.0), c[1] / 500.0, 65536.0 / (c[1] / 500.0), c[2] / 500.0, 65536.0 / (c[2] / 500.0),
               cudaGetErrorString(e));
    }
    // ---- test 3: RZ bias of n-step accumulation chains, random-sign data
    for (int K : {8, 32, 64, 128}) {
        double shrink = 0, abserr = 0, absref = 0;
        int reps = 20;
        for (int rep = 0; rep < reps; rep++) {
            std::vector<float> A(128 * K), B(128 * K), D(128 * 128);
            for (auto& v : A) v = tf32_trunc((float)rand() / RAND_MAX - 0.5f);
            for (auto& v : B) v = tf32_trunc((float)rand() / RAND_MAX - 0.5f);
            cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
            cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
            probe_kernel<<<1, 128, smem>>>(dA, dB, dD, K, 3, dcyc, 0);
            cudaDeviceSynchronize();
            cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
            for (int m = 0; m < 128; m++) for (int n = 0; n < 128; n++) {
                double r = 0;
                for (int k = 0; k < K; k++) r += (double)A[m * K + k] * (double)B[n * K + k];
                shrink += (D[m * 128 + n] - r) * (r > 0 ? 1 : -1);
                abserr += fabs(D[m * 128 + n] - r); absref += fabs(r);
            }
        }
        printf("test3 chain of %2d MMAs (K=%3d): mean signed err / mean|ref| = %+.3e, mean |err| / mean|ref| = %.3e\n",
               K / 8, K, shrink / absref, abserr / absref);
    }
    return 0;
}
