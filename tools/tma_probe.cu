// TMA probe: one CTA issues one 4-D tiled load with the given geometry and dumps what landed in shared memory.
// usage: tma_probe W H C N  bw bh bc  cx cy cc cn   (float32 tensor [N][C][H][W], value = linear index + 1)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned su32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap map, float* out, int nfl, int cx, int cy, int cc, int cn, int bytes) {
    extern __shared__ __align__(1024) unsigned char sm[];
    float* dst = (float*)sm;
    unsigned long long* bar = (unsigned long long*)(sm + ((nfl * 4 + 127) / 128) * 128);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(su32(dst)), "l"(&map), "r"(su32(bar)), "r"(cx), "r"(cy), "r"(cc), "r"(cn) : "memory");
    }
    unsigned done = 0; int spins = 0;
    while (!done && spins++ < (1 << 22))
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(su32(bar)) : "memory");
    if (threadIdx.x == 0 && !done) printf("TIMEOUT waiting for TMA\n");
    for (int i = threadIdx.x; i < nfl; i += blockDim.x) out[i] = dst[i];
}
int main(int argc, char** argv) {
    if (argc < 12) return 2;
    int W = atoi(argv[1]), H = atoi(argv[2]), C = atoi(argv[3]), N = atoi(argv[4]);
    int bw = atoi(argv[5]), bh = atoi(argv[6]), bc = atoi(argv[7]);
    int cx = atoi(argv[8]), cy = atoi(argv[9]), cc = atoi(argv[10]), cn = atoi(argv[11]);
    size_t n = (size_t)W * H * C * N;
    std::vector<float> h(n);
    for (size_t i = 0; i < n; i++) h[i] = (float)(i + 1);
    float *d, *o;
    cudaMalloc(&d, n * 4); cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
    int nfl = bw * bh * bc;
    cudaMalloc(&o, nfl * 4);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap m;
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)N};
    cuuint64_t st[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
    cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bc, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = ((EncodeTiledFn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d  ", (int)r);
    if (r) { printf("\n"); return 1; }
    int smem = ((nfl * 4 + 127) / 128) * 128 + 64;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<<<1, 128, smem>>>(m, o, nfl, cx, cy, cc, cn, nfl * 4);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s  ", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<float> g(nfl);
        cudaMemcpy(g.data(), o, nfl * 4, cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int c = 0; c < bc; c++) for (int y = 0; y < bh; y++) for (int x = 0; x < bw; x++) {
            int gx = cx + x, gy = cy + y, gc = cc + c;
            float want = (gx >= 0 && gx < W && gy >= 0 && gy < H && gc >= 0 && gc < C) ? (float)(((size_t)(cn * C + gc) * H + gy) * W + gx + 1) : 0.f;
            if (g[((size_t)c * bh + y) * bw + x] != want) bad++;
        }
        printf("mismatches=%ld of %d", bad, nfl);
    }
    printf("\n");
    return 0;
}
