// Microbenchmark: how fast can one persistent CTA per SM stream a corpus into shared memory?
//  mode 0: 2-D tensor-map TMA, box = 128 B x 128 rows out of a [rows][pitch] matrix (what the scan kernels do)
//  mode 1: 1-D cp.async.bulk of contiguous 16 KB blocks (a tile-major, pre-swizzled layout would allow this)
//  mode 2: 2-D tensor-map TMA, box = 128 B x 128 rows, matrix pitch = 128 B (contiguous via tensor map)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream tma_stream.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t ph) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
}

__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap tmap, const uint8_t *base, uint32_t ntiles,
                                                       int kchunks, int stages, int mode, unsigned long long *sink) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *full = (uint64_t *)(smem + (size_t)stages * 16384);
    uint64_t *empty = full + stages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // producer
        uint32_t s = 0, ph = 0;
        for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            for (int kc = 0; kc < kchunks; ++kc) {
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect(&full[s], 16384);
                uint8_t *dst = smem + (size_t)s * 16384;
                if (mode == 1) {
                    const uint8_t *src = base + ((size_t)tile * kchunks + kc) * 16384;
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(16384), "r"(smem_u32(&full[s])) : "memory");
                } else {
                    int c0 = mode == 0 ? kc * 128 : 0;
                    int c1 = mode == 0 ? (int)(tile * 128) : (int)((tile * kchunks + kc) * 128);
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)), "l"(&tmap), "r"(smem_u32(&full[s])), "r"(c0), "r"(c1) : "memory");
                }
                if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (threadIdx.x == 32) {  // consumer: touch one word, release the stage
        uint32_t s = 0, ph = 0;
        unsigned long long acc = 0;
        for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            for (int kc = 0; kc < kchunks; ++kc) {
                mbar_wait(&full[s], ph);
                acc += *(volatile uint32_t *)(smem + (size_t)s * 16384 + 64);
                mbar_arrive(&empty[s]);
                if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
            }
        }
        if (acc == 0x1234567ull) *sink = acc;
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
    const size_t rows = argc > 1 ? atoll(argv[1]) : 4000000;
    const int dim = 768, kchunks = dim / 128;
    uint8_t *d;
    CK(cudaMalloc(&d, rows * dim));
    CK(cudaMemset(d, 1, rows * dim));
    unsigned long long *sink;
    CK(cudaMalloc(&sink, 8));
    void *fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fnp;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const uint32_t ntiles = rows / 128;
    for (int mode = 0; mode < 3; ++mode) {
        for (int promo = 0; promo < 2; ++promo) {
            if (mode == 1 && promo) continue;
            CUtensorMap m;
            cuuint64_t dims[2] = {(cuuint64_t)(mode == 2 ? 128 : dim), (cuuint64_t)(mode == 2 ? rows * kchunks : rows)};
            cuuint64_t strides[1] = {(cuuint64_t)(mode == 2 ? 128 : dim)};
            cuuint32_t box[2] = {128, 128}, es[2] = {1, 1};
            CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
            for (int stages = 2; stages <= 12; stages += 2) {
                size_t smem = 1024 + (size_t)stages * 16384 + 2 * stages * 8 + 64;
                CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0); cudaEventCreate(&e1);
                float best = 1e9;
                for (int it = 0; it < 4; ++it) {
                    cudaEventRecord(e0);
                    stream_kernel<<<sms, 64, smem>>>(m, d, ntiles, kchunks, stages, mode, sink);
                    cudaEventRecord(e1);
                    CK(cudaEventSynchronize(e1));
                    float ms; cudaEventElapsedTime(&ms, e0, e1);
                    if (it > 0 && ms < best) best = ms;
                }
                printf("mode %d promo %d stages %2d (%4zu KB in flight/SM): %7.3f ms  %7.1f GB/s\n", mode, promo ? 256 : 128, stages,
                       (size_t)stages * 16, best, rows * dim / best / 1e6);
            }
        }
    }
    return 0;
}
