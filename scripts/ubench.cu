// Micro-benchmarks that size the GEMM pipeline (run on the B200 through gpurun):
//   1. latency of tcgen05.commit -> mbarrier phase completion, with no MMA pending and after one MMA
//   2. latency and per-SM throughput of a 16 KB TMA tile load (128 rows x 128 B, row stride 512 B)
//   3. round trip of an empty producer -> 128 consumers -> committer mbarrier ring
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench scripts/ubench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t par)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) { while (!mbar_try(b, par)) {} }
__device__ __forceinline__ void commit(uint64_t* b) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ uint64_t make_desc(uint32_t a) { return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61); }
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(kIdesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tma3(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// ---- 1. commit latency ----
__global__ void k_commit(long long* out, int n_mma)
{
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(128u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        long long best = 1ll << 60, sum = 0;
        for (int i = 0; i < 64; i++) {
            const long long t0 = clock64();
            for (int k = 0; k < n_mma; k++) mma(tm, make_desc(smem_u32(smem)), make_desc(smem_u32(smem + 16384)), k);
            commit(&bar);
            mbar_wait(&bar, i & 1);
            const long long dt = clock64() - t0;
            if (dt < best) best = dt;
            if (i >= 8) sum += dt;
        }
        out[0] = best; out[1] = sum / 56;
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(128u));
}

// ---- 2. TMA: latency of one tile, throughput with `depth` tiles in flight ----
__global__ void k_tma(const __grid_constant__ CUtensorMap map, long long* out, int depth, int iters, int n_rows_total)
{
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[16];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; i++) mbar_init(&bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
        // latency
        long long best = 1ll << 60;
        for (int i = 0; i < 16; i++) {
            const long long t0 = clock64();
            mbar_expect(&bar[0], 16384);
            tma3(smem, &map, &bar[0], 0, (int)((blockIdx.x * 131 + i * 977) % (n_rows_total - 128)), 0);
            mbar_wait(&bar[0], i & 1);
            const long long dt = clock64() - t0;
            if (dt < best) best = dt;
        }
        // throughput: keep `depth` loads in flight
        const long long t0 = clock64();
        for (int i = 0; i < iters + depth; i++) {
            const int s = i % depth;
            if (i >= depth) mbar_wait(&bar[1 + s], ((i / depth) - 1) & 1);
            if (i < iters) {
                mbar_expect(&bar[1 + s], 16384);
                tma3(smem + s * 16384, &map, &bar[1 + s], 0, (int)((blockIdx.x * 977 + i * 128) % (n_rows_total - 128)), 0);
            }
        }
        const long long dt = clock64() - t0;
        if (blockIdx.x == 0) { out[0] = best; out[1] = dt; }
    }
}

// ---- 3. empty handshake ring: producer -> 128 consumers -> committer (tcgen05.commit) -> producer ----
__global__ void k_ring(long long* out, int stages, int iters, int use_commit)
{
    __shared__ uint64_t full[8], split[8], empty[8];
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 8; s++) { mbar_init(&full[s], 1); mbar_init(&split[s], 128); mbar_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    const long long t0 = clock64();
    if (warp == 0) {
        if (threadIdx.x == 0)
            for (int it = 0; it < iters; it++) { const int s = it % stages; mbar_wait(&empty[s], ((it / stages) & 1) ^ 1); mbar_arrive(&full[s]); }
    } else if (warp == 1) {
        if (threadIdx.x == 32)
            for (int it = 0; it < iters; it++) {
                const int s = it % stages;
                mbar_wait(&split[s], (it / stages) & 1);
                if (use_commit) commit(&empty[s]); else mbar_arrive(&empty[s]);
            }
    } else {
        for (int it = 0; it < iters; it++) { const int s = it % stages; mbar_wait(&full[s], (it / stages) & 1); mbar_arrive(&split[s]); }
    }
    __syncthreads();
    if (threadIdx.x == 0) out[0] = clock64() - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    long long* out; cudaMallocManaged(&out, 64);
    cudaFuncSetAttribute(k_commit, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    for (int n = 0; n <= 12; n += (n < 4 ? 1 : 4)) {
        k_commit<<<1, 128, 40000>>>(out, n); cudaDeviceSynchronize();
        printf("commit after %2d MMA(128x128x8 tf32): min %lld avg %lld cycles  (%s)\n", n, out[0], out[1], cudaGetErrorString(cudaGetLastError()));
    }
    for (int c = 0; c < 2; c++)
        for (int st = 2; st <= 8; st *= 2) {
            k_ring<<<1, 192, 0>>>(out, st, 2000, c); cudaDeviceSynchronize();
            printf("empty ring, %d stages, %s: %.1f cycles / iteration\n", st, c ? "tcgen05.commit" : "mbarrier.arrive", out[0] / 2000.0);
        }
    // TMA: tensor [32 ch of 128][rows] with row stride 512 B like a 128-channel activation
    const int rows = 1 << 18;
    float* buf; cudaMalloc(&buf, (size_t)rows * 512); cudaMemset(buf, 0, (size_t)rows * 512);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap map;
    cuuint64_t dims[3] = {128, (cuuint64_t)rows, 1}, strides[2] = {512, (cuuint64_t)rows * 512};
    cuuint32_t box[3] = {32, 128, 1}, es[3] = {1, 1, 1};
    CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r);
    cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
    for (int grid = 1; grid <= 148; grid *= 148)
        for (int depth = 1; depth <= 8; depth *= 2) {
            k_tma<<<grid, 32, 200000>>>(map, out, depth, 512, rows); cudaDeviceSynchronize();
            printf("TMA 16KB tile, grid %3d, depth %d: latency %lld cyc; %.1f cycles / tile -> %.1f B/clk/SM  (%s)\n", grid, depth, out[0], out[1] / 512.0, 16384.0 * 512 / out[1], cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
