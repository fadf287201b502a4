// Micro-benchmark for the arithmetic-coder step (run on the B200 through gpurun): cycles per iteration of
// the dependent chain  range -> floor(p * range) -> new range  with one warp per SM,
//   1. the reference's arithmetic: I2F.F64.U32, DMUL, F2I.U32.F64.FLOOR
//   2. a 48-bit fixed-point multiply (two IMADs and a shift)
//   3. a plain dependent integer add chain (the per-instruction issue latency)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_coder scripts/ubench_coder.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__global__ void k_fp64(long long* out, uint32_t* sink, double p, int iters)
{
    uint32_t r = 40000u + threadIdx.x;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        const uint32_t f = __double2uint_rd(__dmul_rn(p, (double)r));
        r = (f | 0x8000u) + (r & 1u);       // keeps the chain dependent and the value in range
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    sink[threadIdx.x] = r;
}

__global__ void k_fixed(long long* out, uint32_t* sink, uint64_t q, int iters)
{
    uint32_t r = 40000u + threadIdx.x;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        const uint32_t f = (uint32_t)((q * (uint64_t)r) >> 48);
        r = (f | 0x8000u) + (r & 1u);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[1] = t1 - t0;
    sink[threadIdx.x] = r;
}

__global__ void k_add(long long* out, uint32_t* sink, uint32_t c, int iters)
{
    uint32_t r = threadIdx.x;
    const long long t0 = clock64();
    #pragma unroll 16
    for (int i = 0; i < iters; i++) r = (r ^ c) + (r >> 3);   // 3 dependent ops: LOP, SHF, IADD (SHF and LOP in parallel)
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[2] = t1 - t0;
    sink[threadIdx.x] = r;
}

__global__ void k_flo(long long* out, uint32_t* sink, uint32_t c, int iters)
{
    uint32_t r = threadIdx.x + 12345u;
    const long long t0 = clock64();
    #pragma unroll 16
    for (int i = 0; i < iters; i++) r = (r << (__clz((int)(r | 1u)) & 7)) ^ c;   // FLO + SHF + LOP dependent
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[3] = t1 - t0;
    sink[threadIdx.x] = r;
}

int main()
{
    long long* out;
    uint32_t* sink;
    cudaMallocManaged(&out, 64);
    cudaMalloc(&sink, 4096);
    const int iters = 1 << 16;
    for (int rep = 0; rep < 2; rep++) {
        k_fp64<<<1, 32>>>(out, sink, 0.7312345, iters);
        k_fixed<<<1, 32>>>(out, sink, (uint64_t)(0.7312345 * 281474976710656.0), iters);
        k_add<<<1, 32>>>(out, sink, 0x9E3779B9u, iters);
        k_flo<<<1, 32>>>(out, sink, 0x9E3779B9u, iters);
        cudaDeviceSynchronize();
    }
    printf("fp64 chain (I2F.F64 + DMUL + F2I.F64 + LOP + IADD): %.1f cycles/iter\n", (double)out[0] / iters);
    printf("fixed-point chain (IMAD.WIDE + IMAD + SHF + LOP + IADD): %.1f cycles/iter\n", (double)out[1] / iters);
    printf("xor/shift/add chain (2 dependent levels): %.1f cycles/iter\n", (double)out[2] / iters);
    printf("clz/shift/xor chain (FLO + LOP + SHF + LOP): %.1f cycles/iter\n", (double)out[3] / iters);
    return 0;
}
