// Shared host/device helpers for libeae_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/eae_b200.h"

namespace eae {

// Thread-local last-error text (eae_last_error()).
void set_error(const char* fmt, ...);
// Counts kernel launches made by this library (eae_launch_count()).
void count_launch(int n = 1);

#define EAE_CUDA_OK(expr)                                                                     \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ::eae::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                             __LINE__);                                                       \
            return EAE_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#define EAE_TRY(expr)              \
    do {                           \
        int _r = (expr);           \
        if (_r != 0) return _r;    \
    } while (0)

// After a kernel launch: catches launch-configuration errors without synchronising.
#define EAE_LAUNCH_OK()                                                                           \
    do {                                                                                          \
        cudaError_t _e = cudaGetLastError();                                                      \
        ::eae::count_launch();                                                                    \
        if (_e != cudaSuccess) {                                                                  \
            ::eae::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, \
                             __LINE__);                                                           \
            return EAE_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

static inline uint32_t ceil_div_u32(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

// RAII device buffer used by the _host entry points.
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t n)
    {
        if (p) { cudaFree(p); p = nullptr; }
        bytes = n;
        if (n == 0) n = 16;
        cudaError_t e = cudaMalloc(&p, n);
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("cudaMalloc(%zu) failed: %s", n, cudaGetErrorString(e));
            return EAE_ERR_CUDA;
        }
        return 0;
    }
    template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};

// Kernel classes of the optional device-side profile (eae_profile_*).
enum ProfClass : int {
    kProfGemmConv = 0, kProfGemmTconv, kProfGemmGdn, kProfGemmThin, kProfIm2col, kProfCol2im, kProfQuantize,
    kProfDequantize, kProfCoderEncode, kProfCoderDecode, kProfPack, kProfBinarize, kProfHist, kProfCount
};
// Brackets the launches issued during its lifetime with a CUDA event pair when profiling is on.
struct ProfScope {
    ProfScope(int cls, cudaStream_t st);
    ~ProfScope();
    void close();      // ends the scope early (scopes may nest: a class can be a part of another)
    int slot;
    cudaStream_t stream;
};

// The tensor kernels need (almost) the whole shared memory of an SM, i.e. the largest carveout. An SM whose resident
// CTAs were launched under a smaller carveout cannot take such a CTA until it has drained and been reconfigured - and
// the arithmetic coder's CTAs live for milliseconds. So every kernel that runs beside the tensor kernels in the
// multi-stream pipeline asks for the same (maximum) carveout. first_use_on_device: true the first time it is called
// with this mask on the current device.
bool first_use_on_device(uint64_t* seen_mask);
template <typename K> inline void prefer_max_shared(K kernel)
{
    cudaFuncSetAttribute(reinterpret_cast<const void*>(kernel), cudaFuncAttributePreferredSharedMemoryCarveout,
                         (int)cudaSharedmemCarveoutMaxShared);
}

// Is the per-kernel-class device profile (eae_profile_enable) recording?
bool profiling_enabled();

// Returns 0 if a CUDA device is usable, EAE_ERR_CUDA (with message) otherwise.
int require_device();

}  // namespace eae
