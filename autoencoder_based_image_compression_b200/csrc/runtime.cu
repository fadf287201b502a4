// Runtime plumbing of libeae_b200.so: last-error text, device checks, pinned/device memory, streams
// and events for callers (ctypes) that have no CUDA binding of their own.
#include <atomic>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace eae {

static thread_local char g_error[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int require_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); libeae_b200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return EAE_ERR_CUDA;
    }
    return 0;
}

}  // namespace eae

using namespace eae;

extern "C" const char* eae_last_error(void) { return g_error; }

extern "C" int eae_abi_version(void) { return 1; }

extern "C" int eae_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int eae_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem)
{
    EAE_TRY(require_device());
    cudaDeviceProp prop;
    EAE_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (total_mem) *total_mem = prop.totalGlobalMem;
    return 0;
}

extern "C" uint64_t eae_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" void* eae_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (require_device()) return nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault);
    if (e != cudaSuccess) { set_error("cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return nullptr; }
    return p;
}

extern "C" void eae_host_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" void* eae_device_alloc(size_t bytes)
{
    void* p = nullptr;
    if (require_device()) return nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
    if (e != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return nullptr; }
    return p;
}

extern "C" void eae_device_free(void* p) { if (p) cudaFree(p); }

extern "C" int eae_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream)
{
    if (!dst || !src) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    EAE_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return 0;
}

extern "C" int eae_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream)
{
    if (!dst || !src) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    EAE_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return 0;
}

extern "C" int eae_stream_create(void** stream)
{
    if (!stream) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    EAE_TRY(require_device());
    cudaStream_t s;
    EAE_CUDA_OK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (void*)s;
    return 0;
}

extern "C" int eae_stream_destroy(void* stream)
{
    if (stream) EAE_CUDA_OK(cudaStreamDestroy((cudaStream_t)stream));
    return 0;
}

extern "C" int eae_stream_synchronize(void* stream)
{
    EAE_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

extern "C" int eae_event_create(void** event)
{
    if (!event) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    EAE_TRY(require_device());
    cudaEvent_t ev;
    EAE_CUDA_OK(cudaEventCreate(&ev));
    *event = (void*)ev;
    return 0;
}

extern "C" int eae_event_destroy(void* event)
{
    if (event) EAE_CUDA_OK(cudaEventDestroy((cudaEvent_t)event));
    return 0;
}

extern "C" int eae_event_record(void* event, void* stream)
{
    if (!event) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    EAE_CUDA_OK(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream));
    return 0;
}

extern "C" int eae_event_elapsed_ms(void* start, void* stop, float* ms)
{
    if (!start || !stop || !ms) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    EAE_CUDA_OK(cudaEventSynchronize((cudaEvent_t)stop));
    EAE_CUDA_OK(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return 0;
}
