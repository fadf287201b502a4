// Runtime plumbing of libeae_b200.so: last-error text, device checks, pinned/device memory, streams
// and events for callers (ctypes) that have no CUDA binding of their own.
#include <stdlib.h>
#include <atomic>
#include <mutex>
#include <vector>
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

bool first_use_on_device(uint64_t* seen_mask)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return false; }
    const uint64_t bit = 1ull << (dev & 63);
    if (*seen_mask & bit) return false;
    *seen_mask |= bit;
    return true;
}

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

// ---- profile ----
struct ProfEvent { cudaEvent_t a, b; int cls; };
static std::mutex g_prof_mutex;
static std::vector<ProfEvent> g_prof_events;
static std::atomic<int> g_prof_on{0};
bool profiling_enabled() { return g_prof_on.load(std::memory_order_relaxed) != 0; }
static uint64_t g_prof_launches[kProfCount];
static double g_prof_ms[kProfCount];

ProfScope::ProfScope(int cls, cudaStream_t st) : slot(-1), stream(st)
{
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    ProfEvent ev;
    ev.cls = cls;
    if (cudaEventCreate(&ev.a) != cudaSuccess || cudaEventCreate(&ev.b) != cudaSuccess) return;
    cudaEventRecord(ev.a, st);
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    g_prof_events.push_back(ev);
    slot = (int)g_prof_events.size() - 1;
}

void ProfScope::close()
{
    if (slot < 0) return;
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    cudaEventRecord(g_prof_events[slot].b, stream);
    slot = -1;
}

ProfScope::~ProfScope() { close(); }

static void prof_drain()
{
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    for (ProfEvent& ev : g_prof_events) {
        float ms = 0.f;
        if (cudaEventSynchronize(ev.b) == cudaSuccess && cudaEventElapsedTime(&ms, ev.a, ev.b) == cudaSuccess) {
            g_prof_launches[ev.cls]++;
            g_prof_ms[ev.cls] += ms;
        }
        cudaEventDestroy(ev.a);
        cudaEventDestroy(ev.b);
    }
    g_prof_events.clear();
}

}  // namespace eae

using namespace eae;

extern "C" int eae_set_device(int device)
{
    EAE_TRY(require_device());
    EAE_CUDA_OK(cudaSetDevice(device));
    return 0;
}

extern "C" int eae_set_blocking_sync(int on)
{
    EAE_TRY(require_device());
    EAE_CUDA_OK(cudaSetDeviceFlags(on ? cudaDeviceScheduleBlockingSync : cudaDeviceScheduleAuto));
    return 0;
}

extern "C" int eae_profile_enable(int on)
{
    if (!on) prof_drain();
    g_prof_on.store(on ? 1 : 0);
    return 0;
}

extern "C" int eae_profile_reset(void)
{
    prof_drain();
    for (int i = 0; i < kProfCount; i++) { g_prof_launches[i] = 0; g_prof_ms[i] = 0.; }
    return 0;
}

extern "C" int eae_profile_read(int cls, uint64_t* launches, double* total_ms)
{
    if (cls < 0 || cls >= kProfCount) { set_error("bad kernel class %d", cls); return EAE_ERR_ARGUMENT; }
    prof_drain();
    if (launches) *launches = g_prof_launches[cls];
    if (total_ms) *total_ms = g_prof_ms[cls];
    return 0;
}

extern "C" const char* eae_profile_name(int cls)
{
    static const char* names[kProfCount] = {"gemm_conv", "gemm_tconv", "gemm_gdn", "gemm_thin", "im2col", "col2im",
                                            "quantize", "dequantize", "coder_encode", "coder_decode", "pack", "binarize", "hist"};
    return (cls >= 0 && cls < kProfCount) ? names[cls] : "";
}

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
    const char* pr = getenv("EAE_STREAM_PRIORITY");      // debug: "high" creates the streams at the greatest priority
    if (pr && pr[0] == 'h') {
        int least = 0, greatest = 0;
        EAE_CUDA_OK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        EAE_CUDA_OK(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, greatest));
    } else {
        EAE_CUDA_OK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    }
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
