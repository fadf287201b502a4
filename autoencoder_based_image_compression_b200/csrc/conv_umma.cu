// tcgen05 / TMEM / TMA tap-list GEMM (placeholder until the kernel lands).
#include "common.cuh"
#include "transforms.cuh"

namespace eae {

int umma_available()
{
    set_error("the tcgen05 path is not built into this library yet");
    return EAE_ERR_ARGUMENT;
}

int launch_gemm_umma(const GemmPlan&, const float*, bool, cudaStream_t)
{
    return umma_available();
}

}  // namespace eae
