// k_mid.cu -- the fused middle kernel of the long-line convlv / correl pipeline (conv_mid.cuh), built for rows of
// 4096 points (the production size: conv_split always leaves 4096-point rows) and for 16 / 64 points (small cases the
// tests can reach by lowering row_max_log2)
#include <cuda_runtime.h>

#include "conv_mid.cuh"
#include "plan.h"

namespace nrb {

template <int LOG2R>
__global__ void __launch_bounds__(GeoM<LOG2R>::NT, (LOG2R <= 11 ? 2 : 1)) conv_mid_kernel(const __grid_constant__ ConvMidParams M)
{
    extern __shared__ double2 nrb_mid_smem[];
    conv_mid_cta<LOG2R>(M, nrb_mid_smem, blockIdx.x, (int)threadIdx.x);
}

template <int LOG2R> static int launch_mid(const ConvMidParams &m, u64 ntiles, cudaStream_t s)
{
    constexpr size_t smem = GeoM<LOG2R>::SMEM_BYTES;
    static bool ready[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!ready[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(conv_mid_kernel<LOG2R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        ready[dev & 63] = true;
    }
    if (ntiles == 0) return 0;
    if (ntiles > 0x7fffffffull) return (int)cudaErrorInvalidConfiguration;
    conv_mid_kernel<LOG2R><<<(unsigned)ntiles, GeoM<LOG2R>::NT, smem, s>>>(m);
    return (int)cudaGetLastError();
}

int launch_conv_mid(int log2rest, const ConvMidParams &m, u64 ntiles, cudaStream_t s)
{
    switch (log2rest) {
    case 4: return launch_mid<4>(m, ntiles, s);
    case 6: return launch_mid<6>(m, ntiles, s);
    case 11: return launch_mid<11>(m, ntiles, s);
    case 12: return launch_mid<12>(m, ntiles, s);
    default: return (int)cudaErrorInvalidValue;
    }
}

} // namespace nrb
