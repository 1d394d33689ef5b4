// k_trig.cu -- the one-kernel cosft1 / cosft2 / sinft and twofft of trig_fused.cuh, built for 8 .. 8192 complex points per line
#include <cuda_runtime.h>

#include "trig_fused.cuh"
#include "plan.h"

namespace nrb {

template <int LOG2N>
__global__ void __launch_bounds__(GeoT<LOG2N>::NT, trig_min_ctas(LOG2N)) trig_kernel(const __grid_constant__ TrigParams T)
{
    extern __shared__ double2 nrb_trig_smem[];
    trig_cta<LOG2N>(T, nrb_trig_smem, blockIdx.x, (int)threadIdx.x);
}

template <int LOG2N>
__global__ void __launch_bounds__(GeoT<LOG2N>::NT, trig_min_ctas(LOG2N)) twofft_kernel(const __grid_constant__ TwoFFTParams T)
{
    extern __shared__ double2 nrb_trig_smem[];
    twofft_cta<LOG2N>(T, nrb_trig_smem, blockIdx.x, (int)threadIdx.x);
}

template <class K> static int prepare(K kernel, size_t smem, bool *ready)
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (!ready[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        ready[dev & 63] = true;
    }
    return 0;
}

template <int LOG2N> static int launch_trig_n(const TrigParams &t, cudaStream_t s)
{
    typedef GeoT<LOG2N> G;
    static bool ready[64] = {false};
    if (int rc = prepare(trig_kernel<LOG2N>, G::SMEM_BYTES, ready)) return rc;
    const u64 ntiles = (t.count + G::L - 1) / G::L;
    if (ntiles == 0) return 0;
    if (ntiles > 0x7fffffffull) return (int)cudaErrorInvalidConfiguration;
    trig_kernel<LOG2N><<<(unsigned)ntiles, G::NT, G::SMEM_BYTES, s>>>(t);
    return (int)cudaGetLastError();
}

template <int LOG2N> static int launch_twofft_n(const TwoFFTParams &t, cudaStream_t s)
{
    typedef GeoT<LOG2N> G;
    static bool ready[64] = {false};
    if (int rc = prepare(twofft_kernel<LOG2N>, G::SMEM_BYTES, ready)) return rc;
    const u64 ntiles = (t.count + G::L - 1) / G::L;
    if (ntiles == 0) return 0;
    if (ntiles > 0x7fffffffull) return (int)cudaErrorInvalidConfiguration;
    twofft_kernel<LOG2N><<<(unsigned)ntiles, G::NT, G::SMEM_BYTES, s>>>(t);
    return (int)cudaGetLastError();
}

#define NRB_TRIG_CASES(F, arg)                                                                                         \
    switch (log2n) {                                                                                                   \
    case 3: return F<3>(arg, s);                                                                                       \
    case 4: return F<4>(arg, s);                                                                                       \
    case 5: return F<5>(arg, s);                                                                                       \
    case 6: return F<6>(arg, s);                                                                                       \
    case 7: return F<7>(arg, s);                                                                                       \
    case 8: return F<8>(arg, s);                                                                                       \
    case 9: return F<9>(arg, s);                                                                                       \
    case 10: return F<10>(arg, s);                                                                                     \
    case 11: return F<11>(arg, s);                                                                                     \
    case 12: return F<12>(arg, s);                                                                                     \
    case 13: return F<13>(arg, s);                                                                                     \
    default: return (int)cudaErrorInvalidValue;                                                                        \
    }

int launch_trig(int log2n, const TrigParams &t, cudaStream_t s) { NRB_TRIG_CASES(launch_trig_n, t) }
int launch_twofft(int log2n, const TwoFFTParams &t, cudaStream_t s) { NRB_TRIG_CASES(launch_twofft_n, t) }

} // namespace nrb
