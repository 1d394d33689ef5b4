// kernels_inst.cuh -- __global__ wrappers and per-translation-unit kernel registration.
// Each k_*.cu instantiates a slice of the (log2n, layout, dir, variant) space so the
// instantiations compile in parallel.
#pragma once
#include <cuda_runtime.h>

#include "aux_kernels.cuh"
#include "plan.h"

// resident CTAs per SM the register allocator targets for the 4096-point tiles
// (3 -> 80 registers/thread, 2 -> 128).  ROW kernels carry per-thread twiddles and spill at 80.
#ifndef NRB_MIN_BLOCKS_ROW
#define NRB_MIN_BLOCKS_ROW 2
#endif
#ifndef NRB_MIN_BLOCKS_COL
#define NRB_MIN_BLOCKS_COL (NRB_RADIX16 ? 2 : 3)
#endif

namespace nrb {

typedef int (*PassLaunchFn)(const PassParams &, u64, cudaStream_t);
// table[log2n][layout][dir>0][variant]
struct PassTable { PassLaunchFn fn[kMaxLog2N + 1][2][2][3]; };
PassTable &pass_table();

template <int LOG2N, int LAYOUT, int DIR, int VARIANT>
__global__ void __launch_bounds__(cta_threads(LOG2N), (tile_log2(LOG2N) == 12 ? (LAYOUT == LAYOUT_ROW ? NRB_MIN_BLOCKS_ROW : NRB_MIN_BLOCKS_COL) : 1))
fft_pass_kernel(const __grid_constant__ PassParams P)
{
    extern __shared__ double2 nrb_smem[];
    fft_pass_body<LOG2N, LAYOUT, DIR, VARIANT>(P, nrb_smem, blockIdx.x, (int)threadIdx.x);
}

template <int LOG2N, int LAYOUT, int DIR, int VARIANT>
int launch_pass_t(const PassParams &p, u64 ntiles, cudaStream_t s)
{
    constexpr size_t smem = smem_elems(LOG2N, LAYOUT, VARIANT) * sizeof(double2);
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(fft_pass_kernel<LOG2N, LAYOUT, DIR, VARIANT>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured[dev & 63] = true;
    }
    if (ntiles == 0) return 0;
    fft_pass_kernel<LOG2N, LAYOUT, DIR, VARIANT><<<(unsigned)ntiles, cta_threads(LOG2N), smem, s>>>(p);
    return (int)cudaGetLastError();
}

template <int LOG2N, int LAYOUT> void register_size(PassTable &t)
{
    if constexpr (LAYOUT == LAYOUT_ROW) {
        t.fn[LOG2N][LAYOUT_ROW][1][VAR_PLAIN] = launch_pass_t<LOG2N, LAYOUT_ROW, +1, VAR_PLAIN>;
        t.fn[LOG2N][LAYOUT_ROW][0][VAR_PLAIN] = launch_pass_t<LOG2N, LAYOUT_ROW, -1, VAR_PLAIN>;
        t.fn[LOG2N][LAYOUT_ROW][1][VAR_REAL] = launch_pass_t<LOG2N, LAYOUT_ROW, +1, VAR_REAL>;
        t.fn[LOG2N][LAYOUT_ROW][0][VAR_REAL] = launch_pass_t<LOG2N, LAYOUT_ROW, -1, VAR_REAL>;
    } else {
        t.fn[LOG2N][LAYOUT_COL][1][VAR_PLAIN] = launch_pass_t<LOG2N, LAYOUT_COL, +1, VAR_PLAIN>;
        t.fn[LOG2N][LAYOUT_COL][0][VAR_PLAIN] = launch_pass_t<LOG2N, LAYOUT_COL, -1, VAR_PLAIN>;
        t.fn[LOG2N][LAYOUT_COL][1][VAR_XPOSE] = launch_pass_t<LOG2N, LAYOUT_COL, +1, VAR_XPOSE>;
        t.fn[LOG2N][LAYOUT_COL][0][VAR_XPOSE] = launch_pass_t<LOG2N, LAYOUT_COL, -1, VAR_XPOSE>;
    }
}

} // namespace nrb
