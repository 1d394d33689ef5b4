// kernels_inst.cuh -- __global__ wrappers and per-translation-unit kernel registration.
// Each k_*.cu instantiates a slice of the (log2n, layout, dir, variant) space so the
// instantiations compile in parallel.
#pragma once
#include <cuda_runtime.h>

#include "aux_kernels.cuh"
#include "plan.h"

// resident CTAs per SM the register allocator targets for the 4096-point tiles
#ifndef NRB_MIN_BLOCKS_ROW
#define NRB_MIN_BLOCKS_ROW 2
#endif
#ifndef NRB_MIN_BLOCKS_COL
#define NRB_MIN_BLOCKS_COL 2
#endif

namespace nrb {

typedef int (*PassLaunchFn)(const PassParams &, u64, cudaStream_t);
// table[log2n][layout][dir>0][variant]
struct PassTable { PassLaunchFn fn[kMaxLog2N + 1][2][2][3]; };
PassTable &pass_table();

// NRB_PERSISTENT=1 sizes the grid to what is resident at once (SMs x CTAs/SM) and lets every CTA walk
// tiles blockIdx.x, blockIdx.x + gridDim.x, ...  Measured on B200 (profiles/r01_tuning.md) it is 3-7 %
// SLOWER than one CTA per tile for these kernels (the barrier between tiles serialises what the block
// scheduler otherwise overlaps), so the default is one CTA per tile.
#ifndef NRB_PERSISTENT
#define NRB_PERSISTENT 0
#endif

template <int LOG2N, int LAYOUT, int DIR, int VARIANT>
__global__ void __launch_bounds__(cta_threads(LOG2N, LAYOUT), (tile_log2(LOG2N) == 12 ? (LAYOUT == LAYOUT_ROW ? NRB_MIN_BLOCKS_ROW : NRB_MIN_BLOCKS_COL) : 1))
fft_pass_kernel(const __grid_constant__ PassParams P, const unsigned ntiles)
{
    extern __shared__ double2 nrb_smem[];
    for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        fft_pass_body<LOG2N, LAYOUT, DIR, VARIANT>(P, nrb_smem, tile, (int)threadIdx.x);
        __syncthreads();   // shared memory is reused by the next tile
    }
}

template <int LOG2N, int LAYOUT, int DIR, int VARIANT>
int launch_pass_t(const PassParams &p, u64 ntiles, cudaStream_t s)
{
    constexpr size_t smem = smem_elems(LOG2N, LAYOUT, VARIANT) * sizeof(double2);
    constexpr int NT = cta_threads(LOG2N, LAYOUT);
    static int resident[64] = {0};   // CTAs resident on the whole device, per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!resident[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(fft_pass_kernel<LOG2N, LAYOUT, DIR, VARIANT>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        int per_sm = 0, sms = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fft_pass_kernel<LOG2N, LAYOUT, DIR, VARIANT>, NT, smem);
        if (e != cudaSuccess) return (int)e;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        resident[dev & 63] = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 148);
    }
    if (ntiles == 0) return 0;
    u64 grid = ntiles;
    if (NRB_PERSISTENT && grid > (u64)resident[dev & 63]) grid = (u64)resident[dev & 63];
    fft_pass_kernel<LOG2N, LAYOUT, DIR, VARIANT><<<(unsigned)grid, NT, smem, s>>>(p, (unsigned)ntiles);
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
