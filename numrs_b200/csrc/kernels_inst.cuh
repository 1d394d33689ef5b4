// kernels_inst.cuh -- __global__ wrappers and per-translation-unit kernel registration.
// Each k_*.cu instantiates a slice of the (log2n, layout, dir, variant) space so the
// instantiations compile in parallel.
#pragma once
#include <cuda_runtime.h>

#include "aux_kernels.cuh"
#include "fft_pass2.cuh"
#include "plan.h"

// resident CTAs per SM the register allocator targets for the 4096-point tiles
#ifndef NRB_MIN_BLOCKS_ROW
#define NRB_MIN_BLOCKS_ROW 2
#endif
#ifndef NRB_MIN_BLOCKS_COL
#define NRB_MIN_BLOCKS_COL 2
#endif

namespace nrb {

// CTAs per SM the register allocator targets: the 64 K registers of an SM divided by
// (threads x registers) with 64 registers at 8 points/thread and 128 at 16
NRB_HD constexpr int min_blocks(int log2n, int layout)
{
    return tile_log2(log2n, layout) > 12 ? 1
         : layout == LAYOUT_ROW ? NRB_MIN_BLOCKS_ROW * (4096 >> tile_log2(log2n, layout))
                                : NRB_MIN_BLOCKS_COL;
}

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
__global__ void __launch_bounds__(cta_threads(LOG2N, LAYOUT), min_blocks(LOG2N, LAYOUT))
fft_pass_kernel(const __grid_constant__ PassParams P, const unsigned ntiles)
{
    extern __shared__ double2 nrb_smem[];
    for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (P.prefetch_dist > 0 && tile + (unsigned)P.prefetch_dist < ntiles)
            prefetch_tile<LOG2N, LAYOUT, VARIANT>(P, tile + (unsigned)P.prefetch_dist, (int)threadIdx.x);
        if constexpr (simple_built(LOG2N, LAYOUT, VARIANT)) {
            if (P.simple) fft_pass_body<LOG2N, LAYOUT, DIR, VARIANT, true>(P, nrb_smem, tile, (int)threadIdx.x);
            else fft_pass_body<LOG2N, LAYOUT, DIR, VARIANT, false>(P, nrb_smem, tile, (int)threadIdx.x);
        } else {
            fft_pass_body<LOG2N, LAYOUT, DIR, VARIANT, false>(P, nrb_smem, tile, (int)threadIdx.x);
        }
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
    if (p.grid_cap > 0 && grid > (u64)p.grid_cap) grid = (u64)p.grid_cap;
    fft_pass_kernel<LOG2N, LAYOUT, DIR, VARIANT><<<(unsigned)grid, NT, smem, s>>>(p, (unsigned)ntiles);
    return (int)cudaGetLastError();
}

// ---- two dependent passes in one persistent launch (see FuseSched in nrb_common.h) ----
typedef int (*FusedLaunchFn)(const PassParams &, const PassParams &, const FuseSched &, cudaStream_t);
NRB_HD constexpr unsigned long long fused_key(int la, int lya, int va, int lb, int lyb, int vb, int dir)
{
    return ((unsigned long long)la << 40) | ((unsigned long long)lya << 36) | ((unsigned long long)va << 32) |
           ((unsigned long long)lb << 24) | ((unsigned long long)lyb << 20) | ((unsigned long long)vb << 16) | (dir > 0 ? 1ull : 0ull);
}
void register_fused(unsigned long long key, FusedLaunchFn fn);

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int LA, int LYA, int VA, int LB, int LYB, int VB, int DIR>
__global__ void __launch_bounds__(cta_threads(LA, LYA), 2)
fused_pass_kernel(const __grid_constant__ PassParams PA, const __grid_constant__ PassParams PB, const FuseSched F)
{
    static_assert(cta_threads(LA, LYA) == cta_threads(LB, LYB), "fused passes must use the same CTA size");
    extern __shared__ double2 nrb_smem[];
    __shared__ unsigned long long s_ticket;
    const int tid = (int)threadIdx.x;
    const unsigned long long per = (unsigned long long)F.ta + F.tb;
    const unsigned long long head = (unsigned long long)F.lag * F.ta;                    // A-only tickets
    const unsigned long long mid = (unsigned long long)(F.units - F.lag) * per;          // interleaved
    const unsigned long long total = (unsigned long long)F.units * per;
    for (;;) {
        if (tid == 0) s_ticket = atomicAdd(F.ticket, 1ull);
        __syncthreads();
        const unsigned long long t = s_ticket;
        __syncthreads();
        if (t >= total) break;
        bool is_a;
        unsigned unit, r;
        if (t < head) { is_a = true; unit = (unsigned)(t / F.ta); r = (unsigned)(t % F.ta); }
        else if (t < head + mid) {
            const unsigned long long u = t - head;
            const unsigned i = F.lag + (unsigned)(u / per);
            const unsigned rr = (unsigned)(u % per);
            if (rr < F.ta) { is_a = true; unit = i; r = rr; }
            else { is_a = false; unit = i - F.lag; r = rr - F.ta; }
        } else {
            const unsigned long long u = t - head - mid;
            is_a = false; unit = F.units - F.lag + (unsigned)(u / F.tb); r = (unsigned)(u % F.tb);
        }
        if (is_a) {
            fft_pass_body<LA, LYA, DIR, VA>(PA, nrb_smem, unit * F.ta + r, tid);
            __threadfence();                      // this thread's stores are visible device-wide ...
            __syncthreads();                      // ... for every thread of the CTA ...
            if (tid == 0) atomicAdd(&F.done[unit], 1u);   // ... before the tile is published
        } else {
            if (tid == 0) {
                while (ld_acquire_u32(&F.done[unit]) < F.ta) __nanosleep(64);
            }
            __syncthreads();
            fft_pass_body<LB, LYB, DIR, VB>(PB, nrb_smem, unit * F.tb + r, tid);
        }
    }
}

template <int LA, int LYA, int VA, int LB, int LYB, int VB, int DIR>
int launch_fused_t(const PassParams &pa, const PassParams &pb, const FuseSched &fs, cudaStream_t s)
{
    constexpr size_t sa = smem_elems(LA, LYA, VA) * sizeof(double2), sb = smem_elems(LB, LYB, VB) * sizeof(double2);
    constexpr size_t smem = sa > sb ? sa : sb;
    constexpr int NT = cta_threads(LA, LYA);
    auto kern = fused_pass_kernel<LA, LYA, VA, LB, LYB, VB, DIR>;
    static int resident[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!resident[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        int per_sm = 0, sms = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem);
        if (e != cudaSuccess) return (int)e;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        resident[dev & 63] = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 148);
    }
    cudaError_t e = cudaMemsetAsync(fs.ticket, 0, 16 + 4 * (size_t)((fs.units + 3) & ~3u), s);   // ticket + done[] are contiguous
    if (e != cudaSuccess) return (int)e;
    unsigned long long total = (unsigned long long)fs.units * ((unsigned long long)fs.ta + fs.tb);
    unsigned grid = (unsigned)(total < (unsigned long long)resident[dev & 63] ? total : (unsigned long long)resident[dev & 63]);
    if (grid == 0) return 0;
    kern<<<grid, NT, smem, s>>>(pa, pb, fs);
    return (int)cudaGetLastError();
}

// rlft3: z real pass (ROW REAL, 2^LZ) + y pass (COL PLAIN, 2^LY); forward = z then y, inverse = y then z
template <int LZ, int LY> void register_fused_zy()
{
    if constexpr (cta_threads(LZ, LAYOUT_ROW) != cta_threads(LY, LAYOUT_COL)) return;   // needs equal CTA sizes
    else {
    register_fused(fused_key(LZ, LAYOUT_ROW, VAR_REAL, LY, LAYOUT_COL, VAR_PLAIN, +1),
                   launch_fused_t<LZ, LAYOUT_ROW, VAR_REAL, LY, LAYOUT_COL, VAR_PLAIN, +1>);
    register_fused(fused_key(LY, LAYOUT_COL, VAR_PLAIN, LZ, LAYOUT_ROW, VAR_REAL, -1),
                   launch_fused_t<LY, LAYOUT_COL, VAR_PLAIN, LZ, LAYOUT_ROW, VAR_REAL, -1>);
    }
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

// ---- big-tile passes (fft_pass2.cuh): persistent CTAs, one per resident slot ----
PassTable &pass2_table();

template <int LOG2N, int LAYOUT, int DIR, int VARIANT>
__global__ void __launch_bounds__(Geo2<LOG2N, LAYOUT, VARIANT>::NT, (Geo2<LOG2N, LAYOUT, VARIANT>::TL > 12 ? 1 : 2))
fft_pass2_kernel(const __grid_constant__ PassParams P, const unsigned ntiles)
{
    extern __shared__ double2 nrb_smem[];
    fft_pass2_cta<LOG2N, LAYOUT, DIR, VARIANT>(P, nrb_smem, blockIdx.x, gridDim.x, ntiles, (int)threadIdx.x);
}

// `ntiles_v1` is ignored: the tile count follows from the lines of the pass and this kernel's own geometry
template <int LOG2N, int LAYOUT, int DIR, int VARIANT>
int launch_pass2_t(const PassParams &p, u64, cudaStream_t s)
{
    typedef Geo2<LOG2N, LAYOUT, VARIANT> G;
    constexpr size_t smem = G::SMEM_BYTES;
    auto kern = fft_pass2_kernel<LOG2N, LAYOUT, DIR, VARIANT>;
    static int resident[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!resident[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        int per_sm = 0, sms = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, G::NT, smem);
        if (e != cudaSuccess) return (int)e;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        resident[dev & 63] = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 148);
    }
    const u64 lines = p.q_end - p.q_begin;
    const u64 ntiles = (lines + G::L - 1) / G::L;
    if (ntiles == 0) return 0;
    if (ntiles > 0x7fffffffull) return (int)cudaErrorInvalidConfiguration;
    const u64 grid = (NRB_V2_DIRECT || ntiles < (u64)resident[dev & 63]) ? ntiles : (u64)resident[dev & 63];
    kern<<<(unsigned)grid, G::NT, smem, s>>>(p, (unsigned)ntiles);
    return (int)cudaGetLastError();
}

template <int LOG2N, int LAYOUT> void register_size2(PassTable &t)
{
    if constexpr (LAYOUT == LAYOUT_ROW) {
        t.fn[LOG2N][LAYOUT_ROW][1][VAR_PLAIN] = launch_pass2_t<LOG2N, LAYOUT_ROW, +1, VAR_PLAIN>;
        t.fn[LOG2N][LAYOUT_ROW][0][VAR_PLAIN] = launch_pass2_t<LOG2N, LAYOUT_ROW, -1, VAR_PLAIN>;
    } else {
        t.fn[LOG2N][LAYOUT_COL][1][VAR_PLAIN] = launch_pass2_t<LOG2N, LAYOUT_COL, +1, VAR_PLAIN>;
        t.fn[LOG2N][LAYOUT_COL][0][VAR_PLAIN] = launch_pass2_t<LOG2N, LAYOUT_COL, -1, VAR_PLAIN>;
        t.fn[LOG2N][LAYOUT_COL][1][VAR_XPOSE] = launch_pass2_t<LOG2N, LAYOUT_COL, +1, VAR_XPOSE>;
        t.fn[LOG2N][LAYOUT_COL][0][VAR_XPOSE] = launch_pass2_t<LOG2N, LAYOUT_COL, -1, VAR_XPOSE>;
    }
}

} // namespace nrb
