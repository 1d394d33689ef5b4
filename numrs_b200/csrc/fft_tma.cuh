// fft_tma.cuh -- the strided (COL) FFT pass with its tile moved by the Tensor Memory Accelerator (sm_100a).
//
// fft_pass.cuh's COL pass loads its tile with 16-byte LDG instructions straight into registers (first stage) and stores
// it with STG from the last stage: every thread computes ~13 integer instructions of address arithmetic per element and
// the L1/TEX pipe sees one wavefront per 128-byte run.  Here ONE elected thread describes the whole tile to the TMA unit
// (`cp.async.bulk.tensor.3d`, box = 2L doubles x 256 rows, two or four boxes per tile), the tile lands in shared memory
// in exactly the [n][line] layout the Stockham stages use, an mbarrier tells the CTA when the bytes are there, all NST
// stages run shared -> shared, and the finished tile goes back with one bulk tensor store.  No per-thread global
// addresses at all; the price is two more passes over shared memory than the register-fed version (NST + 1 instead of
// NST - 1 round trips).  A/B against fft_pass.cuh: option `tma_col_mask` (bit log2 N), profiles/r02_tuning.md.
//
// Eligible launches (be_launch_pass checks): LAYOUT_COL, VAR_PLAIN, in place or out of place with the same geometry, no
// four-step twiddle, no split element index, lines = [outer][inner] with inner a multiple of the tile's line count.
// The tensor is described in doubles: dim0 = 2 * inner (re, im interleaved, contiguous), dim1 = N (stride inner * 16 B),
// dim2 = outer (stride N * inner * 16 B).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "fft_pass.cuh"

namespace nrb {

__device__ __forceinline__ unsigned tma_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tma_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tma_mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tma_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(tma_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(tma_smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(tma_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, int c0, int c1, int c2, const void *src)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(tma_smem_u32(src)) : "memory");
}

// One CTA per tile (2 CTAs per SM overlap each other's copy and compute phases), or `PERSIST`: a CTA walks tiles
// blockIdx.x, + gridDim.x, ... with TWO tile buffers, so the bulk load of its next tile is in flight while it computes.
// Resident CTAs per SM: the tile (64 KiB) allows three; the 256-thread geometries (16 points per thread: N = 512, 64) fit
// three in the register file at 80 registers per thread, the 512-thread ones two at 64.  Three matter: with two, 35 % of the
// warp samples sit in the mbarrier wait for the bulk load (ncu, profiles/r02_ncu_full_rlft3_512_tma.md) -- y / x pass of
// rlft3 512^3 5 416 / 5 296 GB/s with two CTAs, 6 417 / 5 931 GB/s with three (register-fed kernel: 6 272 / 5 613).
NRB_HD constexpr int tma_ctas_per_sm(int log2n) { return cta_threads(log2n, LAYOUT_COL) <= 256 ? 3 : 2; }

__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap *map, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

template <int LOG2N, int DIR, bool PERSIST>
__global__ void __launch_bounds__(cta_threads(LOG2N, LAYOUT_COL), PERSIST ? 1 : tma_ctas_per_sm(LOG2N))
fft_col_tma_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
                   const __grid_constant__ PassParams P, const unsigned ntiles, const unsigned log2_inner)
{
    typedef Geo<LOG2N, LAYOUT_COL, VAR_PLAIN> G;
    constexpr int ROWS = G::N < 256 ? G::N : 256;          // rows per box (boxDim <= 256)
    constexpr int NBOX = G::N / ROWS;
    constexpr unsigned TILE_BYTES = (unsigned)G::TILE * 16u;
    extern __shared__ __align__(128) double2 nrb_tma_smem[];
    __shared__ __align__(8) unsigned long long bar[2];
    const int tid = (int)threadIdx.x;
    if (tid == 0) {
        tma_mbar_init(&bar[0], 1);
        tma_mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned inner_mask = (1u << log2_inner) - 1u;
    auto coords = [&](unsigned tile, int &c0, int &c2) {
        const unsigned long long q0 = P.q_begin + (unsigned long long)tile * G::L;
        c0 = 2 * (int)((unsigned)q0 & inner_mask);
        c2 = (int)(q0 >> log2_inner);
    };
    auto issue_load = [&](unsigned tile, int buf) {
        int c0, c2;
        coords(tile, c0, c2);
        tma_mbar_expect_tx(&bar[buf], TILE_BYTES);
#pragma unroll
        for (int b = 0; b < NBOX; ++b) tma_load_3d(nrb_tma_smem + (size_t)buf * G::TILE + (size_t)b * ROWS * G::L, &tm_in, c0, b * ROWS, c2, &bar[buf]);
    };
    unsigned phase[2] = {0u, 0u};
    int buf = 0;
    if (tid == 0 && blockIdx.x < ntiles) issue_load(blockIdx.x, 0);
    if (!PERSIST && tid == 32 && P.prefetch_dist > 0 && blockIdx.x + (unsigned)P.prefetch_dist < ntiles) {
        // ask L2 for the tile a later CTA of this SM slot will load (two instructions for 64 KiB)
        int c0, c2;
        coords(blockIdx.x + (unsigned)P.prefetch_dist, c0, c2);
#pragma unroll
        for (int b = 0; b < NBOX; ++b) tma_prefetch_3d(&tm_in, c0, b * ROWS, c2);
    }
    for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        double2 *sm = nrb_tma_smem + (size_t)buf * G::TILE;
        if (PERSIST && tid == 0 && tile + gridDim.x < ntiles) {
            // the other buffer was handed to a bulk store one iteration ago: wait until that store has READ it
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            issue_load(tile + gridDim.x, buf ^ 1);
        }
        tma_mbar_wait(&bar[buf], phase[buf]);
        phase[buf] ^= 1u;
        StageRunner<LOG2N, LAYOUT_COL, DIR, VAR_PLAIN, 0, false, false, false, 3>::run(P, sm, tile, tid);
        // generic-proxy writes of the last stage -> visible to the async proxy, then one thread hands the tile to TMA
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            int c0, c2;
            coords(tile, c0, c2);
#pragma unroll
            for (int b = 0; b < NBOX; ++b) tma_store_3d(&tm_out, c0, b * ROWS, c2, sm + (size_t)b * ROWS * G::L);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (PERSIST) buf ^= 1;
        else break;
    }
    // the bulk stores must have READ the tile before the CTA's shared memory goes away (the global writes finish on their own)
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// TMA on the INPUT side only: the tile is loaded by bulk tensor copies, the stages run shared -> shared except the last,
// which stores to global memory from registers as in fft_pass.cuh -- so the output side may be anything the generic pass
// can address (four-step twiddle + transposed strides, split element index, exchange pointer tables), and the pass makes
// NST shared-memory round trips instead of the NST + 1 of the load-and-store version.
template <int LOG2N, int DIR, int CTAS>
__global__ void __launch_bounds__(cta_threads(LOG2N, LAYOUT_COL), CTAS)
fft_col_tma_in_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ PassParams P, const unsigned ntiles, const unsigned log2_inner)
{
    typedef Geo<LOG2N, LAYOUT_COL, VAR_PLAIN> G;
    constexpr int ROWS = G::N < 256 ? G::N : 256, NBOX = G::N / ROWS;
    extern __shared__ __align__(128) double2 nrb_tma_smem[];
    __shared__ __align__(8) unsigned long long bar;
    const int tid = (int)threadIdx.x;
    const unsigned tile = blockIdx.x;
    if (tile >= ntiles) return;
    if (tid == 0) {
        tma_mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned long long q0 = P.q_begin + (unsigned long long)tile * G::L;
        const int c0 = 2 * (int)((unsigned)q0 & ((1u << log2_inner) - 1u)), c2 = (int)(q0 >> log2_inner);
        tma_mbar_expect_tx(&bar, (unsigned)G::TILE * 16u);
#pragma unroll
        for (int b = 0; b < NBOX; ++b) tma_load_3d(nrb_tma_smem + (size_t)b * ROWS * G::L, &tm_in, c0, b * ROWS, c2, &bar);
    }
    __syncthreads();
    tma_mbar_wait(&bar, 0u);
    if constexpr (simple_built(LOG2N, LAYOUT_COL, VAR_PLAIN)) {
        if (P.simple) { fft_pass_body<LOG2N, LAYOUT_COL, DIR, VAR_PLAIN, true, true>(P, nrb_tma_smem, tile, tid); return; }
    }
    fft_pass_body<LOG2N, LAYOUT_COL, DIR, VAR_PLAIN, false, true>(P, nrb_tma_smem, tile, tid);
}

// The TRANSPOSING pass of a multi-step transform (VAR_XPOSE) with its strided loads done by TMA: only for 1024-point lines,
// whose 4-lines-per-tile shared layout is XOR-swizzled (Geo::phys: a ^ ((a >> 3) & 3) in 16-byte units) -- exactly the
// TMA's 64-byte swizzle for a 64-byte inner box, so the bulk copy can write the layout the stages expect.  The line-
// contiguous store with the four-step twiddle stays as it is (fft_pass_body's epilogue).
template <int DIR>
__global__ void __launch_bounds__(cta_threads(10, LAYOUT_COL), 2)
fft_xpose_tma_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ PassParams P, const unsigned ntiles, const unsigned log2_inner)
{
    typedef Geo<10, LAYOUT_COL, VAR_XPOSE> G;
    static_assert(G::L == 4, "the 64-byte swizzle matches the 4-lines-per-tile layout only");
    constexpr int ROWS = 256, NBOX = G::N / ROWS;
    extern __shared__ __align__(1024) double2 nrb_tma_smem[];
    __shared__ __align__(8) unsigned long long bar;
    const int tid = (int)threadIdx.x;
    const unsigned tile = blockIdx.x;
    if (tile >= ntiles) return;
    if (tid == 0) {
        tma_mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned long long q0 = P.q_begin + (unsigned long long)tile * G::L;
        const int c0 = 2 * (int)((unsigned)q0 & ((1u << log2_inner) - 1u)), c2 = (int)(q0 >> log2_inner);
        tma_mbar_expect_tx(&bar, (unsigned)G::TILE * 16u);
#pragma unroll
        for (int b = 0; b < NBOX; ++b) tma_load_3d(nrb_tma_smem + (size_t)b * ROWS * G::L, &tm_in, c0, b * ROWS, c2, &bar);
    }
    __syncthreads();
    tma_mbar_wait(&bar, 0u);
    if constexpr (simple_built(10, LAYOUT_COL, VAR_XPOSE)) {
        if (P.simple) { fft_pass_body<10, LAYOUT_COL, DIR, VAR_XPOSE, true, true>(P, nrb_tma_smem, tile, tid); return; }
    }
    fft_pass_body<10, LAYOUT_COL, DIR, VAR_XPOSE, false, true>(P, nrb_tma_smem, tile, tid);
}

} // namespace nrb
