// prefetch_row_bench.cu -- memory skeleton of the long-line ROW passes (N = 4096: 2 CTAs/SM, N = 8192: 1 CTA/SM):
// load a tile, NR round trips through shared memory (write / barrier / read / barrier, the Stockham exchange),
// store the tile.  The real kernels run AT this skeleton's speed (profiles/r01_tuning.md #7), i.e. they are bound
// by the load -> exchange -> store structure at low occupancy, not by math.
//
// Variant D ("direct", what the library does): one CTA per tile, first stage loads global memory into registers.
// Variant P ("prefetch"): persistent CTAs; every thread cp.async's the 16 elements IT will need for the next tile
// into thread-private shared-memory slots while the current tile is in its exchange rounds (no barrier and no
// mbarrier is needed for the landing zone: a thread waits for its own copies and reuses its own slots).
// The landing zone costs TILE*16 bytes, so the exchange buffer is halved by exchanging re and im parts
// separately (SPLIT): 8192-point tile = 128 KiB landing + 72 KiB exchange = 200 KiB, 1 CTA/SM;
// 4096-point tile = 64 + 36 = 100 KiB, 2 CTAs/SM.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void cp16(double2 *smem, const double2 *g)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(g) : "memory");
}

template <int TILE, int NT, int NR, bool PREFETCH, bool SPLIT>
__global__ void __launch_bounds__(NT) skel(double2 *data, long long ntiles)
{
    extern __shared__ double2 sm[];
    constexpr int PPT = TILE / NT;
    double2 *S = sm;                                               // landing zone [PPT][NT] (PREFETCH only)
    double2 *E = sm + (PREFETCH ? TILE : 0);                       // exchange buffer (padded 1 per 8)
    double *Ed = reinterpret_cast<double *>(E);
    const int tid = threadIdx.x;
    auto wr_idx = [&](int i) { const int n = tid + i * NT; return n + (n >> 3); };                       // Stockham-like write
    auto rd_idx = [&](int i) { const int n = tid * 8 + (i & 7) + (i >> 3) * (TILE / 2); return n + (n >> 3); };   // strided read-back
    auto rd_idx_s = [&](int i) { const int n = tid * 8 + (i & 7) + (i >> 3) * (TILE / 2); return n + (n >> 4); };
    auto wr_idx_s = [&](int i) { const int n = tid + i * NT; return n + (n >> 4); };

    if (PREFETCH) {
        const long long t0 = blockIdx.x;
        if (t0 < ntiles) {
#pragma unroll
            for (int i = 0; i < PPT; ++i) cp16(S + i * NT + tid, data + t0 * TILE + tid + i * NT);
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        double2 v[PPT];
        double2 *b = data + t * TILE;
        if (PREFETCH) {
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
#pragma unroll
            for (int i = 0; i < PPT; ++i) v[i] = S[i * NT + tid];
            const long long tn = t + gridDim.x;
            if (tn < ntiles) {
#pragma unroll
                for (int i = 0; i < PPT; ++i) cp16(S + i * NT + tid, data + tn * TILE + tid + i * NT);
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        } else {
#pragma unroll
            for (int i = 0; i < PPT; ++i) v[i] = __ldcg(b + tid + i * NT);
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            if (SPLIT) {
#pragma unroll
                for (int i = 0; i < PPT; ++i) Ed[wr_idx_s(i)] = v[i].x;
                __syncthreads();
#pragma unroll
                for (int i = 0; i < PPT; ++i) v[i].x = Ed[rd_idx_s(i)] + 1.0;
                __syncthreads();
#pragma unroll
                for (int i = 0; i < PPT; ++i) Ed[wr_idx_s(i)] = v[i].y;
                __syncthreads();
#pragma unroll
                for (int i = 0; i < PPT; ++i) v[i].y = Ed[rd_idx_s(i)];
                __syncthreads();
            } else {
#pragma unroll
                for (int i = 0; i < PPT; ++i) E[wr_idx(i)] = v[i];
                __syncthreads();
#pragma unroll
                for (int i = 0; i < PPT; ++i) { v[i] = E[rd_idx(i)]; v[i].x += 1.0; }
                __syncthreads();
            }
        }
#pragma unroll
        for (int i = 0; i < PPT; ++i) b[tid + i * NT] = v[i];
    }
}

template <int TILE, int NT, int NR, bool PREFETCH, bool SPLIT> void run(double2 *d, long long total, const char *tag)
{
    auto k = skel<TILE, NT, NR, PREFETCH, SPLIT>;
    const size_t smem = (PREFETCH ? (size_t)TILE * 16 : 0) + (SPLIT ? (size_t)(TILE + TILE / 16) * 8 : (size_t)(TILE + TILE / 8) * 16);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per = 0, sms = 148;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k, NT, smem);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const long long ntiles = total / TILE;
    const unsigned grid = PREFETCH ? (unsigned)(sms * (per > 0 ? per : 1)) : (unsigned)ntiles;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k<<<grid, NT, smem>>>(d, ntiles);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    printf("%-44s tile %5d  threads %4d  smem %6zu B  CTAs/SM %d  grid %6u : %.3f ms  %5.0f GB/s  (%s %s)\n", tag, TILE, NT, smem, per, grid, best,
           2.0 * total * 16 / best / 1e6, cudaGetErrorString(e), cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const long long total = 1ll << 26;   // 1 GiB
    double2 *d;
    cudaMalloc(&d, total * 16);
    cudaMemset(d, 0, total * 16);
    run<8192, 512, 4, false, false>(d, total, "N=8192 direct (library structure)");
    run<8192, 512, 4, false, true>(d, total, "N=8192 direct, split re/im exchange");
    run<8192, 512, 4, true, true>(d, total, "N=8192 prefetch + split exchange");
    run<8192, 1024, 4, true, true>(d, total, "N=8192 prefetch + split, 8 points/thread");
    run<4096, 256, 3, false, false>(d, total, "N=4096 direct (library structure)");
    run<4096, 256, 3, false, true>(d, total, "N=4096 direct, split re/im exchange");
    run<4096, 256, 3, true, true>(d, total, "N=4096 prefetch + split exchange");
    run<4096, 256, 3, true, false>(d, total, "N=4096 prefetch, full exchange (1 CTA/SM)");
    run<4096, 512, 3, true, true>(d, total, "N=4096 prefetch + split, 8 points/thread");
    run<4096, 512, 3, false, false>(d, total, "N=4096 direct, 8 points/thread");
    double2 h[2];
    cudaMemcpy(h, d + 4242, sizeof(h), cudaMemcpyDeviceToHost);
    printf("check %.0f %.0f (%s)\n", h[0].x, h[1].x, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
