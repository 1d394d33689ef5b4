// pipelined_copy_bench.cu -- does prefetching the next tile (cp.async into a second/third smem buffer,
// persistent CTAs) raise the ceiling of the strided 128-byte-segment pattern over "3 CTAs/SM, load all /
// sync / store all"?  Pattern: volume [512][512][256] complex f64, tile = 512 rows x 8 columns (x- or y-axis).
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_pipeline.h>

__device__ __forceinline__ void cp16(double2 *smem, const double2 *g)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(g));
}

template <int NBUF, int NT>
__global__ void __launch_bounds__(NT) pipelined(double2 *data, long long row_stride, long long tiles_per_block, long long row_block_stride,
                                                 long long ntiles)
{
    extern __shared__ double2 sm[];
    constexpr int TILE = 4096, SEG = 8, PPT = TILE / NT;
    auto tile_base = [&](long long t) { return data + (t / tiles_per_block) * row_block_stride + (t % tiles_per_block) * SEG; };
    auto issue = [&](long long t, int buf) {
        if (t < ntiles) {
            const double2 *b = tile_base(t);
#pragma unroll
            for (int i = 0; i < PPT; ++i) {
                int idx = threadIdx.x + i * NT;
                cp16(sm + buf * TILE + idx, b + (long long)(idx / SEG) * row_stride + (idx % SEG));
            }
        }
        asm volatile("cp.async.commit_group;\n");
    };
    long long t = blockIdx.x;
    for (int k = 0; k < NBUF - 1; ++k) issue(t + (long long)k * gridDim.x, k);
    int buf = 0;
    for (; t < ntiles; t += gridDim.x) {
        issue(t + (long long)(NBUF - 1) * gridDim.x, (buf + NBUF - 1) % NBUF);
        asm volatile("cp.async.wait_group %0;\n" ::"n"(NBUF - 1));
        __syncthreads();
        double2 *b = tile_base(t);
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            int idx = threadIdx.x + i * NT;
            double2 x = sm[buf * TILE + (idx + 17) % TILE];
            x.x += 1.0;
            b[(long long)(idx / SEG) * row_stride + (idx % SEG)] = x;
        }
        __syncthreads();
        buf = (buf + 1) % NBUF;
    }
}

template <int NBUF, int NT> void run(double2 *d, int pass, int ctas_per_sm, const char *tag)
{
    const long long row_stride = pass == 0 ? 131072 : 256;
    const long long tiles_per_block = (pass == 0 ? 131072 : 256) / 8;
    const long long nblocks = pass == 0 ? 1 : 512;
    const long long ntiles = tiles_per_block * nblocks;
    const size_t smem = (size_t)NBUF * 4096 * 16;
    cudaFuncSetAttribute(pipelined<NBUF, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        pipelined<NBUF, NT><<<148 * ctas_per_sm, NT, smem>>>(d, row_stride, tiles_per_block, 131072, ntiles);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("%s-pass  %-28s nbuf %d threads %4d ctas/sm %d : %.3f ms  %.0f GB/s (%s)\n", pass == 0 ? "x" : "y", tag, NBUF, NT, ctas_per_sm,
           ms, 2.0 * (1 << 26) * 16 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    double2 *d;
    cudaMalloc(&d, (size_t)(1 << 26) * 16);
    cudaMemset(d, 0, (size_t)(1 << 26) * 16);
    for (int pass = 0; pass < 2; ++pass) {
        run<1, 256>(d, pass, 3, "persistent, no prefetch");
        run<2, 256>(d, pass, 1, "double buffer");
        run<3, 256>(d, pass, 1, "triple buffer");
        run<3, 512>(d, pass, 1, "triple buffer");
        run<2, 512>(d, pass, 1, "double buffer");
        run<1, 128>(d, pass, 3, "persistent, no prefetch");
    }
    return 0;
}
