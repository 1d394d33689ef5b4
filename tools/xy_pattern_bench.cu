// xy_pattern_bench.cu -- diagnostic for the two strided passes of rlft3 512^3 (volume [512][512][256] complex):
// 64 KiB tiles of 512 rows x 128 bytes, 512 threads x 8 elements, 2 CTAs/SM (the COL kernel's geometry).
// "y pattern": rows 4 KiB apart (a tile spans 2 MiB);  "x pattern": rows 2 MiB apart (a tile touches 512 pages).
// Question: is the x pattern's cost on the load side (TLB / long scoreboard) or on both sides -- i.e. would a
// transposed intermediate (every pass loads with the y pattern and stores with the x pattern) be faster?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/xy_pattern_bench tools/xy_pattern_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>

__global__ void __launch_bounds__(512, 2) tile_pass(const double2 *__restrict__ in, double2 *__restrict__ out, long long ro, long long rr,
                                                    long long wo, long long wr)
{
    extern __shared__ double2 sm[];
    const long long t = blockIdx.x, zg = t & 31, o = t >> 5;
    const int tid = threadIdx.x, c = tid & 7, r0 = tid >> 3;
    double2 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __ldcg(in + o * ro + (long long)(r0 + 64 * i) * rr + zg * 8 + c);
#pragma unroll
    for (int i = 0; i < 8; ++i) sm[tid + 512 * i] = v[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        double2 x = sm[(tid + 512 * i + 8 * 37) & 4095];     // another row, same column
        x.x += 1.0;
        out[o * wo + (long long)(((r0 + 64 * i) + 37) & 511) * wr + zg * 8 + c] = x;
    }
}

int main()
{
    const size_t total = (size_t)1 << 26;
    double2 *a, *b;
    cudaMalloc(&a, total * 16); cudaMalloc(&b, total * 16);
    cudaMemset(a, 0, total * 16); cudaMemset(b, 0, total * 16);
    cudaFuncSetAttribute(tile_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const long long P = 512LL * 256, R = 256;     // plane and row strides in elements
    struct Case { const char *name; int inplace; long long ro, rr, wo, wr; } cases[] = {
        {"y pattern in place            (load 4 KiB rows, store 4 KiB rows)", 1, P, R, P, R},
        {"x pattern in place            (load 2 MiB rows, store 2 MiB rows)", 1, R, P, R, P},
        {"y pattern out of place        (load 4 KiB rows, store 4 KiB rows)", 0, P, R, P, R},
        {"x pattern out of place        (load 2 MiB rows, store 2 MiB rows)", 0, R, P, R, P},
        {"transposing: load y, store x  (load 4 KiB rows, store 2 MiB rows)", 0, P, R, R, P},
        {"transposing: load x, store y  (load 2 MiB rows, store 4 KiB rows)", 0, R, P, P, R},
    };
    for (const Case &cs : cases) {
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            tile_pass<<<512 * 32, 512, 65536>>>(a, cs.inplace ? a : b, cs.ro, cs.rr, cs.wo, cs.wr);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep >= 2 && ms < best) best = ms;
        }
        printf("%s : %.3f ms  %.0f GB/s  (%s)\n", cs.name, best, 2.0 * total * 16 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
