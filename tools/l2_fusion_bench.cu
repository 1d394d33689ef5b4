// l2_fusion_bench.cu -- can two dependent passes over a 1 GiB volume share one HBM round trip through L2?
// Persistent CTAs pull tickets in order: A(c) = contiguous-row RMW tiles of chunk c (z-pass pattern),
// B(c) = strided RMW tiles of chunk c (y-pass pattern, needs ALL A tiles of chunk c).  Ticket order
// A(0) A(1) B(0) A(2) B(1) ... keeps the dependency one chunk behind.  Compare with A-only + B-only.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

struct Sched { unsigned long long ticket; unsigned done[1024]; };

template <int MODE>   // 0 = fused, 1 = A only, 2 = B only
__global__ void __launch_bounds__(256, 3) fused(double2 *data, Sched *S, int planes_per_chunk, int nchunks)
{
    extern __shared__ double2 sm[];
    __shared__ unsigned long long s_ticket;
    const int TILE = 4096, NT = 256, PPT = 16;
    const long long tiles_per_chunk = (long long)planes_per_chunk * 512 * 256 / TILE;
    const long long total = (MODE == 0 ? 2 : 1) * tiles_per_chunk * nchunks;
    for (;;) {
        if (threadIdx.x == 0) s_ticket = atomicAdd(&S->ticket, 1ull);
        __syncthreads();
        const long long t = (long long)s_ticket;
        __syncthreads();
        if (t >= total) break;
        int phase, chunk; long long ti;
        if (MODE == 0) {
            // slots of tiles_per_chunk tickets: slot 0: A0, slot 1: A1, slot 2: B0, slot 3: A2, slot 4: B1, ...
            const long long slot = t / tiles_per_chunk; ti = t % tiles_per_chunk;
            if (slot == 0) { phase = 0; chunk = 0; }
            else if (slot == 2LL * nchunks - 1) { phase = 1; chunk = nchunks - 1; }
            else if (slot & 1) { phase = 0; chunk = (int)((slot + 1) / 2); }
            else { phase = 1; chunk = (int)(slot / 2 - 1); }
        } else { phase = MODE - 1; chunk = (int)(t / tiles_per_chunk); ti = t % tiles_per_chunk; }
        double2 *cb = data + (long long)chunk * planes_per_chunk * 131072;
        double2 v[PPT];
        if (phase == 0) {
            double2 *b = cb + ti * TILE;   // 16 contiguous rows
#pragma unroll
            for (int i = 0; i < PPT; ++i) v[i] = __ldcg(b + threadIdx.x + i * NT);
#pragma unroll
            for (int i = 0; i < PPT; ++i) sm[threadIdx.x + i * NT] = v[i];
            __syncthreads();
#pragma unroll
            for (int i = 0; i < PPT; ++i) { double2 x = sm[(threadIdx.x + i * NT + 17) % TILE]; x.x += 1.0; b[threadIdx.x + i * NT] = x; }
            if (MODE == 0) {
                __threadfence();
                __syncthreads();
                if (threadIdx.x == 0) atomicAdd(&S->done[chunk], 1u);
            }
        } else {
            if (MODE == 0) {
                if (threadIdx.x == 0) {
                    volatile unsigned *d = &S->done[chunk];
                    while (*d < (unsigned)tiles_per_chunk) __nanosleep(100);
                    __threadfence();
                }
                __syncthreads();
            }
            const long long plane = ti / 32, seg = ti % 32;   // 32 column tiles of 8 per plane
            double2 *b = cb + plane * 131072 + seg * 8;
#pragma unroll
            for (int i = 0; i < PPT; ++i) { int idx = threadIdx.x + i * NT; v[i] = __ldcg(b + (long long)(idx >> 3) * 256 + (idx & 7)); }
#pragma unroll
            for (int i = 0; i < PPT; ++i) sm[threadIdx.x + i * NT] = v[i];
            __syncthreads();
#pragma unroll
            for (int i = 0; i < PPT; ++i) { int idx = threadIdx.x + i * NT; double2 x = sm[(idx + 17) % TILE]; x.x += 1.0; b[(long long)(idx >> 3) * 256 + (idx & 7)] = x; }
        }
        __syncthreads();
    }
}

int main()
{
    const size_t total = (size_t)1 << 26;
    double2 *d; Sched *S;
    cudaMalloc(&d, total * 16); cudaMalloc(&S, sizeof(Sched));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFuncSetAttribute(fused<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(fused<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(fused<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int ppc : {4, 8, 16, 32, 64}) {
        const int nchunks = 512 / ppc;
        float ms[3];
        for (int mode = 0; mode < 3; ++mode) {
            for (int rep = 0; rep < 3; ++rep) {
                cudaMemsetAsync(S, 0, sizeof(Sched));
                if (rep == 2 && mode == 0) cudaMemsetAsync(d, 0, total * 16);
                cudaEventRecord(e0);
                if (mode == 0) fused<0><<<148 * 3, 256, 65536>>>(d, S, ppc, nchunks);
                if (mode == 1) fused<1><<<148 * 3, 256, 65536>>>(d, S, ppc, nchunks);
                if (mode == 2) fused<2><<<148 * 3, 256, 65536>>>(d, S, ppc, nchunks);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms[mode], e0, e1);
            }
            if (mode == 0) {   // every element must have been incremented exactly twice
                double2 *h = (double2 *)malloc(1 << 20);
                cudaMemcpy(h, d + 12345678, 1 << 20, cudaMemcpyDeviceToHost);
                double s = 0; for (int i = 0; i < (1 << 16); ++i) s += h[i].x;
                if (s != 2.0 * (1 << 16)) printf("  !! check failed: sum %.1f\n", s);
                free(h);
            }
        }
        printf("chunk %3d planes (%4d MiB): fused %.3f ms (%.0f GB/s algorithmic for 2 passes)   A-only %.3f  B-only %.3f  sum %.3f   (%s)\n", ppc, ppc * 2,
               ms[0], 2.0 * total * 16 / ms[0] / 1e6, ms[1], ms[2], ms[1] + ms[2], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
