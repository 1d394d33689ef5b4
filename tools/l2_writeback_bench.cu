// l2_writeback_bench.cu -- can an intermediate result that is overwritten in place by the next pass stay in L2
// and never reach DRAM?  (Decides whether rlft3's z + y passes can share ONE HBM round trip: tools/l2_fusion_bench.cu
// found that pairing two passes through L2 only saves the re-read, i.e. the dirty lines of pass A are written back
// before pass B overwrites them.  This bench repeats the experiment with the L2 eviction-priority controls that
// were not tried: evict_last store hints, evict_first loads, and a persisting access-policy window.)
//
// One persistent launch walks a 1 GiB buffer chunk by chunk: pass A = in-place read-modify-write of the chunk,
// grid barrier, pass B = in-place read-modify-write of the same chunk by different CTAs, next chunk.
// DRAM traffic is 4 units (2 reads + 2 writes of the buffer) if A's output is written back and re-read,
// 3 units if only the re-read is saved, 2 units if the write-back is elided too.  Time tells which.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned ld_acq(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void grid_bar(unsigned *ctr, unsigned &target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(ctr, 1u);
        while (ld_acq(ctr) < target) __nanosleep(32);
        __threadfence();
    }
    __syncthreads();
}
__device__ __forceinline__ unsigned long long policy_last()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long policy_first()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void st_hint(double2 *p, double2 v, unsigned long long pol)
{
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ double2 ld_hint(const double2 *p, unsigned long long pol)
{
    double2 v;
    asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol) : "memory");
    return v;
}

// HINT 0: plain; 1: pass-A stores evict_last; 2: + pass-B loads evict_first; 3: pass-A stores AND loads of pass B evict_last
template <int HINT>
__global__ void __launch_bounds__(256) sweep(double2 *data, long long chunk_elems, int nchunks, int passes, unsigned *ctr)
{
    unsigned target = 0;
    const unsigned long long pl = policy_last(), pf = policy_first();
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    for (int c = 0; c < nchunks; ++c) {
        double2 *b = data + (long long)c * chunk_elems;
        for (int ps = 0; ps < passes; ++ps) {
            // pass B is done by "the other end" of the grid so the data really crosses SMs / L2 slices
            const long long me = ps == 0 ? (long long)blockIdx.x * blockDim.x + threadIdx.x
                                         : (long long)(gridDim.x - 1 - blockIdx.x) * blockDim.x + threadIdx.x;
            for (long long i = me; i < chunk_elems; i += 4 * nthreads) {
                double2 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const long long k = i + u * nthreads;
                    if (k < chunk_elems) {
                        if (ps == 1 && HINT == 2) v[u] = ld_hint(b + k, pf);
                        else if (ps == 1 && HINT == 3) v[u] = ld_hint(b + k, pl);
                        else v[u] = __ldcg(b + k);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const long long k = i + u * nthreads;
                    if (k < chunk_elems) {
                        v[u].x += 1.0;
                        if (ps == 0 && passes == 2 && HINT >= 1) st_hint(b + k, v[u], pl);
                        else b[k] = v[u];
                    }
                }
            }
            if (passes == 2) grid_bar(ctr, target);
        }
    }
}

template <int HINT> float run(double2 *d, unsigned *ctr, long long total, long long chunk, int passes, int grid, cudaStream_t s)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaMemsetAsync(ctr, 0, 4, s);
        cudaEventRecord(e0, s);
        sweep<HINT><<<grid, 256, 0, s>>>(d, chunk, (int)(total / chunk), passes, ctr);
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    return best;
}

int main()
{
    const long long total = 1ll << 26;   // 1 GiB of double2
    double2 *d; unsigned *ctr;
    cudaMalloc(&d, total * 16); cudaMalloc(&ctr, 4);
    cudaMemset(d, 0, total * 16);
    cudaStream_t s; cudaStreamCreate(&s);
    int sms = 148, per = 1, dev = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, sweep<0>, 256, 0);
    if (per > 4) per = 4;
    const int grid = sms * per;
    const float one = run<0>(d, ctr, total, total, 1, grid, s);
    printf("grid %d CTAs; one in-place sweep of 1 GiB: %.3f ms (%.0f GB/s); two sweeps = 4 traffic units = %.3f ms\n", grid, one,
           2.0 * total * 16 / one / 1e6, 2 * one);
    for (int window = 0; window < 2; ++window) {
        if (window) {
            // persisting access-policy window over the whole buffer (hits of the window stay in the set-aside L2)
            cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, p.persistingL2CacheMaxSize);
            cudaStreamAttrValue a = {};
            a.accessPolicyWindow.base_ptr = d;
            a.accessPolicyWindow.num_bytes = (size_t)p.accessPolicyMaxWindowSize < (size_t)total * 16 ? (size_t)p.accessPolicyMaxWindowSize : (size_t)total * 16;
            a.accessPolicyWindow.hitRatio = 1.0f;
            a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            a.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cudaError_t e = cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &a);
            printf("-- persisting window: set-aside %d MiB, window %zu MiB (%s)\n", p.persistingL2CacheMaxSize >> 20, a.accessPolicyWindow.num_bytes >> 20,
                   cudaGetErrorString(e));
        }
        for (long long mib : {4ll, 8ll, 16ll, 32ll, 64ll}) {
            const long long chunk = mib << 16;   // elements
            const float t0 = run<0>(d, ctr, total, chunk, 2, grid, s), t1 = run<1>(d, ctr, total, chunk, 2, grid, s),
                        t2 = run<2>(d, ctr, total, chunk, 2, grid, s), t3 = run<3>(d, ctr, total, chunk, 2, grid, s);
            printf("chunk %3lld MiB: A+B per chunk  plain %.3f  A-st evict_last %.3f  +B-ld evict_first %.3f  +B-ld evict_last %.3f ms   (units of one sweep: %.2f %.2f %.2f %.2f; 2.0 = nothing saved, 1.5 = re-read saved, 1.0 = write-back elided too)\n",
                   mib, t0, t1, t2, t3, t0 / one, t1 / one, t2 / one, t3 / one);
        }
    }
    // every element must have been incremented the same number of times
    double2 h[4];
    cudaMemcpy(h, d + 12345, sizeof(h), cudaMemcpyDeviceToHost);
    printf("check: %.0f %.0f %.0f %.0f (%s)\n", h[0].x, h[1].x, h[2].x, h[3].x, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
