// p2p_store_bench.cu -- diagnostic: how fast can SMs of GPU 0 write into GPU 1's memory over NVLink, as a
// function of the store pattern?  Decides the layout of the slab-exchange receive buffers.
//   coalesced : every warp stores 512 contiguous bytes, warps walk the buffer linearly
//   run128    : the x-pass pattern of the fused exchange: 8 lanes store one 128-byte run, the 4 runs of a warp
//               instruction are `stride` bytes apart (receive layout [X][Y][N3], stride = Y*N3*16)
//   run512    : receive layout [Y][N3/8][X][8]: the 4 runs of a warp instruction are adjacent (512 B contiguous)
//   bulk      : cp.async.bulk shared -> peer global, one chunk of `chunk` bytes per elected thread
//   memcpy    : cudaMemcpyPeerAsync
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/p2p_store_bench tools/p2p_store_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void k_coalesced(double2 *dst, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = make_double2((double)i, 1.0);
}

// element index space: tiles of 4096 elements (512 "x" values times 8 "z" values); thread t of a 512-thread CTA
// stores 8 elements: z = t & 7, x = (t >> 3) + 64 * r.  Address of (tile, x, z):
//   run128: x * xstride + tile * 8 + z             (xstride in elements, >= tiles * 8)
//   run512: tile * 4096 + x * 8 + z  -> the 4 x values of a warp instruction are adjacent
__global__ void __launch_bounds__(512, 2) k_runs(double2 *dst, size_t tiles, size_t xstride, int mode)
{
    const int t = threadIdx.x, z = t & 7, x0 = t >> 3;
    for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int x = x0 + 64 * r;
            const size_t a = mode == 0 ? (size_t)x * xstride + tile * 8 + z : tile * 4096 + (size_t)x * 8 + z;
            dst[a] = make_double2((double)t, (double)r);
        }
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// every CTA owns 64 KiB of shared memory = one tile; it is sent as 65536 / chunk bulk copies
__global__ void __launch_bounds__(512, 2) k_bulk(double2 *dst, size_t tiles, int chunk)
{
    extern __shared__ __align__(128) unsigned char sm[];
    double2 *s = reinterpret_cast<double2 *>(sm);
    const int t = threadIdx.x;
    for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
#pragma unroll
        for (int r = 0; r < 8; ++r) s[t + 512 * r] = make_double2((double)t, (double)r);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        const int nchunks = 65536 / chunk;
        for (int c = t; c < nchunks; c += 512) {
            char *g = reinterpret_cast<char *>(dst) + tile * 65536 + (size_t)c * chunk;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(smem_u32(sm + (size_t)c * chunk)), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem may be overwritten
        __syncthreads();
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <class F> static float time_ms(F f, int reps, cudaStream_t s)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); f();
    CK(cudaStreamSynchronize(s));
    CK(cudaEventRecord(a, s));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(b, s));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}

int main(int argc, char **argv)
{
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { printf("needs 2 GPUs (have %d)\n", ndev); return 0; }
    const size_t bytes = (size_t)256 << 20, n = bytes / 16, tiles = n / 4096;
    double2 *remote = nullptr, *local = nullptr;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&remote, bytes)); CK(cudaMemset(remote, 0, bytes));
    CK(cudaSetDevice(0)); CK(cudaMalloc(&local, bytes)); CK(cudaMemset(local, 0, bytes));
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, 0, 1));
    if (!can) { printf("no peer access 0 -> 1\n"); return 0; }
    CK(cudaDeviceEnablePeerAccess(1, 0));
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    const int reps = 10;
    auto report = [&](const char *name, float ms) { printf("%-34s %8.3f ms  %7.1f GB/s\n", name, ms, bytes / ms * 1e-6); fflush(stdout); };
    for (int tgt = 0; tgt < 2; ++tgt) {
        double2 *dst = tgt == 0 ? local : remote;
        printf("---- destination: %s\n", tgt == 0 ? "local HBM (reference)" : "peer GPU over NVLink");
        report("coalesced 16B/thread", time_ms([&] { k_coalesced<<<sms * 8, 512, 0, s>>>(dst, n); }, reps, s));
        for (int grid_mul = 2; grid_mul <= 2; ++grid_mul) {
            report("run128, x stride 256 KiB", time_ms([&] { k_runs<<<sms * grid_mul, 512, 0, s>>>(dst, tiles, tiles * 8, 0); }, reps, s));
            report("run512 (4 adjacent runs / instr)", time_ms([&] { k_runs<<<sms * grid_mul, 512, 0, s>>>(dst, tiles, 0, 1); }, reps, s));
        }
        for (int chunk = 128; chunk <= 65536; chunk *= 4) {
            char nm[64];
            snprintf(nm, sizeof(nm), "bulk smem->global, chunk %d B", chunk);
            report(nm, time_ms([&] { k_bulk<<<sms * 2, 512, 65536, s>>>(dst, tiles, chunk); }, reps, s));
        }
    }
    report("cudaMemcpyPeerAsync 0 -> 1", time_ms([&] { cudaMemcpyPeerAsync(remote, 1, local, 0, bytes, s); }, reps, s));
    CK(cudaGetLastError());
    return 0;
}
