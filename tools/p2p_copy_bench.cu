// p2p_copy_bench.cu -- diagnostic for the slab exchange: how should 117 ... 270 MB per direction be
// pushed from one GPU's HBM into a peer's over NVLink while the local FFT passes keep the SMs and the HBM busy?
//   1. SM copy kernel (16-byte loads from local memory, stores into the peer) as a function of the grid size: how few
//      CTAs saturate the link?  (an exchange done by a narrow "pusher" kernel leaves the other SMs to the FFT passes)
//   2. copy engines: cudaMemcpyPeerAsync of the same bytes split over 1 / 2 / 4 / 8 streams
//   3. SM pusher and copy engines at the same time, half the bytes each: do the two paths add up?
//   4. 1 - 3 again while a local HBM-bound kernel (in-place read-modify-write of 1 GiB) runs beside them on its own stream
// Both directions at once (GPU 0 -> 1 and 1 -> 0), as in the all-to-all.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/p2p_copy_bench tools/p2p_copy_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(256) k_push(double2 *__restrict__ dst, const double2 *__restrict__ src, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        const double2 a = __ldcg(src + i), b = __ldcg(src + i + stride), c = __ldcg(src + i + 2 * stride), d = __ldcg(src + i + 3 * stride);
        dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
    }
    for (; i < n; i += stride) dst[i] = __ldcg(src + i);
}

__global__ void __launch_bounds__(256) k_local_rmw(double2 *buf, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double2 v = __ldcg(buf + i);
        v.x += 1.0; v.y -= 1.0;
        buf[i] = v;
    }
}

struct Dev {
    int id;
    double2 *send, *recv, *work;
    cudaStream_t push, load, ce[8];
    cudaEvent_t a, b;
};

int main(int argc, char **argv)
{
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { printf("needs 2 GPUs (have %d)\n", ndev); return 0; }
    const size_t bytes = (size_t)(argc > 1 ? atol(argv[1]) : 256) << 20, n = bytes / 16;
    const size_t work_bytes = (size_t)1 << 30, wn = work_bytes / 16;
    Dev d[2];
    for (int g = 0; g < 2; ++g) {
        d[g].id = g;
        CK(cudaSetDevice(g));
        CK(cudaMalloc(&d[g].send, bytes)); CK(cudaMalloc(&d[g].recv, bytes)); CK(cudaMalloc(&d[g].work, work_bytes));
        CK(cudaMemset(d[g].send, 1, bytes)); CK(cudaMemset(d[g].recv, 0, bytes)); CK(cudaMemset(d[g].work, 0, work_bytes));
        CK(cudaDeviceEnablePeerAccess(1 - g, 0));
        CK(cudaStreamCreateWithFlags(&d[g].push, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&d[g].load, cudaStreamNonBlocking));
        for (int k = 0; k < 8; ++k) CK(cudaStreamCreateWithFlags(&d[g].ce[k], cudaStreamNonBlocking));
        CK(cudaEventCreate(&d[g].a)); CK(cudaEventCreate(&d[g].b));
        cudaDeviceProp p;
        CK(cudaGetDeviceProperties(&p, g));
        if (g == 0) printf("device: %s, %d SMs, asyncEngineCount %d; %zu MiB per direction, both directions at once\n", p.name, p.multiProcessorCount, p.asyncEngineCount, bytes >> 20);
    }
    // one measurement: `issue(g)` enqueues the exchange work of device g; optional local load beside it; returns the
    // slower device's time (events on a gate stream that joins all the streams used)
    auto measure = [&](const char *name, int sm_grid, int ce_streams, double sm_share, bool load) {
        float worst = 0.f, load_ms = 0.f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEvent_t la[2], lb[2];
            for (int g = 0; g < 2; ++g) {
                CK(cudaSetDevice(g));
                CK(cudaDeviceSynchronize());
            }
            for (int g = 0; g < 2; ++g) {
                CK(cudaSetDevice(g));
                Dev &D = d[g], &P = d[1 - g];
                if (load) {
                    CK(cudaEventCreate(&la[g])); CK(cudaEventCreate(&lb[g]));
                    CK(cudaEventRecord(la[g], D.load));
                    for (int r = 0; r < 2; ++r) k_local_rmw<<<148 * 8, 256, 0, D.load>>>(D.work, wn);
                    CK(cudaEventRecord(lb[g], D.load));
                }
                const size_t n_sm = (size_t)((double)n * sm_share) & ~(size_t)255, n_ce = n - n_sm;
                CK(cudaEventRecord(D.a, D.push));
                for (int k = 0; k < ce_streams; ++k) CK(cudaStreamWaitEvent(D.ce[k], D.a, 0));
                if (sm_grid > 0 && n_sm) k_push<<<sm_grid, 256, 0, D.push>>>(P.recv, D.send, n_sm);
                for (int k = 0; k < ce_streams && n_ce; ++k) {
                    const size_t lo = n_sm + n_ce * k / ce_streams, hi = n_sm + n_ce * (k + 1) / ce_streams;
                    CK(cudaMemcpyPeerAsync(P.recv + lo, P.id, D.send + lo, D.id, (hi - lo) * 16, D.ce[k]));
                    cudaEvent_t done;
                    CK(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
                    CK(cudaEventRecord(done, D.ce[k]));
                    CK(cudaStreamWaitEvent(D.push, done, 0));
                    CK(cudaEventDestroy(done));
                }
                CK(cudaEventRecord(D.b, D.push));
            }
            float w = 0.f;
            for (int g = 0; g < 2; ++g) {
                CK(cudaSetDevice(g));
                CK(cudaDeviceSynchronize());
                float ms = 0.f;
                CK(cudaEventElapsedTime(&ms, d[g].a, d[g].b));
                if (ms > w) w = ms;
                if (load) {
                    float l = 0.f;
                    CK(cudaEventElapsedTime(&l, la[g], lb[g]));
                    if (g == 0) load_ms = l;
                    CK(cudaEventDestroy(la[g])); CK(cudaEventDestroy(lb[g]));
                }
            }
            if (rep == 0 || w < worst) worst = w;
        }
        if (load) printf("%-58s %8.3f ms  %7.1f GB/s per direction   (local 2 x 1 GiB RMW beside it: %.3f ms = %.0f GB/s)\n", name, worst, bytes / worst * 1e-6, load_ms, 4.0 * work_bytes / load_ms * 1e-6);
        else printf("%-58s %8.3f ms  %7.1f GB/s per direction\n", name, worst, bytes / worst * 1e-6);
        fflush(stdout);
    };
    char nm[128];
    for (int load = 0; load < 2; ++load) {
        printf("---- %s\n", load ? "with a local HBM-bound kernel running beside the exchange" : "exchange alone");
        for (int grid : {4, 8, 16, 32, 64, 148, 296, 1184}) {
            snprintf(nm, sizeof(nm), "SM pusher, %d CTAs x 256 threads", grid);
            measure(nm, grid, 0, 1.0, load);
        }
        for (int k : {1, 2, 4, 8}) {
            snprintf(nm, sizeof(nm), "copy engines, %d stream(s)", k);
            measure(nm, 0, k, 0.0, load);
        }
        for (int grid : {16, 32, 148}) {
            for (int k : {1, 2}) {
                snprintf(nm, sizeof(nm), "half by a %d-CTA pusher + half by %d copy-engine stream(s)", grid, k);
                measure(nm, grid, k, 0.5, load);
            }
        }
    }
    return 0;
}
