// cluster_exchange_bench.cu -- diagnostic for a two-pass rlft3 (round-2 design question): can a thread-block cluster
// hold a tile that is too big for one CTA (one 64 KiB sub-tile per CTA) and do the last radix-CS butterflies across
// CTAs through distributed shared memory at HBM speed?
// Every CTA loads its own 64 KiB tile (512 rows x 128 B, the COL kernel's geometry), cluster.sync(), then computes
// "butterflies" for 512/CS rows: reads element (row, col) of EVERY CTA of the cluster (CS-1 of CS reads are remote
// DSMEM loads), and stores CS outputs, one into each CTA's tile region; cluster.sync() before exit.
// CS = 1 is the plain pass (no exchange).  Patterns as in xy_pattern_bench.cu (rows 4 KiB / 2 MiB apart).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/cluster_exchange_bench tools/cluster_exchange_bench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
namespace cg = cooperative_groups;

template <int CS>
__global__ void __launch_bounds__(512, 2) cluster_pass(double2 *data, long long so, long long sr)
{
    extern __shared__ double2 sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = CS > 1 ? cluster.block_rank() : 0;
    const long long t = blockIdx.x, zg = t & 31, o = t >> 5;       // consecutive CTAs of a cluster: consecutive zg
    const int tid = threadIdx.x, c = tid & 7, r0 = tid >> 3;
    double2 *base = data + o * so + zg * 8;
    double2 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __ldcg(base + (long long)(r0 + 64 * i) * sr + c);
#pragma unroll
    for (int i = 0; i < 8; ++i) sm[(r0 + 64 * i) * 8 + c] = v[i];
    if (CS > 1) cluster.sync(); else __syncthreads();
    // rows [rank*512/CS, (rank+1)*512/CS): 4096/CS butterflies, CS inputs and CS outputs each
    constexpr int ROWS = 512 / CS;
#pragma unroll
    for (int i = 0; i < 8 / CS + (8 % CS ? 1 : 0); ++i) {
        const int row = rank * ROWS + r0 + 64 * i;
        if (r0 + 64 * i >= ROWS) break;
        double2 in[CS];
#pragma unroll
        for (int p = 0; p < CS; ++p) {
            const double2 *src = CS > 1 ? cluster.map_shared_rank(sm, p) : sm;
            in[p] = src[row * 8 + c];
        }
#pragma unroll
        for (int q = 0; q < CS; ++q) {                  // output q of the butterfly -> tile of CTA q
            double2 acc = make_double2(0.0, 0.0);
#pragma unroll
            for (int p = 0; p < CS; ++p) { acc.x += in[p].x * (double)(q + 1) - in[p].y; acc.y += in[p].y + in[p].x * (double)(p + q); }
            double2 *ob = data + o * so + ((zg & ~(long long)(CS - 1)) + q) * 8;
            ob[(long long)row * sr + c] = acc;
        }
    }
    if (CS > 1) cluster.sync();
}

template <int CS> float run(double2 *d, long long so, long long sr)
{
    cudaFuncSetAttribute(cluster_pass<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    if (CS > 8) cudaFuncSetAttribute(cluster_pass<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(512 * 32);
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = 65536;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        cudaLaunchKernelEx(&cfg, cluster_pass<CS>, d, so, sr);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 2 && ms < best) best = ms;
    }
    return best;
}

int main()
{
    const size_t total = (size_t)1 << 26;
    double2 *d;
    cudaMalloc(&d, total * 16);
    cudaMemset(d, 0, total * 16);
    const long long P = 512LL * 256, R = 256;
    for (int pat = 0; pat < 2; ++pat) {
        const long long so = pat == 0 ? P : R, sr = pat == 0 ? R : P;
        const char *nm = pat == 0 ? "y pattern (rows 4 KiB apart)" : "x pattern (rows 2 MiB apart)";
        float ms;
        ms = run<1>(d, so, sr);  printf("%s  cluster 1 (no exchange): %.3f ms  %.0f GB/s  (%s)\n", nm, ms, 2.0 * total * 16 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        ms = run<2>(d, so, sr);  printf("%s  cluster 2             : %.3f ms  %.0f GB/s  (%s)\n", nm, ms, 2.0 * total * 16 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        ms = run<4>(d, so, sr);  printf("%s  cluster 4             : %.3f ms  %.0f GB/s  (%s)\n", nm, ms, 2.0 * total * 16 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        ms = run<8>(d, so, sr);  printf("%s  cluster 8             : %.3f ms  %.0f GB/s  (%s)\n", nm, ms, 2.0 * total * 16 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        ms = run<16>(d, so, sr); printf("%s  cluster 16 (opt-in)   : %.3f ms  %.0f GB/s  (%s)\n", nm, ms, 2.0 * total * 16 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
