// strided_copy_bench.cu -- upper bound for the strided-tile access pattern of the COL passes:
// each CTA reads a tile of ROWS segments of SEG bytes (row stride STRIDE bytes), then writes it back
// in place.  Sweeps segment size / resident CTAs to find what HBM sustains for this pattern.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int PPT>
__global__ void tile_copy(double2 *data, long long row_stride /*elems*/, int seg /*elems*/, int rows, long long tiles_per_row_block,
                          long long row_block_stride)
{
    // tile t: segment index s = t % segs_per_row, block = t / segs_per_row
    extern __shared__ double2 sm[];
    const long long t = blockIdx.x;
    const long long s = t % tiles_per_row_block, blk = t / tiles_per_row_block;
    double2 *base = data + blk * row_block_stride + s * seg;
    double2 v[PPT];
    const int nt = blockDim.x;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        int idx = threadIdx.x + i * nt;
        int r = idx / seg, c = idx % seg;
        v[i] = base[(long long)r * row_stride + c];
    }
#pragma unroll
    for (int i = 0; i < PPT; ++i) sm[threadIdx.x + i * nt] = v[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        int idx = threadIdx.x + i * nt;
        int r = idx / seg, c = idx % seg;
        double2 x = sm[(threadIdx.x + i * nt + 17) % (nt * PPT)];
        x.x += 1.0;
        base[(long long)r * row_stride + c] = x;
        (void)r; (void)c;
    }
}

int main()
{
    const size_t total = (size_t)1 << 26;   // complex elements = 1 GiB
    double2 *d;
    cudaMalloc(&d, total * sizeof(double2));
    cudaMemset(d, 0, total * sizeof(double2));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int PPT = 16;
    // volume view [512][512][256] complex: x-pass: rows = 512 (stride 131072), y-pass: stride 256
    for (int pass = 0; pass < 2; ++pass) {
        for (int segb = 64; segb <= 1024; segb *= 2) {
            const int seg = segb / 16;
            const int rows = 512;
            const int tile = rows * seg;
            const int nt = tile / PPT;
            if (nt > 1024 || nt < 32) continue;
            const long long row_stride = pass == 0 ? 131072 : 256;
            const long long tiles_per_block = (pass == 0 ? 131072 : 256) / seg;
            const long long nblocks = pass == 0 ? 1 : 512;
            const long long row_block_stride = 131072;
            const long long ntiles = tiles_per_block * nblocks;
            const size_t smem = (size_t)tile * 16;
            cudaFuncSetAttribute(tile_copy<PPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                tile_copy<PPT><<<(unsigned)ntiles, nt, smem>>>(d, row_stride, seg, rows, tiles_per_block, row_block_stride);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
            }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("%s-pass pattern  seg %4d B  tile %3zu KiB  threads %4d  : %.3f ms  %.0f GB/s  (%s)\n", pass == 0 ? "x" : "y", segb,
                   smem >> 10, nt, ms, 2.0 * total * 16 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    }
    // plain streaming copy in place for reference
    return 0;
}
