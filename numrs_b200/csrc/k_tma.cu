// k_tma.cu -- host side of the TMA-fed strided pass (fft_tma.cuh): tensor-map construction and launch.
#include <cuda.h>
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <string>
#include <tuple>

#include "fft_tma.cuh"
#include "plan.h"

namespace nrb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// can this launch take the TMA-fed kernel?  (geometry: plan.h pass_takes_tma; here the pointer alignment and the encoder)
bool tma_pass_eligible(const KernelKey &key, const PassParams &p, u64 ntiles)
{
    if (!pass_takes_tma(key, p) || ntiles == 0) return false;
    if (p.out_peer_on && !pass_takes_tma_in(key, p)) return false;
    if ((((size_t)p.in) | ((size_t)p.out)) & 15) return false;
    return encode_fn() != nullptr;
}

static int make_map(CUtensorMap *m, const double2 *base, u64 inner, u64 N, u64 outer, int log2n, bool swizzle64 = false)
{
    const u64 L = (u64)lines_per_tile(log2n, LAYOUT_COL);
    const cuuint64_t gdim[3] = {2 * inner, N, outer};
    const cuuint64_t gstr[2] = {inner * 16, N * inner * 16};
    const cuuint32_t box[3] = {(cuuint32_t)(2 * L), (cuuint32_t)(N < 256 ? N : 256), 1};
    const cuuint32_t est[3] = {1, 1, 1};
    const CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)base, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

template <int LOG2N, int DIR, bool PERSIST>
static int launch_tma_t(const PassParams &p, u64 ntiles, cudaStream_t s)
{
    typedef Geo<LOG2N, LAYOUT_COL, VAR_PLAIN> G;
    constexpr size_t smem = (size_t)G::TILE * 16 * (PERSIST ? 2 : 1);
    auto kern = fft_col_tma_kernel<LOG2N, DIR, PERSIST>;
    static int resident[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!resident[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        int per_sm = 0, sms = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, G::NT, smem);
        if (e != cudaSuccess) return (int)e;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        resident[dev & 63] = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 148);
    }
    const u64 inner = 1ull << p.logB, N = 1ull << LOG2N;
    const u64 outer = (p.q_end + inner - 1) >> p.logB;          // the map covers outer indices [0, outer)
    CUtensorMap tin, tout;
    int rc = make_map(&tin, p.in, inner, N, outer, LOG2N);
    if (rc == 0) rc = make_map(&tout, p.out, inner, N, outer, LOG2N);
    if (rc != 0) return rc;
    u64 grid = ntiles;
    if (PERSIST && grid > (u64)resident[dev & 63]) grid = (u64)resident[dev & 63];
    kern<<<(unsigned)grid, G::NT, smem, s>>>(tin, tout, p, (unsigned)ntiles, (unsigned)p.logB);
    return (int)cudaGetLastError();
}

template <int LOG2N, int DIR, int CTAS = 2> static int launch_tma_in_t(const PassParams &p, u64 ntiles, int log2_inner, cudaStream_t s)
{
    typedef Geo<LOG2N, LAYOUT_COL, VAR_PLAIN> G;
    constexpr size_t smem = (size_t)G::TILE * 16;
    auto kern = fft_col_tma_in_kernel<LOG2N, DIR, CTAS>;
    static bool ready[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!ready[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        ready[dev & 63] = true;
    }
    const u64 inner = 1ull << log2_inner;
    const u64 outer = (p.q_end + inner - 1) >> log2_inner;
    CUtensorMap tin;
    const int rc = make_map(&tin, p.in, inner, 1ull << LOG2N, outer, LOG2N);
    if (rc != 0) return rc;
    kern<<<(unsigned)ntiles, G::NT, smem, s>>>(tin, p, (unsigned)ntiles, (unsigned)log2_inner);
    return (int)cudaGetLastError();
}

template <int DIR> static int launch_xpose_tma_t(const PassParams &p, u64 ntiles, cudaStream_t s)
{
    typedef Geo<10, LAYOUT_COL, VAR_XPOSE> G;
    constexpr size_t smem = (size_t)G::TILE * 16;
    auto kern = fft_xpose_tma_kernel<DIR>;
    static bool ready[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!ready[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        ready[dev & 63] = true;
    }
    const u64 rest = 1ull << p.logA;
    const u64 outer = (p.q_end + rest - 1) >> p.logA;
    CUtensorMap tin;
    const int rc = make_map(&tin, p.in, rest, 1024, outer, 10, true);
    if (rc != 0) return rc;
    kern<<<(unsigned)ntiles, G::NT, smem, s>>>(tin, p, (unsigned)ntiles, (unsigned)p.logA);
    return (int)cudaGetLastError();
}

int launch_tma_pass(const KernelKey &key, const PassParams &p, u64 ntiles, int persist, cudaStream_t s)
{
    if (key.variant == VAR_XPOSE) return key.dir > 0 ? launch_xpose_tma_t<+1>(p, ntiles, s) : launch_xpose_tma_t<-1>(p, ntiles, s);
    int li = 0;
    if (pass_takes_tma_in(key, p, &li)) {
        switch (key.log2n) {
        case 7: return key.dir > 0 ? launch_tma_in_t<7, +1>(p, ntiles, li, s) : launch_tma_in_t<7, -1>(p, ntiles, li, s);
        case 8: return key.dir > 0 ? launch_tma_in_t<8, +1>(p, ntiles, li, s) : launch_tma_in_t<8, -1>(p, ntiles, li, s);
        case 9:
            if (tunables().tma_in_ctas >= 3) return key.dir > 0 ? launch_tma_in_t<9, +1, 3>(p, ntiles, li, s) : launch_tma_in_t<9, -1, 3>(p, ntiles, li, s);
            return key.dir > 0 ? launch_tma_in_t<9, +1>(p, ntiles, li, s) : launch_tma_in_t<9, -1>(p, ntiles, li, s);
        case 10: return key.dir > 0 ? launch_tma_in_t<10, +1>(p, ntiles, li, s) : launch_tma_in_t<10, -1>(p, ntiles, li, s);
        default: return (int)cudaErrorInvalidValue;
        }
    }
#define NRB_TMA_CASE(LG) \
    case LG: \
        if (key.dir > 0) return persist ? launch_tma_t<LG, +1, true>(p, ntiles, s) : launch_tma_t<LG, +1, false>(p, ntiles, s); \
        return persist ? launch_tma_t<LG, -1, true>(p, ntiles, s) : launch_tma_t<LG, -1, false>(p, ntiles, s);
    switch (key.log2n) {
        NRB_TMA_CASE(7)
        NRB_TMA_CASE(8)
        NRB_TMA_CASE(9)
        NRB_TMA_CASE(10)
    default: return (int)cudaErrorInvalidValue;
    }
#undef NRB_TMA_CASE
}

} // namespace nrb
