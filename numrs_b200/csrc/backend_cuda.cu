// backend_cuda.cu -- CUDA backend of the planner: kernel dispatch table, the elementwise
// kernel, memory and stream helpers.  There is deliberately no other backend in the product.
#include <dirent.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "kernels_inst.cuh"

namespace nrb {

void register_row_a(PassTable &);
void register_row_b(PassTable &);
void register_row_c(PassTable &);
void register_col_a(PassTable &);
void register_col_b(PassTable &);
void register_col_c(PassTable &);

PassTable &pass_table()
{
    static PassTable t;
    static std::once_flag once;
    std::call_once(once, [] {
        memset(&t, 0, sizeof(t));
        register_row_a(t); register_row_b(t); register_row_c(t);
        register_col_a(t); register_col_b(t); register_col_c(t);
    });
    return t;
}

// big-tile passes (fft_pass2.cuh, k_big.cu): only in builds made with `make BIG=1`; the shipped library has none, and
// big_row_mask / big_col_mask then select nothing
#ifdef NRB_WITH_BIG_TILES
void register_big(PassTable &);
#endif
PassTable &pass2_table()
{
    static PassTable t;
    static std::once_flag once;
    std::call_once(once, [] {
        memset(&t, 0, sizeof(t));
#ifdef NRB_WITH_BIG_TILES
        register_big(t);
#endif
    });
    return t;
}

void register_fused_a();
void register_fused_b();
void register_fused_c();
static std::map<unsigned long long, FusedLaunchFn> &fused_table()
{
    static std::map<unsigned long long, FusedLaunchFn> t;
    return t;
}
void register_fused(unsigned long long key, FusedLaunchFn fn) { fused_table()[key] = fn; }
static void init_fused()
{
    static std::once_flag once;
    std::call_once(once, [] { register_fused_a(); register_fused_b(); register_fused_c(); });
}
static unsigned long long key_of(const KernelKey &a, const KernelKey &b)
{
    return fused_key(a.log2n, a.layout, a.variant, b.log2n, b.layout, b.variant, a.dir);
}

static thread_local std::string g_be_err;
static int fail(cudaError_t e)
{
    g_be_err = std::string(cudaGetErrorName(e)) + ": " + cudaGetErrorString(e);
    return (int)e;
}
const char *be_last_error() { return g_be_err.c_str(); }

// TMA-fed strided pass (fft_tma.cuh, k_tma.cu): taken for the line lengths in the option tma_col_mask when eligible
bool tma_pass_eligible(const KernelKey &key, const PassParams &p, u64 ntiles);
int launch_tma_pass(const KernelKey &key, const PassParams &p, u64 ntiles, int persist, cudaStream_t s);

int be_launch_pass(const KernelKey &key, const PassParams &p, u64 ntiles, void *stream)
{
    if (tma_pass_eligible(key, p, ntiles)) {
        PassParams pt = p;
        pt.simple = pass_is_simple(key, p) ? 1 : 0;
        const int rc = launch_tma_pass(key, pt, ntiles, tunables().tma_persist, (cudaStream_t)stream);
        if (rc != 0) return fail((cudaError_t)rc);
        return 0;
    }
    if (key.log2n < 1 || key.log2n > kMaxLog2N || key.layout < 0 || key.layout > 1 || key.variant < 0 || key.variant > 2) {
        g_be_err = "no such kernel";
        return -1;
    }
    PassLaunchFn fn = pass_table().fn[key.log2n][key.layout][key.dir > 0 ? 1 : 0][key.variant];
    if (use_big_tiles(key, p)) {
        PassLaunchFn big = pass2_table().fn[key.log2n][key.layout][key.dir > 0 ? 1 : 0][key.variant];
        if (big) fn = big;
    }
    if (!fn) { g_be_err = "kernel variant not built"; return -1; }
    if (ntiles > 0x7fffffffull) { g_be_err = "grid too large"; return -1; }
    PassParams pp = p;
    pp.simple = pass_is_simple(key, p) ? 1 : 0;
    const int rc = fn(pp, ntiles, (cudaStream_t)stream);
    if (rc != 0) return fail((cudaError_t)rc);
    return 0;
}

int launch_conv_mid(int log2rest, const ConvMidParams &m, u64 ntiles, cudaStream_t s);   // k_mid.cu
int launch_trig(int log2n, const TrigParams &t, cudaStream_t s);                           // k_trig.cu
int launch_twofft(int log2n, const TwoFFTParams &t, cudaStream_t s);
bool be_trig_available(int log2n) { return log2n >= kTrigMinLog2 && log2n <= kTrigMaxLog2; }
int be_launch_trig(int log2n, const TrigParams &t, void *stream)
{
    if (!be_trig_available(log2n)) { g_be_err = "trig kernel not built for this line length"; return -1; }
    const int rc = launch_trig(log2n, t, (cudaStream_t)stream);
    if (rc != 0) { g_be_err = cudaGetErrorString((cudaError_t)rc); return -1; }
    return 0;
}
int be_launch_twofft(int log2n, const TwoFFTParams &t, void *stream)
{
    if (!be_trig_available(log2n)) { g_be_err = "twofft kernel not built for this line length"; return -1; }
    const int rc = launch_twofft(log2n, t, (cudaStream_t)stream);
    if (rc != 0) { g_be_err = cudaGetErrorString((cudaError_t)rc); return -1; }
    return 0;
}
bool be_conv_mid_available(int log2rest) { return log2rest == 4 || log2rest == 6 || log2rest == 11 || log2rest == 12; }
int be_launch_conv_mid(int log2rest, const ConvMidParams &m, u64 ntiles, void *stream)
{
    if (!be_conv_mid_available(log2rest)) { g_be_err = "conv_mid kernel not built for this row length"; return -1; }
    const int rc = launch_conv_mid(log2rest, m, ntiles, (cudaStream_t)stream);
    if (rc != 0) return fail((cudaError_t)rc);
    return 0;
}

bool be_fused_available(const KernelKey &a, const KernelKey &b)
{
    init_fused();
    return a.dir == b.dir && fused_table().count(key_of(a, b)) != 0;
}

int be_launch_fused(const KernelKey &ka, const PassParams &pa, const KernelKey &kb, const PassParams &pb, const FuseSched &fs,
                    void *stream)
{
    init_fused();
    auto it = fused_table().find(key_of(ka, kb));
    if (it == fused_table().end()) { g_be_err = "fused kernel pair not built"; return -1; }
    const int rc = it->second(pa, pb, fs, (cudaStream_t)stream);
    if (rc != 0) return fail((cudaError_t)rc);
    return 0;
}

__global__ void __launch_bounds__(256) aux_kernel(const __grid_constant__ AuxParams A)
{
    aux_body(A, (u64)blockIdx.x * blockDim.x + threadIdx.x, (u64)gridDim.x * blockDim.x);
}

__global__ void __launch_bounds__(kScanThreads) aux_scan_kernel(const __grid_constant__ AuxParams A)
{
    __shared__ double2 sm[kScanSmem];
    aux_scan_cta(A, sm, blockIdx.x, (int)threadIdx.x);
}

int be_launch_aux(const AuxParams &a, void *stream)
{
    if (a.kind == AUX_SCAN && a.op != 1) {      // block-cooperative phases of the running sum
        const u64 blocks = a.count * ((a.n / 2 + kScanChunk - 1) / kScanChunk);
        if (blocks == 0) return 0;
        if (blocks > 0x7fffffffull) { g_be_err = "grid too large"; return -1; }
        aux_scan_kernel<<<(unsigned)blocks, kScanThreads, 0, (cudaStream_t)stream>>>(a);
        cudaError_t e = cudaGetLastError();
        return e == cudaSuccess ? 0 : fail(e);
    }
    u64 items = 0;
    switch (a.kind) {
    case AUX_UNTANGLE: items = a.count * (a.n >= 2 ? a.n / 2 : 1); break;
    case AUX_SPECTRAL: items = a.count * (a.n / 2); break;
    case AUX_PAD_RESPONSE: items = a.n; break;
    case AUX_FILL: items = a.n; break;
    case AUX_SPECTRAL_Z: items = a.count * (a.n >= 8 ? a.n / 4 : 1); break;
    case AUX_SPECTRAL_ZT: items = a.count * (((1ull << a.m) / 2 + 1) * ((a.n / 2) >> a.m)); break;
    case AUX_SIGNAL: case AUX_WAIT: items = a.count; break;
    case AUX_REDUCE: items = a.count * a.m; break;
    case AUX_STATS_FINAL: items = a.count; break;
    case AUX_NORMALIZE: items = a.count * (a.op == 1 ? a.n / 2 : a.n); break;
    case AUX_PACK2: items = a.count * a.n; break;
    case AUX_POWER: case AUX_SCALE: case AUX_CMUL: items = a.n; break;
    case AUX_TWOFFT_SPLIT: items = a.count * (a.n / 2 + 1); break;
    case AUX_COSFT: items = a.count * a.m; break;
    case AUX_SCAN: items = a.count; break;
    default: items = a.count * a.n; break;
    }
    if (items == 0) return 0;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    u64 blocks = (items + 255) / 256;
    const u64 cap = (u64)sms * 8 * 4;     // grid-stride beyond 4 waves of 8 CTAs/SM
    if (blocks > cap) blocks = cap;
    aux_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail(e);
}

int be_malloc(void **p, size_t bytes)
{
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
    return e == cudaSuccess ? 0 : fail(e);
}
int be_free(void *p)
{
    cudaError_t e = cudaFree(p);
    return e == cudaSuccess ? 0 : fail(e);
}
int be_memset(void *p, int value, size_t bytes, void *stream)
{
    cudaError_t e = cudaMemsetAsync(p, value, bytes, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail(e);
}
int be_ipc_export(void *dptr, unsigned char handle[64])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, dptr);
    if (e != cudaSuccess) return fail(e);
    memcpy(handle, &h, 64);
    return 0;
}
int be_ipc_import(const unsigned char handle[64], void **dptr)
{
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    cudaError_t e = cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess);
    return e == cudaSuccess ? 0 : fail(e);
}
int be_ipc_release(void *dptr)
{
    cudaError_t e = cudaIpcCloseMemHandle(dptr);
    return e == cudaSuccess ? 0 : fail(e);
}
int be_h2d(void *dst, const void *src, size_t bytes, void *stream)
{
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail(e);
}
int be_d2h(void *dst, const void *src, size_t bytes, void *stream)
{
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail(e);
}
int be_d2d(void *dst, const void *src, size_t bytes, void *stream)
{
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail(e);
}
int be_h2d_2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows, void *stream)
{
    cudaError_t e = cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyHostToDevice, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail(e);
}
int be_d2h_2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows, void *stream)
{
    cudaError_t e = cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail(e);
}
int be_enable_peer(int dev, int peer)
{
    cudaError_t e = cudaSetDevice(dev);
    if (e != cudaSuccess) return fail(e);
    if (dev == peer) return 0;
    int can = 0;
    e = cudaDeviceCanAccessPeer(&can, dev, peer);
    if (e != cudaSuccess) return fail(e);
    if (!can) { g_be_err = "devices " + std::to_string(dev) + " and " + std::to_string(peer) + " have no peer access"; return -1; }
    e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
    return e == cudaSuccess ? 0 : fail(e);
}
int be_sync(void *stream)
{
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail(e);
}
void *be_event_record(void *stream)
{
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    cudaEventRecord(e, (cudaStream_t)stream);
    return (void *)e;
}
void *be_event_create()
{
    cudaEvent_t e;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    return (void *)e;
}
int be_event_record_on(void *event, void *stream)
{
    cudaError_t e = cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail(e);
}
int be_stream_wait(void *stream, void *event)
{
    cudaError_t e = cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0);
    return e == cudaSuccess ? 0 : fail(e);
}
int be_stream_create_prio(void **stream, int high_priority)
{
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    cudaStream_t s;
    cudaError_t e = cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, high_priority ? hi : lo);
    if (e != cudaSuccess) return fail(e);
    *stream = (void *)s;
    return 0;
}
float be_event_elapsed_ms(void *a, void *b)
{
    float ms = 0.f;
    cudaEventSynchronize((cudaEvent_t)b);
    cudaEventElapsedTime(&ms, (cudaEvent_t)a, (cudaEvent_t)b);
    return ms;
}
void be_event_destroy(void *e) { cudaEventDestroy((cudaEvent_t)e); }
int be_device_count()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int be_set_device(int dev)
{
    cudaError_t e = cudaSetDevice(dev);
    return e == cudaSuccess ? 0 : fail(e);
}
int be_stream_create(void **stream)
{
    cudaStream_t s;
    cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    if (e != cudaSuccess) return fail(e);
    *stream = (void *)s;
    return 0;
}
int be_stream_destroy(void *stream)
{
    cudaError_t e = cudaStreamDestroy((cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail(e);
}
// Pinned host memory.  On a box with several NUMA nodes a large buffer is interleaved over the nodes (mbind) before it is
// pinned: one host array feeds the PCIe links of ALL the GPUs when a call is spread over them (multi.cpp), and a buffer
// that lives on one socket caps the aggregate at that socket's memory / inter-socket bandwidth (measured on the 8-GPU box:
// 122 GB/s from a single-node buffer against 226 GB/s when every process pinned its own local slab).
static int numa_node_count()
{
    static int n = [] {
        int c = 0;
        DIR *d = opendir("/sys/devices/system/node");
        if (!d) return 1;
        while (struct dirent *e = readdir(d))
            if (!strncmp(e->d_name, "node", 4) && e->d_name[4] >= '0' && e->d_name[4] <= '9') ++c;
        closedir(d);
        return c < 1 ? 1 : c > 62 ? 62 : c;
    }();
    return n;
}
static std::mutex g_host_mu;
static std::map<void *, size_t> &host_maps() { static auto *m = new std::map<void *, size_t>(); return *m; }

void *be_host_alloc(size_t bytes)
{
    const char *env = getenv("NRB_HOST_INTERLEAVE");
    const bool want = !(env && env[0] == '0');
    const int nodes = numa_node_count();
    if (want && nodes > 1 && bytes >= ((size_t)64 << 20)) {
        const size_t len = (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
        void *p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p != MAP_FAILED) {
            unsigned long mask = (1ul << nodes) - 1ul;
            syscall(SYS_mbind, p, len, 3 /* MPOL_INTERLEAVE */, &mask, (unsigned long)(nodes + 1), 0u);   // best effort
            {   // fault the pages in (several threads: the first touch of 8 GiB takes seconds on one)
                const int nt = 8;
                std::vector<std::thread> ts;
                for (int t = 0; t < nt; ++t)
                    ts.emplace_back([=] {
                        const size_t lo = len / nt * t, hi = t == nt - 1 ? len : len / nt * (t + 1);
                        for (size_t o = lo; o < hi; o += 4096) ((volatile char *)p)[o] = 0;
                    });
                for (auto &t : ts) t.join();
            }
            if (cudaHostRegister(p, len, cudaHostRegisterPortable) == cudaSuccess) {
                std::lock_guard<std::mutex> lk(g_host_mu);
                host_maps()[p] = len;
                return p;
            }
            cudaGetLastError();
            munmap(p, len);
        }
    }
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void be_host_free(void *p)
{
    size_t len = 0;
    {
        std::lock_guard<std::mutex> lk(g_host_mu);
        auto it = host_maps().find(p);
        if (it != host_maps().end()) { len = it->second; host_maps().erase(it); }
    }
    if (len) { cudaHostUnregister(p); munmap(p, len); }
    else cudaFreeHost(p);
}
int be_current_device()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    return dev;
}

} // namespace nrb
