// emu_backend.cpp -- TEST-ONLY host emulation of the numrs_b200 kernels.
//
// Compiles the very same device source (fft_pass.cuh, aux_kernels.cuh) with -DNRB_EMU and
// runs every CTA as a set of cooperative fibers (ucontext) whose __syncthreads() is a
// round-robin yield.  It exists so the CPU-only test tier can validate the planner and the
// index arithmetic of every kernel variant in a container without a GPU.  It is built into
// tests/emu/libnrb_emu.so only; the product library has no such path and the numrs_b200
// Python package never loads this file's output.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <ucontext.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../numrs_b200/csrc/aux_kernels.cuh"
#include "../../numrs_b200/csrc/fft_pass2.cuh"
#include "../../numrs_b200/csrc/conv_mid.cuh"
#include "../../numrs_b200/csrc/trig_fused.cuh"
#include "../../numrs_b200/csrc/plan.h"

namespace nrb_emu {

struct Cta;
void barrier();
struct Cta {
    ucontext_t main_ctx;
    std::vector<ucontext_t> fibers;
    std::vector<char> stacks;
    std::vector<char> done;
    std::vector<double2> smem;
    int current;
    void (*body)(const void *, double2 *, unsigned, int);
    const void *params;
    unsigned tile;
};
static thread_local Cta *t_cta = nullptr;

// __shfl_sync for the fibers: publish, CTA-wide barrier, read the source lane of the own warp, barrier
static thread_local std::vector<double> *t_shfl = nullptr;
double shfl(double v, int src_lane)
{
    Cta *c = t_cta;
    const int tid = c->current;
    (*t_shfl)[tid] = v;
    barrier();
    const double r = (*t_shfl)[(tid & ~31) | (src_lane & 31)];
    barrier();
    return r;
}

void barrier()
{
    Cta *c = t_cta;
    swapcontext(&c->fibers[c->current], &c->main_ctx);
}

static void fiber_entry()
{
    Cta *c = t_cta;
    const int tid = c->current;
    c->body(c->params, c->smem.data(), c->tile, tid);
    c->done[tid] = 1;
    swapcontext(&c->fibers[tid], &c->main_ctx);
}

static const size_t kStack = 64 * 1024;

static void run_cta(Cta &c, int nthreads)
{
    t_cta = &c;
    for (int t = 0; t < nthreads; ++t) {
        getcontext(&c.fibers[t]);
        c.fibers[t].uc_stack.ss_sp = c.stacks.data() + (size_t)t * kStack;
        c.fibers[t].uc_stack.ss_size = kStack;
        c.fibers[t].uc_link = nullptr;
        makecontext(&c.fibers[t], fiber_entry, 0);
        c.done[t] = 0;
    }
    int remaining = nthreads;
    while (remaining > 0) {
        remaining = 0;
        for (int t = 0; t < nthreads; ++t) {
            if (c.done[t]) continue;
            c.current = t;
            swapcontext(&c.main_ctx, &c.fibers[t]);
            if (!c.done[t]) ++remaining;
        }
    }
}

template <int LOG2N, int LAYOUT, int DIR, int VARIANT>
static void body_tpl(const void *p, double2 *sm, unsigned tile, int tid)
{
    const nrb::PassParams &P = *(const nrb::PassParams *)p;
    if constexpr (nrb::simple_built(LOG2N, LAYOUT, VARIANT)) {
        if (P.simple) { nrb::fft_pass_body<LOG2N, LAYOUT, DIR, VARIANT, true>(P, sm, tile, tid); return; }
    }
    nrb::fft_pass_body<LOG2N, LAYOUT, DIR, VARIANT, false>(P, sm, tile, tid);
}

typedef void (*BodyFn)(const void *, double2 *, unsigned, int);
struct Entry { BodyFn fn; int nthreads; size_t smem; };
static Entry g_table[nrb::kMaxLog2N + 1][2][2][3];

template <int LOG2N, int LAYOUT> static void reg()
{
    using namespace nrb;
    const int nt = cta_threads(LOG2N, LAYOUT);
    if (LAYOUT == LAYOUT_ROW) {
        g_table[LOG2N][0][1][VAR_PLAIN] = Entry{body_tpl<LOG2N, LAYOUT_ROW, 1, VAR_PLAIN>, nt, smem_elems(LOG2N, LAYOUT_ROW, VAR_PLAIN)};
        g_table[LOG2N][0][0][VAR_PLAIN] = Entry{body_tpl<LOG2N, LAYOUT_ROW, -1, VAR_PLAIN>, nt, smem_elems(LOG2N, LAYOUT_ROW, VAR_PLAIN)};
        g_table[LOG2N][0][1][VAR_REAL] = Entry{body_tpl<LOG2N, LAYOUT_ROW, 1, VAR_REAL>, nt, smem_elems(LOG2N, LAYOUT_ROW, VAR_REAL)};
        g_table[LOG2N][0][0][VAR_REAL] = Entry{body_tpl<LOG2N, LAYOUT_ROW, -1, VAR_REAL>, nt, smem_elems(LOG2N, LAYOUT_ROW, VAR_REAL)};
    } else {
        g_table[LOG2N][1][1][VAR_PLAIN] = Entry{body_tpl<LOG2N, LAYOUT_COL, 1, VAR_PLAIN>, nt, smem_elems(LOG2N, LAYOUT_COL, VAR_PLAIN)};
        g_table[LOG2N][1][0][VAR_PLAIN] = Entry{body_tpl<LOG2N, LAYOUT_COL, -1, VAR_PLAIN>, nt, smem_elems(LOG2N, LAYOUT_COL, VAR_PLAIN)};
        g_table[LOG2N][1][1][VAR_XPOSE] = Entry{body_tpl<LOG2N, LAYOUT_COL, 1, VAR_XPOSE>, nt, smem_elems(LOG2N, LAYOUT_COL, VAR_XPOSE)};
        g_table[LOG2N][1][0][VAR_XPOSE] = Entry{body_tpl<LOG2N, LAYOUT_COL, -1, VAR_XPOSE>, nt, smem_elems(LOG2N, LAYOUT_COL, VAR_XPOSE)};
    }
}

// big-tile passes (fft_pass2.cuh): one emulated CTA runs the whole persistent loop first, first + stride, ...
struct BigParams { nrb::PassParams p; unsigned stride, ntiles; };
template <int LOG2N, int LAYOUT, int DIR, int VARIANT>
static void body2_tpl(const void *p, double2 *sm, unsigned first, int tid)
{
    const BigParams *b = (const BigParams *)p;
    nrb::fft_pass2_cta<LOG2N, LAYOUT, DIR, VARIANT>(b->p, sm, first, b->stride, b->ntiles, tid);
}
struct Entry2 { BodyFn fn; int nthreads; size_t smem; int lines; };
static Entry2 g_table2[nrb::kMaxLog2N + 1][2][2][3];
template <int LOG2N, int LAYOUT, int VARIANT> static void reg2()
{
    typedef nrb::Geo2<LOG2N, LAYOUT, VARIANT> G;
    const size_t smem = (G::SMEM_BYTES + 15) / 16;
    g_table2[LOG2N][LAYOUT][1][VARIANT] = Entry2{body2_tpl<LOG2N, LAYOUT, 1, VARIANT>, G::NT, smem, G::L};
    g_table2[LOG2N][LAYOUT][0][VARIANT] = Entry2{body2_tpl<LOG2N, LAYOUT, -1, VARIANT>, G::NT, smem, G::L};
}

static void fill_table();
static void init_table()
{
    static std::once_flag once;      // several device workers (multi.cpp) may launch their first kernel at the same time
    std::call_once(once, fill_table);
}
static void fill_table()
{
    using namespace nrb;
    reg<1, 0>(); reg<2, 0>(); reg<3, 0>(); reg<4, 0>(); reg<5, 0>(); reg<6, 0>(); reg<7, 0>();
    reg<8, 0>(); reg<9, 0>(); reg<10, 0>(); reg<11, 0>(); reg<12, 0>(); reg<13, 0>();
    reg<1, 1>(); reg<2, 1>(); reg<3, 1>(); reg<4, 1>(); reg<5, 1>(); reg<6, 1>(); reg<7, 1>();
    reg<8, 1>(); reg<9, 1>(); reg<10, 1>(); reg<11, 1>(); reg<12, 1>();
    // same set as the CUDA build (k_big.cu)
    reg2<11, 0, VAR_PLAIN>(); reg2<12, 0, VAR_PLAIN>(); reg2<13, 0, VAR_PLAIN>();
    reg2<9, 1, VAR_PLAIN>(); reg2<10, 1, VAR_PLAIN>(); reg2<9, 1, VAR_XPOSE>(); reg2<10, 1, VAR_XPOSE>();
}

} // namespace nrb_emu

namespace nrb {

static thread_local std::string g_emu_err;
const char *be_last_error() { return g_emu_err.c_str(); }

static long g_pass_launches = 0, g_aux_launches = 0, g_big_launches = 0;

static long g_simple_launches = 0;
int be_launch_pass(const KernelKey &key, const PassParams &p_in, u64 ntiles, void *)
{
    PassParams p = p_in;
    p.simple = pass_is_simple(key, p_in) ? 1 : 0;
    if (p.simple) ++g_simple_launches;
    nrb_emu::init_table();
    if (key.log2n < 1 || key.log2n > kMaxLog2N) { g_emu_err = "no such kernel"; return -1; }
    const nrb_emu::Entry e = nrb_emu::g_table[key.log2n][key.layout][key.dir > 0 ? 1 : 0][key.variant];
    if (use_big_tiles(key, p) && nrb_emu::g_table2[key.log2n][key.layout][key.dir > 0 ? 1 : 0][key.variant].fn) {
        const nrb_emu::Entry2 e2 = nrb_emu::g_table2[key.log2n][key.layout][key.dir > 0 ? 1 : 0][key.variant];
        const u64 lines = p.q_end - p.q_begin;
        nrb_emu::BigParams bp;
        bp.p = p;
        bp.ntiles = (unsigned)((lines + e2.lines - 1) / e2.lines);
        bp.stride = 3;      // a "grid" of 3 persistent CTAs: every CTA walks several tiles when there are more than 3
        ++g_pass_launches;
        ++g_big_launches;
#pragma omp parallel for schedule(dynamic)
        for (int first = 0; first < (int)bp.stride; ++first) {
            nrb_emu::Cta c;
            c.fibers.resize(e2.nthreads);
            c.stacks.resize((size_t)e2.nthreads * nrb_emu::kStack);
            c.done.resize(e2.nthreads);
            c.smem.assign(e2.smem, make_double2(__builtin_nan(""), __builtin_nan("")));
            std::vector<double> shfl_buf(e2.nthreads + 32, 0.0);
            nrb_emu::t_shfl = &shfl_buf;
            c.body = e2.fn;
            c.params = &bp;
            c.tile = (unsigned)first;
            nrb_emu::run_cta(c, e2.nthreads);
        }
        return 0;
    }
    if (!e.fn) { g_emu_err = "kernel variant not built"; return -1; }
    ++g_pass_launches;
#pragma omp parallel
    {
        nrb_emu::Cta c;
        c.fibers.resize(e.nthreads);
        c.stacks.resize((size_t)e.nthreads * nrb_emu::kStack);
        c.done.resize(e.nthreads);
        c.smem.assign(e.smem, make_double2(0.0, 0.0));
        std::vector<double> shfl_buf(e.nthreads + 32, 0.0);
        nrb_emu::t_shfl = &shfl_buf;
        c.body = e.fn;
        c.params = &p;
#pragma omp for schedule(dynamic)
        for (long long t = 0; t < (long long)ntiles; ++t) {
            c.tile = (unsigned)t;
            // poison shared memory so reads of never-written cells show up as NaN
            for (size_t i = 0; i < c.smem.size(); ++i) c.smem[i] = make_double2(__builtin_nan(""), __builtin_nan(""));
            nrb_emu::run_cta(c, e.nthreads);
        }
    }
    return 0;
}

template <int LOG2R> static void mid_body(const void *p, double2 *sm, unsigned tile, int tid)
{
    conv_mid_cta<LOG2R>(*(const ConvMidParams *)p, sm, tile, tid);
}
bool be_conv_mid_available(int log2rest) { return log2rest == 4 || log2rest == 6 || log2rest == 11 || log2rest == 12; }   // as k_mid.cu
static long g_mid_launches = 0;
int be_launch_conv_mid(int log2rest, const ConvMidParams &m, u64 ntiles, void *)
{
    nrb_emu::BodyFn fn = nullptr;
    int nt = 0;
    size_t smem = 0;
    switch (log2rest) {
    case 4: fn = mid_body<4>; nt = GeoM<4>::NT; smem = GeoM<4>::SMEM_BYTES / 16; break;
    case 6: fn = mid_body<6>; nt = GeoM<6>::NT; smem = GeoM<6>::SMEM_BYTES / 16; break;
    case 11: fn = mid_body<11>; nt = GeoM<11>::NT; smem = GeoM<11>::SMEM_BYTES / 16; break;
    case 12: fn = mid_body<12>; nt = GeoM<12>::NT; smem = GeoM<12>::SMEM_BYTES / 16; break;
    default: g_emu_err = "conv_mid kernel not built for this row length"; return -1;
    }
    ++g_mid_launches;
#pragma omp parallel
    {
        nrb_emu::Cta c;
        c.fibers.resize(nt);
        c.stacks.resize((size_t)nt * nrb_emu::kStack);
        c.done.resize(nt);
        std::vector<double> shfl_buf(nt + 32, 0.0);
        nrb_emu::t_shfl = &shfl_buf;
        c.body = fn;
        c.params = &m;
#pragma omp for schedule(dynamic)
        for (long long t = 0; t < (long long)ntiles; ++t) {
            c.tile = (unsigned)t;
            c.smem.assign(smem, make_double2(__builtin_nan(""), __builtin_nan("")));
            nrb_emu::run_cta(c, nt);
        }
    }
    return 0;
}

// one-kernel cosft1 / cosft2 / sinft and twofft (trig_fused.cuh): same set of lengths as k_trig.cu
template <int LOG2N> static void trig_body(const void *p, double2 *sm, unsigned tile, int tid)
{
    trig_cta<LOG2N>(*(const TrigParams *)p, sm, tile, tid);
}
template <int LOG2N> static void twofft_body(const void *p, double2 *sm, unsigned tile, int tid)
{
    twofft_cta<LOG2N>(*(const TwoFFTParams *)p, sm, tile, tid);
}
bool be_trig_available(int log2n) { return log2n >= kTrigMinLog2 && log2n <= kTrigMaxLog2; }
static long g_trig_launches = 0;
template <int LOG2N> static void trig_geo(bool two, nrb_emu::BodyFn &fn, int &nt, size_t &smem, int &lines)
{
    fn = two ? twofft_body<LOG2N> : trig_body<LOG2N>;
    nt = GeoT<LOG2N>::NT; smem = (GeoT<LOG2N>::SMEM_BYTES + 15) / 16; lines = GeoT<LOG2N>::L;
}
static int emu_launch_trig(int log2n, bool two, const void *params, u64 count)
{
    nrb_emu::BodyFn fn = nullptr;
    int nt = 0, lines = 1;
    size_t smem = 0;
    switch (log2n) {
    case 3: trig_geo<3>(two, fn, nt, smem, lines); break;
    case 4: trig_geo<4>(two, fn, nt, smem, lines); break;
    case 5: trig_geo<5>(two, fn, nt, smem, lines); break;
    case 6: trig_geo<6>(two, fn, nt, smem, lines); break;
    case 7: trig_geo<7>(two, fn, nt, smem, lines); break;
    case 8: trig_geo<8>(two, fn, nt, smem, lines); break;
    case 9: trig_geo<9>(two, fn, nt, smem, lines); break;
    case 10: trig_geo<10>(two, fn, nt, smem, lines); break;
    case 11: trig_geo<11>(two, fn, nt, smem, lines); break;
    case 12: trig_geo<12>(two, fn, nt, smem, lines); break;
    case 13: trig_geo<13>(two, fn, nt, smem, lines); break;
    default: g_emu_err = "trig kernel not built for this line length"; return -1;
    }
    ++g_trig_launches;
    const long long ntiles = (long long)((count + (u64)lines - 1) / (u64)lines);
#pragma omp parallel
    {
        nrb_emu::Cta c;
        c.fibers.resize(nt);
        c.stacks.resize((size_t)nt * nrb_emu::kStack);
        c.done.resize(nt);
        std::vector<double> shfl_buf(nt + 32, 0.0);
        nrb_emu::t_shfl = &shfl_buf;
        c.body = fn;
        c.params = params;
#pragma omp for schedule(dynamic)
        for (long long t = 0; t < ntiles; ++t) {
            c.tile = (unsigned)t;
            c.smem.assign(smem, make_double2(__builtin_nan(""), __builtin_nan("")));
            nrb_emu::run_cta(c, nt);
        }
    }
    return 0;
}
int be_launch_trig(int log2n, const TrigParams &t, void *) { return emu_launch_trig(log2n, false, &t, t.count); }
int be_launch_twofft(int log2n, const TwoFFTParams &t, void *) { return emu_launch_trig(log2n, true, &t, t.count); }

bool be_fused_available(const KernelKey &a, const KernelKey &b)
{
    // same set as the CUDA build (k_fused_*.cu): z in {128,256,512} (ROW REAL), y in {256,512,1024} (COL PLAIN)
    const KernelKey &z = a.layout == LAYOUT_ROW ? a : b, &y = a.layout == LAYOUT_ROW ? b : a;
    const bool order_ok = (a.dir > 0) == (a.layout == LAYOUT_ROW);
    return a.dir == b.dir && order_ok && z.layout == LAYOUT_ROW && z.variant == VAR_REAL && y.layout == LAYOUT_COL &&
           y.variant == VAR_PLAIN && z.log2n >= 7 && z.log2n <= 9 && y.log2n >= 8 && y.log2n <= 10 &&
           cta_threads(z.log2n, LAYOUT_ROW) == cta_threads(y.log2n, LAYOUT_COL);
}

int be_launch_fused(const KernelKey &ka, const PassParams &pa, const KernelKey &kb, const PassParams &pb, const FuseSched &fs, void *st)
{
    // the dependency structure (all of A's tiles of a unit before B's) is trivially met by running A, then B
    int rc = be_launch_pass(ka, pa, (u64)fs.units * fs.ta, st);
    if (rc == 0) rc = be_launch_pass(kb, pb, (u64)fs.units * fs.tb, st);
    return rc;
}

static void scan_cta_body(const void *p, double2 *sm, unsigned block, int tid)
{
    aux_scan_cta(*(const AuxParams *)p, sm, block, tid);
}

int be_launch_aux(const AuxParams &a, void *)
{
    ++g_aux_launches;
    if (a.kind == AUX_SCAN && a.op != 1) {      // block-cooperative kernel: one emulated CTA per chunk
        const u64 blocks = a.count * ((a.n / 2 + kScanChunk - 1) / kScanChunk);
        nrb_emu::Cta c;
        c.fibers.resize(kScanThreads);
        c.stacks.resize((size_t)kScanThreads * nrb_emu::kStack);
        c.done.resize(kScanThreads);
        c.body = scan_cta_body;
        c.params = &a;
        for (u64 b = 0; b < blocks; ++b) {
            c.tile = (unsigned)b;
            c.smem.assign(kScanSmem, make_double2(__builtin_nan(""), __builtin_nan("")));
            nrb_emu::run_cta(c, kScanThreads);
        }
        return 0;
    }
    const u64 nthreads = 977;   // deliberately odd: exercises the grid-stride loops
    for (u64 t = 0; t < nthreads; ++t) aux_body(a, t, nthreads);
    return 0;
}

void *be_event_record(void *)
{
    struct timespec *ts = new timespec;
    clock_gettime(CLOCK_MONOTONIC, ts);
    return ts;
}
void *be_event_create() { return new timespec; }
int be_event_record_on(void *, void *) { return 0; }
int be_stream_wait(void *, void *) { return 0; }
int be_stream_create_prio(void **s, int) { *s = (void *)1; return 0; }
float be_event_elapsed_ms(void *a, void *b)
{
    const timespec *x = (const timespec *)a, *y = (const timespec *)b;
    return (float)((y->tv_sec - x->tv_sec) * 1e3 + (y->tv_nsec - x->tv_nsec) * 1e-6);
}
void be_event_destroy(void *e) { delete (timespec *)e; }
int be_malloc(void **p, size_t bytes) { *p = malloc(bytes ? bytes : 16); return *p ? 0 : -1; }
int be_free(void *p) { free(p); return 0; }
int be_memset(void *p, int v, size_t n, void *) { memset(p, v, n); return 0; }
int be_ipc_export(void *p, unsigned char h[64]) { memset(h, 0, 64); memcpy(h, &p, sizeof(p)); return 0; }
int be_ipc_import(const unsigned char h[64], void **p) { memcpy(p, h, sizeof(*p)); return 0; }
int be_ipc_release(void *) { return 0; }
int be_h2d(void *d, const void *s, size_t n, void *) { memcpy(d, s, n); return 0; }
int be_d2h(void *d, const void *s, size_t n, void *) { memcpy(d, s, n); return 0; }
int be_d2d(void *d, const void *s, size_t n, void *) { memmove(d, s, n); return 0; }
int be_sync(void *) { return 0; }
// NRB_EMU_DEVICES = n pretends there are n devices (all of them host memory) so the multi-device host-slice paths of
// multi.cpp can run in the CPU test tier; the current device is per thread, as in CUDA
static int emu_devices()
{
    const char *v = getenv("NRB_EMU_DEVICES");
    const int n = (v && *v) ? atoi(v) : 1;
    return n < 1 ? 1 : n > 8 ? 8 : n;
}
static thread_local int t_emu_device = 0;
int be_current_device() { return t_emu_device; }
int be_device_count() { return emu_devices(); }
int be_set_device(int dev) { if (dev < 0 || dev >= emu_devices()) return -1; t_emu_device = dev; return 0; }
static int copy_2d(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t rows)
{
    for (size_t r = 0; r < rows; ++r) memcpy((char *)d + r * dp, (const char *)s + r * sp, w);
    return 0;
}
int be_h2d_2d(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t rows, void *) { return copy_2d(d, dp, s, sp, w, rows); }
int be_d2h_2d(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t rows, void *) { return copy_2d(d, dp, s, sp, w, rows); }
int be_enable_peer(int dev, int peer) { return (be_set_device(dev) == 0 && peer >= 0 && peer < emu_devices()) ? 0 : -1; }
int be_stream_create(void **s) { *s = (void *)1; return 0; }
int be_stream_destroy(void *) { return 0; }
void *be_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 16); }
void be_host_free(void *p) { free(p); }

} // namespace nrb

extern "C" long nrb_emu_launch_count(int aux) { return aux == 5 ? nrb::g_trig_launches : aux == 4 ? nrb::g_mid_launches : aux == 3 ? nrb::g_simple_launches : aux == 2 ? nrb::g_big_launches : aux ? nrb::g_aux_launches : nrb::g_pass_launches; }
