// api.cpp -- the C ABI of include/numrs_b200.h: device-resident plan API and the host-slice
// drop-in entry points (H2D copy, transform on the device, D2H copy, blocking).
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#pragma GCC visibility push(default)
#include "../../include/numrs_b200.h"
#pragma GCC visibility pop
#include "multi.h"
#include "plan.h"

using namespace nrb;

struct nrb_plan_s {
    Plan plan;
    std::mutex mu;   // serialises exec of one plan (its workspace is shared)
    ~nrb_plan_s()
    {
        if (plan.ws) be_free(plan.ws);
        if (plan.sched) be_free(plan.sched);
        release_side_lane(plan.side);
    }
};
struct nrb_slab_s {
    SlabPlan plan;
    ~nrb_slab_s() { slab_release(plan); }
};

namespace {

// ---- per-thread host-call context ----
struct DevBuf {
    void *p;
    size_t cap;
    DevBuf() : p(nullptr), cap(0) {}
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) be_free(p);
        p = nullptr; cap = 0;
        if (be_malloc(&p, bytes) != 0) return -1;
        cap = bytes;
        return 0;
    }
    void release() { if (p) be_free(p); p = nullptr; cap = 0; }
};

// Every thread that makes host-slice calls owns a stream and three staging buffers.  The contexts are registered so
// that nrb_shutdown can release the device memory of ALL threads, and a thread that exits releases its own.
struct ThreadCtx;
std::mutex g_ctx_mu;
std::set<ThreadCtx *> g_ctxs;

struct ThreadCtx {
    void *stream;
    void *s_in, *s_out;      // copy streams of the pipelined batch calls (created on first use)
    int device;
    DevBuf io, aux, out;
    ThreadCtx() : stream(nullptr), s_in(nullptr), s_out(nullptr), device(-1)
    {
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        g_ctxs.insert(this);
    }
    void release()
    {
        io.release(); aux.release(); out.release();
        if (stream) { be_stream_destroy(stream); stream = nullptr; }
        if (s_in) { be_stream_destroy(s_in); s_in = nullptr; }
        if (s_out) { be_stream_destroy(s_out); s_out = nullptr; }
        device = -1;
    }
    ~ThreadCtx()
    {
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        release();
        g_ctxs.erase(this);
    }
};
thread_local ThreadCtx t_ctx;

// Plan cache of the host-slice entry points: least recently used first out, bounded by entry count and by the device
// workspace the cached plans own (NRB_PLAN_CACHE_MB, default 16 GiB); plans in use stay alive through their shared_ptr.
struct CacheEntry { std::shared_ptr<nrb_plan_s> plan; unsigned long long stamp; };
std::mutex g_cache_mu;
std::map<std::string, CacheEntry> g_plan_cache;
unsigned long long g_cache_clock = 0;

size_t plan_cache_cap_bytes()
{
    static const size_t cap = [] {
        const char *v = getenv("NRB_PLAN_CACHE_MB");
        const long mb = (v && *v) ? atol(v) : 16384;
        return (size_t)(mb < 0 ? 0 : mb) << 20;
    }();
    return cap;
}

void cache_evict_locked(size_t incoming_bytes)
{
    for (;;) {
        size_t total = incoming_bytes;
        auto oldest = g_plan_cache.end();
        for (auto it = g_plan_cache.begin(); it != g_plan_cache.end(); ++it) {
            total += it->second.plan->plan.ws_elems * sizeof(double2);
            if (oldest == g_plan_cache.end() || it->second.stamp < oldest->second.stamp) oldest = it;
        }
        if (oldest == g_plan_cache.end() || (g_plan_cache.size() < 64 && total <= plan_cache_cap_bytes())) return;
        g_plan_cache.erase(oldest);
    }
}

int fail(int code, const std::string &msg) { set_error(msg); return code; }

// No C++ exception may cross the C ABI (the callers are Rust / C / ctypes): every entry point that can allocate is a
// function-try-block ending in this handler.
int on_exception() noexcept
{
    try { throw; }
    catch (const std::bad_alloc &) {
        try { set_error("out of host memory"); } catch (...) {}
        return NRB_ERR_OOM;
    }
    catch (const std::exception &e) {
        try { set_error(std::string("internal error: ") + e.what()); } catch (...) {}
        return NRB_ERR_CUDA;
    }
    catch (...) {
        return NRB_ERR_CUDA;
    }
}

int ensure_ctx()
{
    if (be_device_count() <= 0) return fail(NRB_ERR_CUDA, "no CUDA device available (numrs_b200 has no CPU fallback)");
    const int dev = be_current_device();
    if (dev < 0) return fail(NRB_ERR_CUDA, std::string("cannot query CUDA device: ") + be_last_error());
    if (t_ctx.stream && t_ctx.device != dev) t_ctx.release();   // thread moved to another device
    if (!t_ctx.stream) {
        if (be_stream_create(&t_ctx.stream) != 0) return fail(NRB_ERR_CUDA, std::string("stream creation failed: ") + be_last_error());
        t_ctx.device = dev;
    }
    return NRB_OK;
}

std::shared_ptr<nrb_plan_s> cached_plan(int kind, const size_t *dims, size_t ndim, size_t batch, int *rc)
{
    std::string key = std::to_string(be_current_device()) + ":" + std::to_string(kind) + ":" + std::to_string(batch);
    for (size_t d = 0; d < ndim; ++d) key += ":" + std::to_string(dims[d]);
    std::lock_guard<std::mutex> lk(g_cache_mu);
    auto it = g_plan_cache.find(key);
    if (it != g_plan_cache.end()) { *rc = NRB_OK; it->second.stamp = ++g_cache_clock; return it->second.plan; }
    std::shared_ptr<nrb_plan_s> h(new nrb_plan_s());
    *rc = build_plan(h->plan, kind, dims, ndim, batch);
    if (*rc == NRB_ERR_OOM && !g_plan_cache.empty()) {
        // the cache itself may be what holds the memory: drop it and try once more
        g_plan_cache.clear();
        h.reset(new nrb_plan_s());
        *rc = build_plan(h->plan, kind, dims, ndim, batch);
    }
    if (*rc != NRB_OK) return nullptr;
    cache_evict_locked(h->plan.ws_elems * sizeof(double2));
    g_plan_cache[key] = CacheEntry{h, ++g_cache_clock};
    return h;
}

// staging buffer of a host-slice call; on failure the plan cache (which may be what holds the memory) is dropped and the
// allocation retried once
int ensure_buf(DevBuf &b, size_t bytes)
{
    if (b.ensure(bytes) == 0) return 0;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        g_plan_cache.clear();
    }
    return b.ensure(bytes);
}

int copy_fail(const char *what) { return fail(NRB_ERR_CUDA, std::string(what) + " failed: " + be_last_error()); }

// ---- batches in chunks over three streams: the H2D copy of chunk c + 1, the transforms of chunk c and the D2H copy of
// chunk c - 1 overlap, so a batch call keeps both directions of the PCIe link busy instead of using them one after the
// other (the transforms themselves are a few per cent of the call).  Units of a batch are independent (fft_batch
// FFT_1.rs:185, convlv_batch Convolve.rs:241, correl_batch Correlation.rs:273), so any split gives the same results.
size_t pipeline_chunk(size_t count, size_t unit_bytes)
{
    const size_t total = count * unit_bytes, min_chunk = (size_t)tunables().pipeline_min_kb << 10;
    if (!tunables().pipeline_batches || min_chunk == 0 || count < 4 || total < 4 * min_chunk) return 0;   // not worth it: one shot
    size_t want = total / 8 > min_chunk ? total / 8 : min_chunk;
    size_t per = (want + unit_bytes - 1) / unit_bytes;
    if (per < 1) per = 1;
    return per >= count ? 0 : per;
}

// runs h2d(first, n, stream), exec(first, n, stream), d2h(first, n, stream) for every chunk with the ordering above
std::atomic<long> g_pipelined_calls{0};

template <class H2D, class EXEC, class D2H>
int run_pipelined(size_t count, size_t per, H2D h2d, EXEC exec, D2H d2h)
{
    ++g_pipelined_calls;
    if ((!t_ctx.s_in && be_stream_create(&t_ctx.s_in) != 0) || (!t_ctx.s_out && be_stream_create(&t_ctx.s_out) != 0))
        return fail(NRB_ERR_CUDA, std::string("stream creation failed: ") + be_last_error());
    const size_t nchunks = (count + per - 1) / per;
    std::vector<void *> ev(2 * nchunks, nullptr);
    int rc = NRB_OK;
    for (size_t c = 0; c < nchunks && rc == NRB_OK; ++c) {
        const size_t first = c * per, n = first + per <= count ? per : count - first;
        void *e_in = ev[2 * c] = be_event_create(), *e_cmp = ev[2 * c + 1] = be_event_create();
        if (!e_in || !e_cmp) { rc = fail(NRB_ERR_CUDA, "event creation failed"); break; }
        if ((rc = h2d(first, n, t_ctx.s_in)) != NRB_OK) break;
        if (be_event_record_on(e_in, t_ctx.s_in) != 0 || be_stream_wait(t_ctx.stream, e_in) != 0) { rc = copy_fail("stream ordering"); break; }
        if ((rc = exec(first, n, t_ctx.stream)) != NRB_OK) break;
        if (be_event_record_on(e_cmp, t_ctx.stream) != 0 || be_stream_wait(t_ctx.s_out, e_cmp) != 0) { rc = copy_fail("stream ordering"); break; }
        rc = d2h(first, n, t_ctx.s_out);
    }
    // everything enqueued must have finished before the buffers, the plans or the caller's memory are touched again
    const int s1 = be_sync(t_ctx.s_in), s2 = be_sync(t_ctx.stream), s3 = be_sync(t_ctx.s_out);
    if (rc == NRB_OK && (s1 != 0 || s2 != 0 || s3 != 0)) rc = copy_fail("pipelined batch execution");
    for (void *e : ev) if (e) be_event_destroy(e);
    return rc;
}

// in-place transform of `count` host slices of `doubles` doubles each
int run_inplace(int kind, const size_t *dims, size_t ndim, double *const *ptrs, size_t count, size_t doubles,
                int isign, double *speq, size_t speq_doubles)
{
    int rc = ensure_ctx();
    if (rc) return rc;
    const size_t bytes = doubles * sizeof(double);
    const size_t per = speq ? 0 : pipeline_chunk(count, bytes);
    if (per) {
        // chunks of `per` slices (and a shorter last one): one plan each, both held for the whole call
        const size_t tail = count % per;
        auto hp = cached_plan(kind, dims, ndim, per, &rc);
        if (!hp) return rc;
        std::shared_ptr<nrb_plan_s> ht = hp;
        if (tail && !(ht = cached_plan(kind, dims, ndim, tail, &rc))) return rc;
        if (ensure_buf(t_ctx.io, bytes * count) != 0) return fail(NRB_ERR_OOM, std::string("device allocation failed: ") + be_last_error());
        char *dio = (char *)t_ctx.io.p;
        std::unique_lock<std::mutex> l1(hp->mu), l2;
        if (ht != hp) l2 = std::unique_lock<std::mutex>(ht->mu);
        auto copy = [&](bool in, size_t first, size_t n, void *st) -> int {
            for (size_t b = first; b < first + n;) {
                size_t e = b + 1;
                while (e < first + n && ptrs[e] == ptrs[b] + (e - b) * doubles) ++e;
                const int r = in ? be_h2d(dio + b * bytes, ptrs[b], bytes * (e - b), st) : be_d2h(ptrs[b], dio + b * bytes, bytes * (e - b), st);
                if (r != 0) return copy_fail(in ? "host-to-device copy" : "device-to-host copy");
                b = e;
            }
            return NRB_OK;
        };
        return run_pipelined(count, per,
                             [&](size_t f, size_t n, void *st) { return copy(true, f, n, st); },
                             [&](size_t f, size_t n, void *st) { return exec_plan((n == per ? hp : ht)->plan, (double *)(dio + f * bytes), nullptr, nullptr, isign == 1 ? 1 : -1, 0, st); },
                             [&](size_t f, size_t n, void *st) { return copy(false, f, n, st); });
    }
    auto h = cached_plan(kind, dims, ndim, count, &rc);
    if (!h) return rc;
    if (ensure_buf(t_ctx.io, bytes * count) != 0) return fail(NRB_ERR_OOM, std::string("device allocation failed: ") + be_last_error());
    if (speq && ensure_buf(t_ctx.aux, speq_doubles * sizeof(double)) != 0) return fail(NRB_ERR_OOM, "device allocation failed");
    void *s = t_ctx.stream;
    char *dio = (char *)t_ctx.io.p;
    // coalesce runs of host slices that are contiguous in memory into one copy
    for (size_t b = 0; b < count;) {
        size_t e = b + 1;
        while (e < count && ptrs[e] == ptrs[b] + (e - b) * doubles) ++e;
        if (be_h2d(dio + b * bytes, ptrs[b], bytes * (e - b), s) != 0) return copy_fail("host-to-device copy");
        b = e;
    }
    // rlft3 inverse reads speq; forward overwrites it
    if (speq && isign != 1 && be_h2d(t_ctx.aux.p, speq, speq_doubles * sizeof(double), s) != 0) return copy_fail("host-to-device copy");
    {
        std::lock_guard<std::mutex> lk(h->mu);
        rc = exec_plan(h->plan, (double *)t_ctx.io.p, (double *)t_ctx.aux.p, nullptr, isign == 1 ? 1 : -1, 0, s);
        if (rc == NRB_OK && be_sync(s) != 0) rc = copy_fail("kernel execution");
    }
    if (rc) return rc;
    for (size_t b = 0; b < count;) {
        size_t e = b + 1;
        while (e < count && ptrs[e] == ptrs[b] + (e - b) * doubles) ++e;
        if (be_d2h(ptrs[b], dio + b * bytes, bytes * (e - b), s) != 0) return copy_fail("device-to-host copy");
        b = e;
    }
    if (speq && isign == 1 && be_d2h(speq, t_ctx.aux.p, speq_doubles * sizeof(double), s) != 0) return copy_fail("device-to-host copy");
    if (be_sync(s) != 0) return copy_fail("device-to-host copy");
    return NRB_OK;
}

// out-of-place batch: io (count x n), aux (aux_count x aux_n), out (count x n)
int run_outofplace(int kind, const size_t *dims, size_t ndim, const double *const *in, const double *const *aux,
                   size_t aux_count, size_t aux_n, double *const *out, size_t count, size_t n, int isign, int arg)
{
    int rc = ensure_ctx();
    if (rc) return rc;
    const size_t bytes = n * sizeof(double);
    const size_t per = pipeline_chunk(count, bytes);
    if (per) {
        const size_t tail = count % per;
        const bool per_signal_aux = aux_count == count;
        auto hp = cached_plan(kind, dims, ndim, per, &rc);
        if (!hp) return rc;
        std::shared_ptr<nrb_plan_s> ht = hp;
        if (tail && !(ht = cached_plan(kind, dims, ndim, tail, &rc))) return rc;
        if (ensure_buf(t_ctx.io, bytes * count) != 0 || ensure_buf(t_ctx.out, bytes * count) != 0 ||
            ensure_buf(t_ctx.aux, aux_n * sizeof(double) * aux_count) != 0)
            return fail(NRB_ERR_OOM, std::string("device allocation failed: ") + be_last_error());
        std::unique_lock<std::mutex> l1(hp->mu), l2;
        if (ht != hp) l2 = std::unique_lock<std::mutex>(ht->mu);
        const size_t aux_bytes = aux_n * sizeof(double);
        if (!per_signal_aux) {      // one operand for the whole batch (convlv's response): up front, on the compute stream
            for (size_t b = 0; b < aux_count; ++b)
                if (be_h2d((char *)t_ctx.aux.p + b * aux_bytes, aux[b], aux_bytes, t_ctx.stream) != 0) return copy_fail("host-to-device copy");
        }
        return run_pipelined(count, per,
                             [&](size_t f, size_t cnt, void *st) -> int {
                                 for (size_t b = f; b < f + cnt; ++b) {
                                     if (be_h2d((char *)t_ctx.io.p + b * bytes, in[b], bytes, st) != 0) return copy_fail("host-to-device copy");
                                     if (per_signal_aux && be_h2d((char *)t_ctx.aux.p + b * aux_bytes, aux[b], aux_bytes, st) != 0) return copy_fail("host-to-device copy");
                                 }
                                 return NRB_OK;
                             },
                             [&](size_t f, size_t cnt, void *st) {
                                 return exec_plan((cnt == per ? hp : ht)->plan, (double *)((char *)t_ctx.io.p + f * bytes),
                                                  (double *)((char *)t_ctx.aux.p + (per_signal_aux ? f * aux_bytes : 0)),
                                                  (double *)((char *)t_ctx.out.p + f * bytes), isign, arg, st);
                             },
                             [&](size_t f, size_t cnt, void *st) -> int {
                                 for (size_t b = f; b < f + cnt; ++b)
                                     if (be_d2h(out[b], (char *)t_ctx.out.p + b * bytes, bytes, st) != 0) return copy_fail("device-to-host copy");
                                 return NRB_OK;
                             });
    }
    auto h = cached_plan(kind, dims, ndim, count, &rc);
    if (!h) return rc;
    if (ensure_buf(t_ctx.io, bytes * count) != 0 || ensure_buf(t_ctx.out, bytes * count) != 0 ||
        ensure_buf(t_ctx.aux, aux_n * sizeof(double) * aux_count) != 0)
        return fail(NRB_ERR_OOM, std::string("device allocation failed: ") + be_last_error());
    void *s = t_ctx.stream;
    for (size_t b = 0; b < count; ++b)
        if (be_h2d((char *)t_ctx.io.p + b * bytes, in[b], bytes, s) != 0) return copy_fail("host-to-device copy");
    for (size_t b = 0; b < aux_count; ++b)
        if (be_h2d((char *)t_ctx.aux.p + b * aux_n * sizeof(double), aux[b], aux_n * sizeof(double), s) != 0)
            return copy_fail("host-to-device copy");
    {
        std::lock_guard<std::mutex> lk(h->mu);
        rc = exec_plan(h->plan, (double *)t_ctx.io.p, (double *)t_ctx.aux.p, (double *)t_ctx.out.p, isign, arg, s);
        if (rc == NRB_OK && be_sync(s) != 0) rc = copy_fail("kernel execution");
    }
    if (rc) return rc;
    for (size_t b = 0; b < count; ++b)
        if (be_d2h(out[b], (char *)t_ctx.out.p + b * bytes, bytes, s) != 0) return copy_fail("device-to-host copy");
    if (be_sync(s) != 0) return copy_fail("device-to-host copy");
    return NRB_OK;
}

// general out-of-place call: host inputs -> io / aux, exec, selected ranges of the device `out` buffer -> host
struct Seg { const void *host; size_t dev_off, doubles; };   // offsets / sizes in doubles
int run_segments(int kind, const size_t *dims, size_t ndim, size_t batch, const std::vector<Seg> &io, const std::vector<Seg> &aux,
                 size_t out_doubles, const std::vector<Seg> &outs, int isign, int arg)
{
    int rc = ensure_ctx();
    if (rc) return rc;
    auto h = cached_plan(kind, dims, ndim, batch, &rc);
    if (!h) return rc;
    size_t io_d = 0, aux_d = 0;
    for (const Seg &g : io) if (g.dev_off + g.doubles > io_d) io_d = g.dev_off + g.doubles;
    for (const Seg &g : aux) if (g.dev_off + g.doubles > aux_d) aux_d = g.dev_off + g.doubles;
    if (ensure_buf(t_ctx.io, io_d * sizeof(double)) != 0 || ensure_buf(t_ctx.aux, aux_d * sizeof(double)) != 0 ||
        ensure_buf(t_ctx.out, out_doubles * sizeof(double)) != 0)
        return fail(NRB_ERR_OOM, std::string("device allocation failed: ") + be_last_error());
    void *s = t_ctx.stream;
    for (const Seg &g : io)
        if (be_h2d((double *)t_ctx.io.p + g.dev_off, g.host, g.doubles * sizeof(double), s) != 0) return copy_fail("host-to-device copy");
    for (const Seg &g : aux)
        if (be_h2d((double *)t_ctx.aux.p + g.dev_off, g.host, g.doubles * sizeof(double), s) != 0) return copy_fail("host-to-device copy");
    {
        std::lock_guard<std::mutex> lk(h->mu);
        rc = exec_plan(h->plan, (double *)t_ctx.io.p, (double *)t_ctx.aux.p, (double *)t_ctx.out.p, isign, arg, s);
        if (rc == NRB_OK && be_sync(s) != 0) rc = copy_fail("kernel execution");
    }
    if (rc) return rc;
    for (const Seg &g : outs)
        if (be_d2h(const_cast<void *>(g.host), (double *)t_ctx.out.p + g.dev_off, g.doubles * sizeof(double), s) != 0)
            return copy_fail("device-to-host copy");
    if (be_sync(s) != 0) return copy_fail("device-to-host copy");
    return NRB_OK;
}

// ---- batches over several devices (option num_devices): contiguous batch ranges, no communication ----
// (fft_batch FFT_1.rs:185, convlv_batch Convolve.rs:241, correl_batch Correlation.rs:273 shard by batch)
size_t shard_min_bytes() { return (size_t)tunables().shard_min_kb << 10; }   // smaller calls stay on the calling thread's device

int run_inplace_batch(int kind, const size_t *dims, size_t ndim, double *const *ptrs, size_t count, size_t doubles, int isign)
{
    const int G = multi_device_count();
    if (G < 2 || count < 2 || count * doubles * sizeof(double) < shard_min_bytes())
        return run_inplace(kind, dims, ndim, ptrs, count, doubles, isign, nullptr, 0);
    return multi_shard_batch(count, G, [&](size_t first, size_t cnt) {
        return run_inplace(kind, dims, ndim, ptrs + first, cnt, doubles, isign, nullptr, 0);
    });
}

int run_outofplace_batch(int kind, const size_t *dims, size_t ndim, const double *const *in, const double *const *aux,
                         size_t aux_count, size_t aux_n, double *const *out, size_t count, size_t n, int isign, int arg)
{
    const int G = multi_device_count();
    if (G < 2 || count < 2 || count * n * sizeof(double) < shard_min_bytes())
        return run_outofplace(kind, dims, ndim, in, aux, aux_count, aux_n, out, count, n, isign, arg);
    const bool per_signal_aux = aux_count == count;     // correl: one second operand per pair; convlv: one response for all
    return multi_shard_batch(count, G, [&](size_t first, size_t cnt) {
        return run_outofplace(kind, dims, ndim, in + first, per_signal_aux ? aux + first : aux, per_signal_aux ? cnt : aux_count,
                              aux_n, out + first, cnt, n, isign, arg);
    });
}

} // namespace

extern "C" {

const char *nrb_version(void) { return "numrs_b200 0.1.0 (sm_100a)"; }
const char *nrb_last_error(void) { return get_error().c_str(); }
int nrb_device_count(void) { return be_device_count(); }
int nrb_set_device(int device)
try {
    if (be_set_device(device) != 0) return fail(NRB_ERR_CUDA, std::string("cudaSetDevice failed: ") + be_last_error());
    return NRB_OK;
} catch (...) { return on_exception(); }
int nrb_shutdown(void)
{
    multi_release();
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        g_plan_cache.clear();
    }
    {
        // staging buffers and streams of every thread that ever made a host-slice call (the caller guarantees that no
        // other nrb_* call is in flight, as for any shutdown function)
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        for (ThreadCtx *c : g_ctxs) c->release();
    }
    release_tables();
    return NRB_OK;
}
int nrb_set_option(const char *name, long value)
try {
    if (set_tunable(name, value) != 0) return fail(NRB_ERR_INVALID_DIMS, "unknown option");
    multi_release();
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_plan_cache.clear();
    return NRB_OK;
} catch (...) { return on_exception(); }
long nrb_multi_device_calls(int which) { return which == 2 ? g_pipelined_calls.load() : multi_calls(which); }
int nrb_num_devices_in_use(void) { return multi_device_count(); }
void *nrb_host_alloc(size_t bytes) { return be_host_alloc(bytes); }
void nrb_host_free(void *p) { if (p) be_host_free(p); }

// ------------------------------------------------------------------ plan API
int nrb_plan_create(int kind, const size_t *dims, size_t ndim, size_t batch, nrb_plan_t *plan)
try {
    if (!plan) return fail(NRB_ERR_INVALID_DIMS, "plan pointer is NULL");
    *plan = nullptr;
    if (be_device_count() <= 0) return fail(NRB_ERR_CUDA, "no CUDA device available (numrs_b200 has no CPU fallback)");
    std::unique_ptr<nrb_plan_s> h(new nrb_plan_s());
    const int rc = build_plan(h->plan, kind, dims, ndim, batch);
    if (rc != NRB_OK) return rc;
    *plan = h.release();
    return NRB_OK;
} catch (...) { return on_exception(); }
size_t nrb_plan_workspace_bytes(nrb_plan_t plan) { return plan ? plan->plan.ws_elems * sizeof(double2) : 0; }
int nrb_plan_num_launches(nrb_plan_t plan, int isign)
try {
    return plan ? (int)plan->plan.prog[isign == 1 ? 0 : 1].steps.size() : 0;
} catch (...) { return on_exception(); }
int nrb_plan_exec(nrb_plan_t plan, double *d_io, double *d_aux, double *d_out, int isign, int arg, void *stream)
{
    if (!plan) return fail(NRB_ERR_INVALID_DIMS, "plan is NULL");
    return exec_plan(plan->plan, d_io, d_aux, d_out, isign, arg, stream);
}
int nrb_plan_profile(nrb_plan_t plan, double *d_io, double *d_aux, double *d_out, int isign, int arg, void *stream,
                     float *ms, int cap)
try {
    if (!plan || !ms) return fail(NRB_ERR_INVALID_DIMS, "plan is NULL");
    return profile_plan(plan->plan, d_io, d_aux, d_out, isign, arg, stream, ms, cap);
} catch (...) { return on_exception(); }
int nrb_plan_describe_launch(nrb_plan_t plan, int isign, int idx, char *name, size_t cap, double *bytes)
{
    if (!plan) return fail(NRB_ERR_INVALID_DIMS, "plan is NULL");
    return describe_launch(plan->plan, isign, idx, name, cap, bytes);
}
int nrb_fill_uniform_device(double *d_out, unsigned long long seed, unsigned long long offset, size_t count, void *stream)
try {
    if (be_device_count() <= 0) return fail(NRB_ERR_CUDA, "no CUDA device available (numrs_b200 has no CPU fallback)");
    return fill_uniform_device(d_out, seed, offset, count, stream);
} catch (...) { return on_exception(); }
int nrb_upload(void *d_dst, const void *h_src, size_t bytes, void *stream)
{
    if (be_device_count() <= 0) return fail(NRB_ERR_CUDA, "no CUDA device available (numrs_b200 has no CPU fallback)");
    if (bytes && (!d_dst || !h_src)) return fail(NRB_ERR_EMPTY_INPUT, "null pointer");
    if (bytes && be_h2d(d_dst, h_src, bytes, stream) != 0) return copy_fail("host-to-device copy");
    return NRB_OK;
}
int nrb_download(void *h_dst, const void *d_src, size_t bytes, void *stream)
try {
    if (be_device_count() <= 0) return fail(NRB_ERR_CUDA, "no CUDA device available (numrs_b200 has no CPU fallback)");
    if (bytes && (!h_dst || !d_src)) return fail(NRB_ERR_EMPTY_INPUT, "null pointer");
    if (bytes && be_d2h(h_dst, d_src, bytes, stream) != 0) return copy_fail("device-to-host copy");
    return NRB_OK;
} catch (...) { return on_exception(); }
int nrb_stream_synchronize(void *stream)
{
    if (be_device_count() <= 0) return fail(NRB_ERR_CUDA, "no CUDA device available (numrs_b200 has no CPU fallback)");
    if (be_sync(stream) != 0) return copy_fail("stream synchronisation");
    return NRB_OK;
}
int nrb_complex_multiply_device(double *d_a, const double *d_b, size_t ncomplex, int conj_b, double scale, void *stream)
try {
    if (be_device_count() <= 0) return fail(NRB_ERR_CUDA, "no CUDA device available (numrs_b200 has no CPU fallback)");
    if (ncomplex == 0) return NRB_OK;
    if (!d_a || !d_b) return fail(NRB_ERR_EMPTY_INPUT, "null pointer");
    return complex_multiply_device(d_a, d_b, ncomplex, conj_b, scale, stream);
} catch (...) { return on_exception(); }
int nrb_plan_destroy(nrb_plan_t plan)
{
    delete plan;
    return NRB_OK;
}

// ------------------------------------------------------------------ slab API
int nrb_slab_create(size_t nn1, size_t nn2, size_t nn3, int nranks, int rank, nrb_slab_t *plan)
try {
    if (!plan) return fail(NRB_ERR_INVALID_DIMS, "plan pointer is NULL");
    *plan = nullptr;
    if (be_device_count() <= 0) return fail(NRB_ERR_CUDA, "no CUDA device available (numrs_b200 has no CPU fallback)");
    std::unique_ptr<nrb_slab_s> h(new nrb_slab_s());
    const int rc = build_slab_plan(h->plan, nn1, nn2, nn3, nranks, rank);
    if (rc != NRB_OK) return rc;
    *plan = h.release();
    return NRB_OK;
} catch (...) { return on_exception(); }
int nrb_slab_create_fourn(size_t nn1, size_t nn2, size_t nn3, int nranks, int rank, nrb_slab_t *plan)
try {
    if (!plan) return fail(NRB_ERR_INVALID_DIMS, "plan pointer is NULL");
    *plan = nullptr;
    if (be_device_count() <= 0) return fail(NRB_ERR_CUDA, "no CUDA device available (numrs_b200 has no CPU fallback)");
    std::unique_ptr<nrb_slab_s> h(new nrb_slab_s());
    const int rc = build_slab_plan(h->plan, nn1, nn2, nn3, nranks, rank, false);
    if (rc != NRB_OK) return rc;
    *plan = h.release();
    return NRB_OK;
} catch (...) { return on_exception(); }
size_t nrb_slab_local_doubles(nrb_slab_t p) { return p ? (p->plan.real ? 1 : 2) * p->plan.nn1 * p->plan.nn2 * p->plan.nn3 / (size_t)p->plan.nranks : 0; }
size_t nrb_slab_speq_doubles(nrb_slab_t p) { return p && p->plan.real ? 2 * p->plan.nn1 * p->plan.nn2 / (size_t)p->plan.nranks : 0; }
size_t nrb_slab_xchg_doubles(nrb_slab_t p) { return p ? 2 * p->plan.xchg_elems() : 0; }
int nrb_slab_num_launches(nrb_slab_t p, int isign)
try {
    if (!p) return 0;
    const int s = isign == 1 ? 0 : 1;
    return (int)(p->plan.prog[s][0].steps.size() + p->plan.prog[s][1].steps.size()) + (p->plan.nranks > 1 ? 1 : 0);
} catch (...) { return on_exception(); }
int nrb_slab_exec(nrb_slab_t p, int isign, double *d_slab, double *d_speq, unsigned long long epoch, void *stream)
try {
    if (!p) return fail(NRB_ERR_INVALID_DIMS, "plan is NULL");
    return exec_slab_fused(p->plan, isign, d_slab, d_speq, epoch, stream);
} catch (...) { return on_exception(); }
int nrb_slab_stage(nrb_slab_t p, int stage, int isign, double *d_slab, double *d_speq, double *d_send, double *d_recv,
                   void *stream)
try {
    if (!p) return fail(NRB_ERR_INVALID_DIMS, "plan is NULL");
    return exec_slab_stage(p->plan, stage, isign, d_slab, d_speq, d_send, d_recv, stream);
} catch (...) { return on_exception(); }
int nrb_slab_set_peers(nrb_slab_t p, void *const *peer_recv, int count)
{
    if (!p) return fail(NRB_ERR_INVALID_DIMS, "plan is NULL");
    return slab_set_peers(p->plan, peer_recv, count);
}
int nrb_slab_set_send_peers(nrb_slab_t p, void *const *peer_send, int count)
{
    if (!p) return fail(NRB_ERR_INVALID_DIMS, "plan is NULL");
    return slab_set_send_peers(p->plan, peer_send, count);
}
int nrb_slab_barrier(nrb_slab_t p, int phase, unsigned long long epoch, void *stream)
try {
    if (!p) return fail(NRB_ERR_INVALID_DIMS, "plan is NULL");
    return slab_barrier(p->plan, phase, epoch, stream);
} catch (...) { return on_exception(); }
size_t nrb_slab_recv_bytes(nrb_slab_t p) { return p ? nrb_slab_xchg_doubles(p) * sizeof(double) + 8 * 8 * kSlabMaxChunks : 0; }
int nrb_slab_set_chunks(nrb_slab_t p, int chunks)
try {
    if (!p) return fail(NRB_ERR_INVALID_DIMS, "plan is NULL");
    if (chunks > kSlabMaxChunks) return fail(NRB_ERR_INVALID_DIMS, "slab: at most 16 chunks");
    return slab_set_chunks(p->plan, chunks);
} catch (...) { return on_exception(); }
int nrb_slab_stage_part(nrb_slab_t p, int stage, int part, int isign, double *d_slab, double *d_speq, void *stream)
{
    if (!p) return fail(NRB_ERR_INVALID_DIMS, "plan is NULL");
    return exec_slab_part(p->plan, stage, part, isign, d_slab, d_speq, stream);
}
int nrb_slab_set_dma(nrb_slab_t p, int chunks)
try {
    if (!p) return fail(NRB_ERR_INVALID_DIMS, "plan is NULL");
    return slab_set_dma(p->plan, chunks);
} catch (...) { return on_exception(); }
int nrb_slab_stage_part_xchg(nrb_slab_t p, int stage, int part, int isign, double *d_slab, double *d_speq, double *d_xchg, void *stream)
{
    if (!p) return fail(NRB_ERR_INVALID_DIMS, "plan is NULL");
    return exec_slab_part(p->plan, stage, part, isign, d_slab, d_speq, stream, d_xchg);
}
int nrb_slab_exec_dma(nrb_slab_t p, int isign, double *d_slab, double *d_speq, unsigned long long epoch, void *stream)
try {
    if (!p) return fail(NRB_ERR_INVALID_DIMS, "plan is NULL");
    return exec_slab_dma(p->plan, isign, d_slab, d_speq, epoch, stream);
} catch (...) { return on_exception(); }
int nrb_slab_dma_timeline(nrb_slab_t p, int enable, char *text, size_t cap)
{
    if (!p) return fail(NRB_ERR_INVALID_DIMS, "plan is NULL");
    if (text && cap) {
        const std::string t = slab_dma_timeline(p->plan);
        strncpy(text, t.c_str(), cap - 1);
        text[cap - 1] = 0;
    }
    p->plan.timeline = enable != 0;
    return NRB_OK;
}
int nrb_slab_barrier_chunk(nrb_slab_t p, int phase, int chunk, unsigned long long epoch, void *stream)
try {
    if (!p) return fail(NRB_ERR_INVALID_DIMS, "plan is NULL");
    return slab_barrier_chunk(p->plan, phase, chunk, epoch, stream);
} catch (...) { return on_exception(); }
int nrb_device_alloc(size_t bytes, void **dptr)
{
    if (!dptr) return fail(NRB_ERR_INVALID_DIMS, "null pointer");
    if (be_device_count() <= 0) return fail(NRB_ERR_CUDA, "no CUDA device available (numrs_b200 has no CPU fallback)");
    if (be_malloc(dptr, bytes) != 0) return fail(NRB_ERR_OOM, std::string("device allocation failed: ") + be_last_error());
    if (be_memset(*dptr, 0, bytes, nullptr) != 0 || be_sync(nullptr) != 0) return fail(NRB_ERR_CUDA, std::string("memset failed: ") + be_last_error());
    return NRB_OK;
}
int nrb_device_free(void *dptr) { if (dptr) be_free(dptr); return NRB_OK; }
int nrb_ipc_export(void *dptr, unsigned char handle[64])
try {
    if (be_ipc_export(dptr, handle) != 0) return fail(NRB_ERR_CUDA, std::string("cudaIpcGetMemHandle failed: ") + be_last_error());
    return NRB_OK;
} catch (...) { return on_exception(); }
int nrb_ipc_import(const unsigned char handle[64], void **dptr)
{
    if (be_ipc_import(handle, dptr) != 0) return fail(NRB_ERR_CUDA, std::string("cudaIpcOpenMemHandle failed: ") + be_last_error());
    return NRB_OK;
}
int nrb_ipc_release(void *dptr)
try {
    if (be_ipc_release(dptr) != 0) return fail(NRB_ERR_CUDA, std::string("cudaIpcCloseMemHandle failed: ") + be_last_error());
    return NRB_OK;
} catch (...) { return on_exception(); }
int nrb_slab_destroy(nrb_slab_t p)
{
    delete p;
    return NRB_OK;
}

// ------------------------------------------------------------------ host-slice entry points
int nrb_four1(double *data, size_t nn, int isign)
try {
    double *ptrs[1] = {data};
    size_t sizes[1] = {nn};
    return nrb_four1_batch(ptrs, sizes, 1, isign);
} catch (...) { return on_exception(); }

int nrb_four1_batch(double *const *ptrs, const size_t *nn, size_t count, int isign)
try {
    if (isign != 1 && isign != -1) return fail(NRB_ERR_INVALID_ISIGN, "isign must be 1 or -1");
    if (count == 0) return NRB_OK;
    if (!ptrs || !nn) return fail(NRB_ERR_EMPTY_INPUT, "null batch");
    // group slices by length; each group is one batched plan
    std::map<size_t, std::vector<double *>> groups;
    for (size_t b = 0; b < count; ++b) {
        if (nn[b] <= 1) continue;                 // FFT_1.rs:5-44 is the identity for nn <= 1
        if (!is_pow2(nn[b])) return fail(NRB_ERR_NOT_POW2, "four1: nn must be a power of two");
        if (!ptrs[b]) return fail(NRB_ERR_EMPTY_INPUT, "null slice");
        groups[nn[b]].push_back(ptrs[b]);
    }
    for (auto &g : groups) {
        const size_t dims[1] = {g.first};
        const int rc = run_inplace_batch(NRB_KIND_FOUR1, dims, 1, g.second.data(), g.second.size(), 2 * g.first, isign);
        if (rc) return rc;
    }
    return NRB_OK;
} catch (...) { return on_exception(); }

int nrb_fourn(double *data, const size_t *nn, size_t ndim, int isign)
try {
    // Fourn.rs:367-378 validation order
    if (ndim == 0 || !nn) return fail(NRB_ERR_INVALID_DIMS, "Invalid dimensions");
    if (isign != 1 && isign != -1) return fail(NRB_ERR_INVALID_ISIGN, "isign must be 1 or -1");
    size_t total = 1;
    for (size_t d = 0; d < ndim; ++d) {
        if (nn[d] <= 1) return fail(NRB_ERR_INVALID_DIMS, "Invalid dimension size");
        if (total > ((size_t)1 << 60) / nn[d]) return fail(NRB_ERR_INVALID_DIMS, "Invalid dimension size");   // product overflows
        total *= nn[d];
    }
    if (!data) return fail(NRB_ERR_EMPTY_INPUT, "null data");
    if (ndim == 3 && multi_device_count() > 1) {
        // slab-decomposed over the devices of the box (multi.cpp); shapes it cannot take fall through to one device
        const int rc = multi_transform3d(false, data, nullptr, nn[0], nn[1], nn[2], isign, multi_device_count());
        if (rc != NRB_ERR_UNSUPPORTED) return rc;
    }
    double *ptrs[1] = {data};
    return run_inplace(NRB_KIND_FOURN, nn, ndim, ptrs, 1, 2 * total, isign, nullptr, 0);
} catch (...) { return on_exception(); }

int nrb_realft(double *data, size_t n, int isign)
try {
    double *ptrs[1] = {data};
    return nrb_realft_batch(ptrs, n, 1, isign);
} catch (...) { return on_exception(); }

int nrb_realft_batch(double *const *ptrs, size_t n, size_t count, int isign)
try {
    if (n % 2 != 0) return fail(NRB_ERR_INVALID_DIMS, "n must be even");          // Real_FT.rs:5
    if (n == 0) return fail(NRB_ERR_EMPTY_INPUT, "data length must be at least n"); // Real_FT.rs:6 / :43
    if (!is_pow2(n)) return fail(NRB_ERR_NOT_POW2, "realft: n must be a power of two");
    if (count == 0) return NRB_OK;
    if (!ptrs) return fail(NRB_ERR_EMPTY_INPUT, "null batch");
    const size_t dims[1] = {n};
    // Real_FT.rs:10,15: isign == 1 is forward, anything else inverse
    return run_inplace_batch(NRB_KIND_REALFT, dims, 1, ptrs, count, n, isign == 1 ? 1 : -1);
} catch (...) { return on_exception(); }

int nrb_rlft3(double *data, double *speq, size_t nn1, size_t nn2, size_t nn3, int isign)
try {
    if (isign != 1 && isign != -1) return fail(NRB_ERR_INVALID_ISIGN, "isign must be 1 or -1");   // Real_FT3.rs:17
    if (nn1 == 0 || nn2 == 0 || nn3 < 2) return fail(NRB_ERR_INVALID_DIMS, "data dimensions mismatch");
    if (!data || !speq) return fail(NRB_ERR_EMPTY_INPUT, "null data");
    if (multi_device_count() > 1) {
        const int rc = multi_transform3d(true, data, speq, nn1, nn2, nn3, isign, multi_device_count());
        if (rc != NRB_ERR_UNSUPPORTED) return rc;
    }
    const size_t dims[3] = {nn1, nn2, nn3};
    double *ptrs[1] = {data};
    return run_inplace(NRB_KIND_RLFT3, dims, 3, ptrs, 1, nn1 * nn2 * nn3, isign, speq, 2 * nn1 * nn2);
} catch (...) { return on_exception(); }

int nrb_convlv(const double *data, size_t n, const double *respns, size_t m, int isign, int pad_mode, double *ans)
try {
    const double *in[1] = {data};
    double *out[1] = {ans};
    return nrb_convlv_batch(in, 1, n, respns, m, isign, pad_mode, out);
} catch (...) { return on_exception(); }

int nrb_convlv_batch(const double *const *data, size_t count, size_t n, const double *respns, size_t m, int isign,
                     int pad_mode, double *const *ans)
try {
    // Convolve.rs:13-21 check order
    if (n == 0 || m == 0) return fail(NRB_ERR_EMPTY_INPUT, "Input arrays cannot be empty");
    if (m > n) return fail(NRB_ERR_RESPONSE_TOO_LONG, "Response function longer than data");
    if (isign != 1 && isign != -1) return fail(NRB_ERR_INVALID_ISIGN, "isign must be 1 (convolution) or -1 (deconvolution)");
    if (n < 2 || !is_pow2(n)) return fail(NRB_ERR_NOT_POW2, "convlv: n must be a power of two >= 2");
    if (pad_mode != NRB_PAD_LITERAL && pad_mode != NRB_PAD_NR) return fail(NRB_ERR_INVALID_DIMS, "unknown pad_mode");
    if (count == 0) return NRB_OK;
    if (!data || !respns || !ans) return fail(NRB_ERR_EMPTY_INPUT, "null pointer");
    const size_t dims[2] = {n, m};
    const double *aux[1] = {respns};
    return run_outofplace_batch(NRB_KIND_CONVLV, dims, 2, data, aux, 1, m, ans, count, n, isign, pad_mode);
} catch (...) { return on_exception(); }

int nrb_correl(const double *data1, size_t n1, const double *data2, size_t n2, double *ans)
try {
    // Correlation.rs:11-16 check order
    if (n1 == 0) return fail(NRB_ERR_EMPTY_INPUT, "Input arrays cannot be empty");
    if (n2 != n1) return fail(NRB_ERR_LENGTH_MISMATCH, "Input arrays must have the same length");
    const double *a[1] = {data1}, *b[1] = {data2};
    double *o[1] = {ans};
    return nrb_correl_batch(a, b, 1, n1, o);
} catch (...) { return on_exception(); }

int nrb_correl_batch(const double *const *data1, const double *const *data2, size_t count, size_t n, double *const *ans)
try {
    if (n == 0) return fail(NRB_ERR_EMPTY_INPUT, "Input arrays cannot be empty");
    if (n > 32 && !is_pow2(n)) return fail(NRB_ERR_NOT_POW2, "correl: n > 32 must be a power of two");
    if (count == 0) return NRB_OK;
    if (!data1 || !data2 || !ans) return fail(NRB_ERR_EMPTY_INPUT, "null pointer");
    const size_t dims[1] = {n};
    return run_outofplace_batch(NRB_KIND_CORREL, dims, 1, data1, data2, count, n, ans, count, n, 1, 0);
} catch (...) { return on_exception(); }

// ------------------------------------------------------------------ "next" rows
int nrb_correl_normalized(const double *data1, size_t n1, const double *data2, size_t n2, int fast, double *ans)
try {
    // Correlation.rs:190-196 / :227-233 check order
    if (n1 == 0) return fail(NRB_ERR_EMPTY_INPUT, "Input arrays cannot be empty");
    if (n2 != n1) return fail(NRB_ERR_LENGTH_MISMATCH, "Input arrays must have the same length");
    if (n1 > 32 && !is_pow2(n1)) return fail(NRB_ERR_NOT_POW2, "correl: n > 32 must be a power of two");
    if (!data1 || !data2 || !ans) return fail(NRB_ERR_EMPTY_INPUT, "null pointer");
    const size_t dims[1] = {n1};
    double stats[4] = {0, 0, 0, 0};   // (mean1, std1, mean2, std2)
    // the answers are only copied back when both standard deviations are non-zero: two-step read-back
    int rc = run_segments(fast ? NRB_KIND_CORREL_NORM_FAST : NRB_KIND_CORREL_NORM, dims, 1, 1, {{data1, 0, n1}}, {{data2, 0, n1}},
                          4 + n1, {{stats, 0, 4}}, 1, 0);
    if (rc) return rc;
    if (stats[1] == 0.0 || stats[3] == 0.0) return fail(NRB_ERR_ZERO_STDDEV, "Zero standard deviation");   // Correlation.rs:214-216
    if (be_d2h(ans, (double *)t_ctx.out.p + 4, n1 * sizeof(double), t_ctx.stream) != 0 || be_sync(t_ctx.stream) != 0)
        return copy_fail("device-to-host copy");
    return NRB_OK;
} catch (...) { return on_exception(); }

int nrb_autocorrel_fast(const double *data, size_t n, double *ans)
try {
    if (n == 0) return fail(NRB_ERR_EMPTY_INPUT, "Input arrays cannot be empty");                 // Correlation.rs:288-290
    if (n > 32 && !is_pow2(n)) return fail(NRB_ERR_NOT_POW2, "correl: n > 32 must be a power of two");
    if (!data || !ans) return fail(NRB_ERR_EMPTY_INPUT, "null pointer");
    const size_t dims[1] = {n};
    return run_segments(NRB_KIND_AUTOCORREL_FAST, dims, 1, 1, {{data, 0, n}}, {}, n, {{ans, 0, n}}, 1, 0);
} catch (...) { return on_exception(); }

int nrb_twofft(const double *data1, const double *data2, size_t n, double *fft1, double *fft2)
{
    const double *a[1] = {data1}, *b[1] = {data2};
    double *f1[1] = {fft1}, *f2[1] = {fft2};
    return nrb_twofft_batch(a, b, 1, n, f1, f2);
}

int nrb_twofft_batch(const double *const *data1, const double *const *data2, size_t count, size_t n, double *const *fft1,
                     double *const *fft2)
try {
    if (n == 0) return fail(NRB_ERR_EMPTY_INPUT, "twofft: empty input");
    if (!is_pow2(n)) return fail(NRB_ERR_NOT_POW2, "twofft: n must be a power of two");
    if (count == 0) return NRB_OK;
    if (!data1 || !data2 || !fft1 || !fft2) return fail(NRB_ERR_EMPTY_INPUT, "null pointer");
    const size_t dims[1] = {n};
    const size_t per = 2 * n + 2;   // FFT_2.rs:6-7
    std::vector<Seg> io, aux, outs;
    for (size_t b = 0; b < count; ++b) {
        if (!data1[b] || !data2[b] || !fft1[b] || !fft2[b]) return fail(NRB_ERR_EMPTY_INPUT, "null pointer");
        io.push_back(Seg{data1[b], b * n, n});
        aux.push_back(Seg{data2[b], b * n, n});
        outs.push_back(Seg{fft1[b], b * per, per});
        outs.push_back(Seg{fft2[b], (count + b) * per, per});
    }
    return run_segments(NRB_KIND_TWOFFT, dims, 1, count, io, aux, 2 * count * per, outs, 1, 0);
} catch (...) { return on_exception(); }

int nrb_power_spectrum(const double *complex_data, size_t npoints, int take_sqrt, double *out)
try {
    if (npoints == 0) return NRB_OK;                           // FFT_1.rs:206-228: empty in, empty out
    if (!complex_data || !out) return fail(NRB_ERR_EMPTY_INPUT, "null pointer");
    const size_t dims[1] = {npoints};
    return run_segments(NRB_KIND_POWER, dims, 1, 1, {{complex_data, 0, 2 * npoints}}, {}, npoints, {{out, 0, npoints}}, 1, take_sqrt ? 1 : 0);
} catch (...) { return on_exception(); }

static int run_cosft(int kind, double *y, size_t n, size_t doubles, int isign)
try {
    if (n < 2) return fail(NRB_ERR_INVALID_DIMS, "cosft/sinft: n must be >= 2");
    if (!is_pow2(n)) return fail(NRB_ERR_NOT_POW2, "cosft/sinft: n must be a power of two");
    if (!y) return fail(NRB_ERR_EMPTY_INPUT, "null data");
    const size_t dims[1] = {n};
    double *ptrs[1] = {y};
    return run_inplace(kind, dims, 1, ptrs, 1, doubles, isign, nullptr, 0);
} catch (...) { return on_exception(); }
int nrb_cosft1(double *y, size_t n) { return run_cosft(NRB_KIND_COSFT1, y, n, n + 2, 1); }
int nrb_cosft2(double *y, size_t n, int isign)
try {
    if (isign != 1 && isign != -1) return fail(NRB_ERR_INVALID_ISIGN, "Invalid isign value. Must be 1 or -1");   // Cos_FT2.rs:11
    return run_cosft(NRB_KIND_COSFT2, y, n, n + 1, isign);
} catch (...) { return on_exception(); }
int nrb_sinft(double *y, size_t n) { return run_cosft(NRB_KIND_SINFT, y, n, n + 1, 1); }

} // extern "C"
