// multi.cpp -- see multi.h: one host-slice call over several GPUs inside one process.
#include "multi.h"

#include <condition_variable>
#include <list>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/numrs_b200.h"

namespace nrb {

int multi_device_count()
{
    const int want = tunables().num_devices;
    if (want == 1) return 1;
    int have = be_device_count();
    if (have < 1) return 1;
    if (want > 1 && want < have) have = want;
    if (have > 8) have = 8;
    int g = 1;
    while (2 * g <= have) g *= 2;
    return g;
}

// ------------------------------------------------------------------ worker pool: one persistent thread per device
namespace {

struct Worker {
    std::mutex mu;
    std::condition_variable cv;
    const std::function<int(int)> *job;
    bool busy;
    int rc;
    std::string err;
    Worker() : job(nullptr), busy(false), rc(0) {}
};

std::mutex g_pool_mu;                       // one multi-device call at a time: the workers are shared
std::vector<Worker *> &pool() { static std::vector<Worker *> *p = new std::vector<Worker *>(); return *p; }   // never destroyed

void worker_main(Worker *w, int dev)
{
    for (;;) {
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv.wait(lk, [w] { return w->job != nullptr; });
        const std::function<int(int)> *job = w->job;
        lk.unlock();
        int rc;
        std::string err;
        try {
            if (be_set_device(dev) != 0) { rc = NRB_ERR_CUDA; err = std::string("cudaSetDevice failed: ") + be_last_error(); }
            else {
                rc = (*job)(dev);
                if (rc != 0) err = get_error();
            }
        } catch (const std::bad_alloc &) { rc = NRB_ERR_OOM; err = "out of host memory"; }
        catch (...) { rc = NRB_ERR_CUDA; err = "internal error in a device worker"; }
        lk.lock();
        w->rc = rc;
        w->err = err;
        w->job = nullptr;
        w->busy = false;
        lk.unlock();
        w->cv.notify_all();
    }
}

} // namespace

int multi_run(int G, const std::function<int(int)> &fn)
{
    std::vector<Worker *> &P = pool();
    while ((int)P.size() < G) {
        Worker *w = new Worker();
        std::thread(worker_main, w, (int)P.size()).detach();
        P.push_back(w);
    }
    for (int g = 0; g < G; ++g) {
        std::lock_guard<std::mutex> lk(P[g]->mu);
        P[g]->job = &fn;
        P[g]->busy = true;
        P[g]->cv.notify_all();
    }
    int rc = NRB_OK;
    for (int g = 0; g < G; ++g) {
        std::unique_lock<std::mutex> lk(P[g]->mu);
        P[g]->cv.wait(lk, [&] { return !P[g]->busy; });
        if (P[g]->rc != 0 && rc == NRB_OK) { rc = P[g]->rc; set_error("device " + std::to_string(g) + ": " + P[g]->err); }
    }
    return rc;
}

// ------------------------------------------------------------------ 3-D transforms: slabs over G devices
namespace {

struct SlabDev {
    SlabPlan sp;
    void *stream, *slab, *speq, *recv, *send, *ev0;
    SlabDev() : stream(nullptr), slab(nullptr), speq(nullptr), recv(nullptr), send(nullptr), ev0(nullptr) {}
};

struct MultiSlab {
    bool real;
    size_t nn1, nn2, nn3;
    int G;
    SlabDev d[8];
    MultiSlab() : real(true), nn1(0), nn2(0), nn3(0), G(0) {}
    size_t local_bytes() const { return (real ? 1 : 2) * nn1 * nn2 * nn3 / (size_t)G * sizeof(double); }
    size_t speq_bytes() const { return real ? 2 * nn1 * nn2 / (size_t)G * sizeof(double) : 0; }
    // (from the shape, not from d[0].sp: the devices build their plans concurrently)
    size_t recv_bytes() const
    {
        const size_t g = (size_t)G, blk = (nn1 / g) * (nn2 / g) * ((real ? nn3 / 2 : nn3) + (real ? 1 : 0));
        return g * blk * sizeof(double2) + 8 * 8 * kSlabMaxChunks;
    }
    void release()
    {
        // every buffer is freed with its own device current (streams and events belong to a device)
        const int cur = be_current_device();
        for (int g = 0; g < G; ++g) {
            if (be_set_device(g) != 0) continue;
            SlabDev &D = d[g];
            if (D.stream) be_sync(D.stream);
            slab_release(D.sp);
            D.sp.tables.clear();
            if (D.slab) be_free(D.slab);
            if (D.speq) be_free(D.speq);
            if (D.recv) be_free(D.recv);
            if (D.send) be_free(D.send);
            if (D.ev0) be_event_destroy(D.ev0);
            if (D.stream) be_stream_destroy(D.stream);
            D.slab = D.speq = D.recv = D.send = D.ev0 = D.stream = nullptr;
        }
        if (cur >= 0) be_set_device(cur);
        G = 0;
    }
    ~MultiSlab() { release(); }
};

// a multi-device plan pins (slab + receive buffer) on every device, so only the two most recent shapes are kept
std::list<std::shared_ptr<MultiSlab>> &slab_cache() { static auto *c = new std::list<std::shared_ptr<MultiSlab>>(); return *c; }

std::shared_ptr<MultiSlab> get_multi_slab(bool real, size_t nn1, size_t nn2, size_t nn3, int G, int *rc)
{
    auto &C = slab_cache();
    for (auto it = C.begin(); it != C.end(); ++it) {
        if ((*it)->real == real && (*it)->nn1 == nn1 && (*it)->nn2 == nn2 && (*it)->nn3 == nn3 && (*it)->G == G) {
            std::shared_ptr<MultiSlab> m = *it;
            C.erase(it);
            C.push_front(m);
            *rc = NRB_OK;
            return m;
        }
    }
    while (C.size() >= 2) C.pop_back();
    std::shared_ptr<MultiSlab> m(new MultiSlab());
    m->real = real; m->nn1 = nn1; m->nn2 = nn2; m->nn3 = nn3; m->G = G;
    MultiSlab *M = m.get();
    *rc = multi_run(G, [M](int g) -> int {
        SlabDev &D = M->d[g];
        for (int p = 0; p < M->G; ++p)
            if (be_enable_peer(g, p) != 0) { set_error(std::string("peer access: ") + be_last_error()); return NRB_ERR_UNSUPPORTED; }
        int r = build_slab_plan(D.sp, M->nn1, M->nn2, M->nn3, M->G, g, M->real);
        if (r != NRB_OK) return r;
        if (be_stream_create(&D.stream) != 0) { set_error(std::string("stream creation failed: ") + be_last_error()); return NRB_ERR_CUDA; }
        D.ev0 = be_event_create();
        if (!D.ev0) { set_error("event creation failed"); return NRB_ERR_CUDA; }
        if (be_malloc(&D.slab, M->local_bytes()) != 0 || (M->real && be_malloc(&D.speq, M->speq_bytes()) != 0) ||
            be_malloc(&D.recv, M->recv_bytes()) != 0 || be_malloc(&D.send, M->recv_bytes()) != 0) {
            set_error(std::string("device allocation failed: ") + be_last_error());
            return NRB_ERR_OOM;
        }
        if (be_memset(D.recv, 0, M->recv_bytes(), D.stream) != 0 || be_sync(D.stream) != 0) { set_error(std::string("memset failed: ") + be_last_error()); return NRB_ERR_CUDA; }
        return NRB_OK;
    });
    if (*rc != NRB_OK) return nullptr;     // ~MultiSlab frees what was allocated
    // push + pull exchange (plan.cpp exec_slab_stage): every device's receive AND send buffer is visible to all
    void *peers[8], *sends[8];
    for (int g = 0; g < G; ++g) { peers[g] = M->d[g].recv; sends[g] = M->d[g].send; }
    for (int g = 0; g < G; ++g)
        if ((*rc = slab_set_peers(M->d[g].sp, peers, G)) != NRB_OK || (*rc = slab_set_send_peers(M->d[g].sp, sends, G)) != NRB_OK) return nullptr;
    C.push_front(m);
    return m;
}

} // namespace

static long g_multi_calls[2] = {0, 0};
long multi_calls(int which) { std::lock_guard<std::mutex> lk(g_pool_mu); return g_multi_calls[which ? 1 : 0]; }

int multi_shard_batch(size_t total, int G, const std::function<int(size_t, size_t)> &fn)
{
    std::lock_guard<std::mutex> lk(g_pool_mu);
    ++g_multi_calls[1];
    const size_t per = (total + (size_t)G - 1) / (size_t)G;
    return multi_run(G, [&](int g) -> int {
        const size_t first = (size_t)g * per;
        if (first >= total) return NRB_OK;
        return fn(first, total - first < per ? total - first : per);
    });
}

void multi_release()
{
    std::lock_guard<std::mutex> lk(g_pool_mu);
    slab_cache().clear();
}

int multi_transform3d(bool real, double *data, double *speq, size_t nn1, size_t nn2, size_t nn3, int isign, int G)
{
    if (G < 2 || !is_pow2(nn1) || !is_pow2(nn2) || !is_pow2(nn3) || (size_t)G > nn1 || (size_t)G > nn2 || nn3 < 2) return NRB_ERR_UNSUPPORTED;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    int rc = NRB_OK;
    std::shared_ptr<MultiSlab> m = get_multi_slab(real, nn1, nn2, nn3, G, &rc);
    if (!m) return rc;
    MultiSlab *M = m.get();
    ++g_multi_calls[0];
    const size_t X = nn1 / (size_t)G, Y = nn2 / (size_t)G;
    const size_t zl = (real ? 1 : 2) * nn3;                  // doubles per z line
    const size_t row_bytes = Y * zl * sizeof(double);        // one x-row of an nn2-slab
    const size_t host_pitch = nn2 * zl * sizeof(double);
    const bool fwd = isign == 1;
    // phase A: scatter the input slabs (all PCIe links at once), stage 0 (its last pass stores into the peers' receive
    // buffers over NVLink), one event per device
    rc = multi_run(G, [&](int g) -> int {
        SlabDev &D = M->d[g];
        int e;
        if (fwd) e = be_h2d_2d(D.slab, row_bytes, data + (size_t)g * Y * zl, host_pitch, row_bytes, nn1, D.stream);
        else {
            e = be_h2d(D.slab, data + (size_t)g * X * nn2 * zl, M->local_bytes(), D.stream);
            if (e == 0 && real) e = be_h2d(D.speq, speq + (size_t)g * X * 2 * nn2, M->speq_bytes(), D.stream);
        }
        if (e != 0) { set_error(std::string("host-to-device copy failed: ") + be_last_error()); return NRB_ERR_CUDA; }
        const int r = exec_slab_stage(D.sp, 0, isign, (double *)D.slab, (double *)D.speq, nullptr, nullptr, D.stream);
        if (r != NRB_OK) return r;
        if (be_event_record_on(D.ev0, D.stream) != 0) { set_error(std::string("event record failed: ") + be_last_error()); return NRB_ERR_CUDA; }
        return NRB_OK;
    });
    // (the return of multi_run is the host-side barrier: every device's event has been recorded before anyone waits on it)
    // phase B: stage 1 behind every device's stage 0, gather the output slabs
    const int rcb = multi_run(G, [&](int g) -> int {
        SlabDev &D = M->d[g];
        if (rc != NRB_OK) { be_sync(D.stream); return NRB_OK; }     // a peer failed: just drain
        for (int p = 0; p < G; ++p)
            if (be_stream_wait(D.stream, M->d[p].ev0) != 0) { set_error(std::string("stream wait failed: ") + be_last_error()); return NRB_ERR_CUDA; }
        const int r = exec_slab_stage(D.sp, 1, isign, (double *)D.slab, (double *)D.speq, nullptr, nullptr, D.stream);
        if (r != NRB_OK) { be_sync(D.stream); return r; }
        int e;
        if (fwd) {
            e = be_d2h(data + (size_t)g * X * nn2 * zl, D.slab, M->local_bytes(), D.stream);
            if (e == 0 && real) e = be_d2h(speq + (size_t)g * X * 2 * nn2, D.speq, M->speq_bytes(), D.stream);
        } else {
            e = be_d2h_2d(data + (size_t)g * Y * zl, host_pitch, D.slab, row_bytes, row_bytes, nn1, D.stream);
        }
        if (e != 0) { set_error(std::string("device-to-host copy failed: ") + be_last_error()); be_sync(D.stream); return NRB_ERR_CUDA; }
        if (be_sync(D.stream) != 0) { set_error(std::string("kernel execution failed: ") + be_last_error()); return NRB_ERR_CUDA; }
        return NRB_OK;
    });
    return rc != NRB_OK ? rc : rcb;
}

} // namespace nrb
