// cabi_bench.cpp -- per-kernel timing of one workload through the C ABI only (no Python, no torch: a gpurun call
// with this binary costs seconds).  Usage:
//   cabi_bench <libnumrs_b200.so> <workload> [option=value ...]
//   workload: four1:<log2n>:<batch> | fourn:<n0>x<n1>[x<n2>] | rlft3:<n> | convlv:<log2n>:<batch> | correl:<log2n>:<batch>
//             | autocorrel:<log2n>:<batch>      (out of place, 4096-tap response, one direction; prints a checksum of the
//             first 2^20 outputs instead of a round trip, so that two runs with different options can be compared)
// Prints, for the forward + inverse pair: every launch's kernel name, algorithmic bytes, average ms (CUDA events
// around the launch, nrb_plan_profile) and GB/s; the pair's wall time over back-to-back executions; and the
// round-trip error inverse(forward(x)) / scale vs x on the first 2^20 doubles.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <string>
#include <vector>

#include "../include/numrs_b200.h"

#define SYM(name) decltype(&::name) p_##name = (decltype(&::name))dlsym(h, #name); if (!p_##name) { fprintf(stderr, "missing %s\n", #name); return 2; }

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s lib workload [opt=val ...]\n", argv[0]); return 2; }
    void *h = dlopen(argv[1], RTLD_NOW);
    if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    SYM(nrb_set_option) SYM(nrb_plan_create) SYM(nrb_plan_exec) SYM(nrb_plan_profile) SYM(nrb_plan_describe_launch)
    SYM(nrb_plan_num_launches) SYM(nrb_plan_destroy) SYM(nrb_device_alloc) SYM(nrb_device_free) SYM(nrb_fill_uniform_device)
    SYM(nrb_download) SYM(nrb_stream_synchronize) SYM(nrb_last_error) SYM(nrb_device_count)
    if (p_nrb_device_count() <= 0) { fprintf(stderr, "no CUDA device\n"); return 3; }
    for (int i = 3; i < argc; ++i) {
        char *eq = strchr(argv[i], '=');
        if (!eq) continue;
        *eq = 0;
        if (p_nrb_set_option(argv[i], atol(eq + 1)) != 0) { fprintf(stderr, "option %s: %s\n", argv[i], p_nrb_last_error()); return 2; }
        printf("option %s = %ld\n", argv[i], atol(eq + 1));
    }
    std::string wl = argv[2];
    int kind = 0;
    std::vector<size_t> dims;
    size_t batch = 1, io_doubles = 0, aux_doubles = 0, out_doubles = 0;
    bool oop = false;   // out of place, forward only
    double scale = 1.0, alg_bytes = 0.0;
    if (wl.rfind("four1:", 0) == 0) {
        int lg = 0; unsigned long b = 1;
        sscanf(wl.c_str(), "four1:%d:%lu", &lg, &b);
        kind = NRB_KIND_FOUR1; dims = {(size_t)1 << lg}; batch = b;
        io_doubles = 2 * dims[0] * batch; scale = (double)dims[0]; alg_bytes = 32.0 * dims[0] * batch;
    } else if (wl.rfind("fourn:", 0) == 0) {
        unsigned long a = 1, b = 1, c = 0;
        const int n = sscanf(wl.c_str(), "fourn:%lux%lux%lu", &a, &b, &c);
        kind = NRB_KIND_FOURN; dims = {a, b}; if (n == 3) dims.push_back(c);
        size_t tot = 1; for (size_t d : dims) tot *= d;
        io_doubles = 2 * tot; scale = (double)tot; alg_bytes = 32.0 * tot;
    } else if (wl.rfind("rlft3:", 0) == 0) {
        unsigned long n = 0;
        sscanf(wl.c_str(), "rlft3:%lu", &n);
        kind = NRB_KIND_RLFT3; dims = {n, n, n};
        io_doubles = n * n * n; aux_doubles = 2 * n * n; scale = (double)(n * n * n) / 2.0; alg_bytes = 16.0 * n * n * n + 16.0 * n * n;
    } else if (wl.rfind("convlv:", 0) == 0 || wl.rfind("correl:", 0) == 0 || wl.rfind("autocorrel:", 0) == 0) {
        int lg = 0; unsigned long b = 1;
        sscanf(wl.c_str() + wl.find(':') + 1, "%d:%lu", &lg, &b);
        const size_t n = (size_t)1 << lg;
        batch = b; oop = true;
        io_doubles = n * batch; out_doubles = n * batch;
        if (wl[0] == 'c' && wl[2] == 'n') { kind = NRB_KIND_CONVLV; dims = {n, 4096}; aux_doubles = 4096; alg_bytes = 16.0 * n * batch + 8.0 * n; }
        else if (wl[0] == 'c') { kind = NRB_KIND_CORREL; dims = {n}; aux_doubles = n * batch; alg_bytes = 24.0 * n * batch; }
        else { kind = NRB_KIND_AUTOCORREL_FAST; dims = {n}; alg_bytes = 16.0 * n * batch; }
    } else { fprintf(stderr, "unknown workload\n"); return 2; }

    nrb_plan_t plan = nullptr;
    if (p_nrb_plan_create(kind, dims.data(), dims.size(), batch, &plan) != 0) { fprintf(stderr, "plan: %s\n", p_nrb_last_error()); return 1; }
    const int NBUF = 3;
    double *buf[NBUF], *aux = nullptr;
    for (int i = 0; i < NBUF; ++i) {
        if (p_nrb_device_alloc(io_doubles * 8, (void **)&buf[i]) != 0) { fprintf(stderr, "alloc: %s\n", p_nrb_last_error()); return 1; }
        p_nrb_fill_uniform_device(buf[i], 4242, 0, io_doubles, nullptr);
    }
    if (aux_doubles) {
        p_nrb_device_alloc(aux_doubles * 8, (void **)&aux);
        p_nrb_fill_uniform_device(aux, 4243, 0, aux_doubles, nullptr);
    }
    double *outb = nullptr;
    if (out_doubles) p_nrb_device_alloc(out_doubles * 8, (void **)&outb);
    p_nrb_stream_synchronize(nullptr);
    if (oop) {
        if (p_nrb_plan_exec(plan, buf[0], aux, outb, 1, 0, nullptr) != 0 || p_nrb_stream_synchronize(nullptr) != 0) { fprintf(stderr, "exec: %s\n", p_nrb_last_error()); return 1; }
        const size_t ns = out_doubles < (1u << 20) ? out_doubles : (1u << 20);
        std::vector<double> y(ns);
        p_nrb_download(y.data(), outb + (out_doubles - ns), ns * 8, nullptr);
        p_nrb_stream_synchronize(nullptr);
        double sum = 0, sq = 0;
        for (double t : y) { sum += t; sq += t * t; }
        printf("workload %s: checksum of the last %zu outputs: sum %.15e  sumsq %.15e\n", wl.c_str(), ns, sum, sq);
        const int REPS = 6;
        std::map<std::string, std::pair<double, double>> agg;
        std::vector<std::string> order;
        for (int rep = 0; rep < REPS; ++rep) {
            float ms[256];
            const int n = p_nrb_plan_num_launches(plan, 1);
            if (p_nrb_plan_profile(plan, buf[rep % NBUF], aux, outb, 1, 0, nullptr, ms, 256) != 0) { fprintf(stderr, "profile: %s\n", p_nrb_last_error()); return 1; }
            if (rep == 0) continue;
            for (int i = 0; i < n && i < 256; ++i) {
                char name[128]; double bytes = 0;
                p_nrb_plan_describe_launch(plan, 1, i, name, sizeof(name), &bytes);
                std::string key = std::string(name) + " #" + std::to_string(i);
                if (!agg.count(key)) order.push_back(key);
                agg[key].first += bytes; agg[key].second += ms[i];
            }
        }
        for (const std::string &k : order)
            printf("  %-48s %8.3f ms  %7.0f GB/s\n", k.c_str(), agg[k].second / (REPS - 1), agg[k].first / agg[k].second / 1e6);
        const int STEPS = 12;
        for (int i = 0; i < 3; ++i) p_nrb_plan_exec(plan, buf[i % NBUF], aux, outb, 1, 0, nullptr);
        p_nrb_stream_synchronize(nullptr);
        const auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < STEPS; ++i) p_nrb_plan_exec(plan, buf[i % NBUF], aux, outb, 1, 0, nullptr);
        p_nrb_stream_synchronize(nullptr);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / STEPS;
        printf("  one call: %.3f ms, %.0f GB/s algorithmic (wall clock over %d back-to-back calls)\n", ms, alg_bytes / ms / 1e6, STEPS);
        p_nrb_plan_destroy(plan);
        return 0;
    }

    // correctness first: round trip on buffer 0
    const size_t ns = io_doubles < (1u << 20) ? io_doubles : (1u << 20);
    std::vector<double> x0(ns), x1(ns);
    p_nrb_download(x0.data(), buf[0], ns * 8, nullptr);
    p_nrb_stream_synchronize(nullptr);
    int rc = p_nrb_plan_exec(plan, buf[0], aux, nullptr, 1, 0, nullptr);
    if (rc == 0) rc = p_nrb_plan_exec(plan, buf[0], aux, nullptr, -1, 0, nullptr);
    if (rc != 0 || p_nrb_stream_synchronize(nullptr) != 0) { fprintf(stderr, "exec: %s\n", p_nrb_last_error()); return 1; }
    p_nrb_download(x1.data(), buf[0], ns * 8, nullptr);
    p_nrb_stream_synchronize(nullptr);
    double num = 0, den = 0;
    for (size_t i = 0; i < ns; ++i) { const double d = x1[i] / scale - x0[i]; num += d * d; den += x0[i] * x0[i]; }
    printf("workload %s: round-trip rel-L2 error %.3e (first %zu doubles)\n", wl.c_str(), std::sqrt(num / den), ns);
    p_nrb_fill_uniform_device(buf[0], 4242, 0, io_doubles, nullptr);

    // per-launch profile
    const int REPS = 6;
    std::map<std::string, std::pair<double, double>> agg;   // name -> (bytes, ms) summed
    std::vector<std::string> order;
    for (int rep = 0; rep < REPS; ++rep) {
        for (int isign = 1; isign >= -1; isign -= 2) {
            float ms[256];
            const int n = p_nrb_plan_num_launches(plan, isign);
            if (p_nrb_plan_profile(plan, buf[rep % NBUF], aux, nullptr, isign, 0, nullptr, ms, 256) != 0) { fprintf(stderr, "profile: %s\n", p_nrb_last_error()); return 1; }
            if (rep == 0) continue;   // warm-up
            for (int i = 0; i < n && i < 256; ++i) {
                char name[128]; double bytes = 0;
                p_nrb_plan_describe_launch(plan, isign, i, name, sizeof(name), &bytes);
                std::string key = std::string(name) + (isign > 0 ? " fwd#" : " inv#") + std::to_string(i);
                if (!agg.count(key)) order.push_back(key);
                agg[key].first += bytes; agg[key].second += ms[i];
            }
        }
    }
    for (const std::string &k : order)
        printf("  %-48s %8.3f ms  %7.0f GB/s\n", k.c_str(), agg[k].second / (REPS - 1), agg[k].first / agg[k].second / 1e6);

    // wall time of forward + inverse, back to back, rotating buffers
    const int STEPS = 12;
    for (int i = 0; i < 3; ++i) { p_nrb_plan_exec(plan, buf[i % NBUF], aux, nullptr, 1, 0, nullptr); p_nrb_plan_exec(plan, buf[i % NBUF], aux, nullptr, -1, 0, nullptr); }
    p_nrb_stream_synchronize(nullptr);
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < STEPS; ++i) { p_nrb_plan_exec(plan, buf[i % NBUF], aux, nullptr, 1, 0, nullptr); p_nrb_plan_exec(plan, buf[i % NBUF], aux, nullptr, -1, 0, nullptr); }
    p_nrb_stream_synchronize(nullptr);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / STEPS;
    printf("  forward + inverse: %.3f ms per pair, %.0f GB/s algorithmic (wall clock over %d back-to-back pairs)\n", ms, 2 * alg_bytes / ms / 1e6, STEPS);
    p_nrb_plan_destroy(plan);
    return 0;
}
