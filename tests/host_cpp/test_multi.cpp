// test_multi.cpp -- drives the multi-device host-slice path (numrs_b200/csrc/multi.cpp) from C++ through the host mirror
// of the reference interface, the way a Rust caller of `rlft3(&mut data, &mut speq, ..)` (Real_FT3.rs:8), the in-memory
// `Fourn` (Real_FT3.rs:35), `fft_batch` (FFT_1.rs:185), `convlv_batch` (Convolve.rs:241) and `correl_batch`
// (Correlation.rs:273) would: whole host arrays in, whole host arrays out, no torch, no IPC, one process.
// Every call is made twice -- option num_devices = 1 (the calling thread's device) and = 0 (every visible device) -- and
// the results must agree to rounding (the slab path applies the x and y passes in the other order).
// Linked against libnumrs_b200.so on a GPU box, or against the test-only emulation (NRB_EMU_DEVICES=4) in the CPU tier.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../numrs_b200/host/num_rs.hpp"

using namespace num_rs;

static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); ++failures; } } while (0)

static double rel(const std::vector<double> &a, const std::vector<double> &b)
{
    double num = 0, den = 0;
    for (std::size_t i = 0; i < a.size(); ++i) { num += (a[i] - b[i]) * (a[i] - b[i]); den += b[i] * b[i]; }
    return std::sqrt(num / den);
}
static std::vector<double> noise(std::size_t n, unsigned seed)
{
    std::vector<double> v(n);
    unsigned long long s = 0x9E3779B97F4A7C15ull * (seed + 1);
    for (std::size_t i = 0; i < n; ++i) { s = s * 6364136223846793005ull + 1442695040888963407ull; v[i] = (double)(s >> 11) / 9007199254740992.0 * 2.0 - 1.0; }
    return v;
}

int main(int argc, char **argv)
{
    const std::size_t n1 = argc > 1 ? std::strtoul(argv[1], nullptr, 10) : 16, n2 = argc > 2 ? std::strtoul(argv[2], nullptr, 10) : 32,
                      n3 = argc > 3 ? std::strtoul(argv[3], nullptr, 10) : 8;
    CHECK(nrb_set_option("num_devices", 0) == NRB_OK);
    const int G = nrb_num_devices_in_use();
    std::printf("devices visible %d, used per call %d\n", nrb_device_count(), G);
    CHECK(nrb_set_option("shard_min_kb", 0) == NRB_OK);
    const double tol = 1e-12 * std::log2((double)(n1 * n2 * n3));
    const long t0 = nrb_multi_device_calls(0), b0 = nrb_multi_device_calls(1);
    // ---- rlft3: forward spectrum + speq plane, then the inverse
    {
        const std::vector<double> x = noise(n1 * n2 * n3, 1);
        std::vector<double> d1 = x, s1(2 * n1 * n2), dG = x, sG(2 * n1 * n2);
        nrb_set_option("num_devices", 1);
        Real_FT3::rlft3(d1, s1, n1, n2, n3, 1);
        nrb_set_option("num_devices", 0);
        Real_FT3::rlft3(dG, sG, n1, n2, n3, 1);
        CHECK(rel(dG, d1) <= tol && rel(sG, s1) <= tol);
        Real_FT3::rlft3(dG, sG, n1, n2, n3, -1);
        for (double &v : dG) v *= 2.0 / (double)(n1 * n2 * n3);
        CHECK(rel(dG, x) <= tol);
    }
    // ---- 3-D complex fourn, both signs
    for (int isign = 1; isign >= -1; isign -= 2) {
        const std::vector<double> z = noise(2 * n1 * n2 * n3, 2);
        std::vector<double> a = z, b = z;
        nrb_set_option("num_devices", 1);
        Fourn::fourn(a, {n1, n2, n3}, 3, isign);
        nrb_set_option("num_devices", 0);
        Fourn::fourn(b, {n1, n2, n3}, 3, isign);
        CHECK(rel(b, a) <= tol);
    }
    if (G > 1 && (std::size_t)G <= n1 && (std::size_t)G <= n2) CHECK(nrb_multi_device_calls(0) - t0 == 4);
    // ---- batches: fft_batch, convlv_batch, correl_batch (ragged: 2 G + 1 signals)
    {
        const std::size_t cnt = 2 * (std::size_t)G + 1, nn = 256, n = 512;
        std::vector<std::vector<double>> a(cnt), b(cnt);
        std::vector<std::pair<double *, std::size_t>> pa, pb;
        for (std::size_t i = 0; i < cnt; ++i) { a[i] = noise(2 * nn, 10 + (unsigned)i); b[i] = a[i]; }
        for (std::size_t i = 0; i < cnt; ++i) { pa.push_back({a[i].data(), a[i].size()}); pb.push_back({b[i].data(), b[i].size()}); }
        nrb_set_option("num_devices", 1);
        FFT_1::FFTProcessor().fft_batch(pa, 1);
        nrb_set_option("num_devices", 0);
        FFT_1::FFTProcessor().fft_batch(pb, 1);
        for (std::size_t i = 0; i < cnt; ++i) CHECK(a[i] == b[i]);       // same kernels on every device: bit-identical
        std::vector<std::vector<double>> sig(cnt);
        std::vector<std::pair<std::vector<double>, std::vector<double>>> tm(cnt);
        for (std::size_t i = 0; i < cnt; ++i) { sig[i] = noise(n, 40 + (unsigned)i); tm[i] = {sig[i], noise(n, 80 + (unsigned)i)}; }
        const std::vector<double> resp = noise(9, 99);
        nrb_set_option("num_devices", 1);
        auto c1 = Convolve::convlv_batch(sig, resp, 1);
        auto r1 = Correlation::correl_batch(tm);
        nrb_set_option("num_devices", 0);
        auto cG = Convolve::convlv_batch(sig, resp, 1);
        auto rG = Correlation::correl_batch(tm);
        for (std::size_t i = 0; i < cnt; ++i) CHECK(c1[i] == cG[i] && r1[i] == rG[i]);
    }
    if (G > 1) CHECK(nrb_multi_device_calls(1) - b0 == 3);
    nrb_shutdown();
    std::printf(failures ? "multi-device host calls: %d failures\n" : "multi-device host calls: ok (%d device(s))\n", failures ? failures : G);
    return failures ? 1 : 0;
}
