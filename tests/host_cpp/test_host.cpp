// test_host.cpp -- exercises the C++ host mirror (numrs_b200/host/num_rs.hpp) the way the
// reference's own unit tests exercise the Rust API.  argv[1] == "gpu": run the numeric tests
// (needs a device); otherwise only the argument-validation paths, which must behave like the
// reference before any device is touched.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "../../numrs_b200/host/num_rs.hpp"

using namespace num_rs;

static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); ++failures; } } while (0)

template <class E, class F> static bool throws(F f)
{
    try { f(); } catch (const E &) { return true; } catch (...) { return false; }
    return false;
}

int main(int argc, char **argv)
{
    const bool gpu = argc > 1 && !std::strcmp(argv[1], "gpu");
    // Convolve.rs:406-424
    try { Convolve::convlv({}, {1.0}, 1); CHECK(false); } catch (const Convolve::ConvlvError &e) { CHECK(e.kind == Convolve::ConvlvError::EmptyInput); }
    try { Convolve::convlv({1.0, 2.0}, {1.0, 2.0, 3.0}, 1); CHECK(false); } catch (const Convolve::ConvlvError &e) { CHECK(e.kind == Convolve::ConvlvError::ResponseTooLong); }
    try { Convolve::convlv({1.0, 2.0}, {1.0}, 0); CHECK(false); } catch (const Convolve::ConvlvError &e) { CHECK(e.kind == Convolve::ConvlvError::InvalidIsign); }
    // Correlation.rs:453-462
    try { Correlation::correl({}, {1.0}); CHECK(false); } catch (const Correlation::CorrelError &e) { CHECK(e.kind == Correlation::CorrelError::EmptyInput); }
    try { Correlation::correl({1.0, 2.0}, {1.0}); CHECK(false); } catch (const Correlation::CorrelError &e) { CHECK(e.kind == Correlation::CorrelError::LengthMismatch); }
    // Fourn.rs:467-476
    { std::vector<double> d(2); CHECK(throws<Fourn::InvalidInput>([&] { Fourn::fourn(d, {1}, 1, 1); })); }
    { std::vector<double> d(16); CHECK(throws<Fourn::InvalidInput>([&] { Fourn::fourn(d, {8}, 1, 0); })); }
    // Real_FT.rs:5-6, Real_FT3.rs:17-19
    { std::vector<double> d(6); CHECK(throws<Panic>([&] { Real_FT::realft(d, 5, 1); })); }
    { std::vector<double> a(4), b(4), f(8), g(10); CHECK(throws<Panic>([&] { FFT_2::twofft(a, b, f, g); })); }   // FFT_2.rs:6
    { std::vector<double> y(9); CHECK(throws<Panic>([&] { Cos_FT2::cosft2(y, 8, 0); })); }                        // Cos_FT2.rs:266-272
    { std::vector<double> d(64), s(32); CHECK(throws<Panic>([&] { Real_FT3::rlft3(d, s, 4, 4, 4, 0); })); }
    { std::vector<double> d(64), s(16); CHECK(throws<Panic>([&] { Real_FT3::rlft3(d, s, 4, 4, 4, 1); })); }
    if (gpu) {
        // FFT_1.rs:246-267
        const std::size_t n = 1024;
        std::vector<double> sig(n);
        for (std::size_t i = 0; i < n; ++i) { double t = (double)i / n; sig[i] = std::sin(2 * M_PI * 5 * t) + 0.5 * std::cos(2 * M_PI * 20 * t); }
        auto c = FFT_1::real_to_complex(sig);
        FFT_1::four1(c, n, 1);
        FFT_1::four1(c, n, -1);
        for (std::size_t i = 0; i < n; ++i) CHECK(std::fabs(c[2 * i] / n - sig[i]) < 1e-10);
        // Convolve.rs:347-360 (indices 1..3), :445-454
        auto y = Convolve::convlv({1, 2, 3, 4}, {1, 1}, 1);
        CHECK(std::fabs(y[1] - 3) < 1e-10 && std::fabs(y[2] - 5) < 1e-10 && std::fabs(y[3] - 7) < 1e-10);
        CHECK(std::fabs(Convolve::ConvlvProcessor().process({1, 2, 3, 4}, {1, 1}, 1)[1] - 3.0) < 1e-10);
        // Correlation.rs:407-418, :494-502
        auto r = Correlation::correl({1, 2, 3, 4}, {1, 2, 3, 4});
        CHECK(r[0] == 30.0 && r[0] > r[1]);
        auto r2 = Correlation::correl({1, 2}, {1, 2});
        CHECK(r2[0] == 5.0 && r2[1] == 2.0);
        // Correlation.rs:505-512, ZeroStdDev :214-216; FFT_2.rs:406-428 (fft1[1] == 0, Hermitian symmetry)
        CHECK(std::fabs(Correlation::correl_normalized_fast({1, 2, 3, 4}, {1, 2, 3, 4})[0] - 1.0) < 1e-10);
        try { Correlation::correl_normalized({1, 1, 1, 1}, {1, 2, 3, 4}); CHECK(false); } catch (const Correlation::CorrelError &e) { CHECK(e.kind == Correlation::CorrelError::ZeroStdDev); }
        CHECK(Correlation::autocorrel_fast({1, 2, 1, 2})[0] == 10.0);
        {
            const std::size_t m = 256;
            std::vector<double> a(m), b(m), f1(2 * m + 2), f2(2 * m + 2);
            for (std::size_t i = 0; i < m; ++i) { double t = (double)i / m; a[i] = std::sin(2 * M_PI * 5 * t); b[i] = std::cos(2 * M_PI * 10 * t); }
            FFT_2::twofft(a, b, f1, f2);
            CHECK(f1[1] == 0.0 && f2[1] == 0.0 && std::fabs(f2[20] - m / 2.0) < 1e-9);
            for (std::size_t k = 1; k < m / 2; ++k) CHECK(std::fabs(f1[2 * k] - f1[2 * (m - k)]) < 1e-10 && std::fabs(f1[2 * k + 1] + f1[2 * (m - k) + 1]) < 1e-10);
            auto pw = FFT_1::power_spectrum_device(f2), hp = FFT_1::power_spectrum(f2);
            for (std::size_t k = 0; k < m; ++k) CHECK(std::fabs(pw[k] - hp[k]) <= 4e-16 * hp[k]);   // the device contracts x*x + y*y into an FMA
        }
        {   // FFT_2.rs:258 process_batch = twofft on every tuple; Real_FT.rs:365 process_batch = realft on every item
            std::vector<double> a1(64), b1(64), a2(256), b2(256), f1(130), g1(130), f2(514), g2(514), r1(130), s1(130);
            for (std::size_t i = 0; i < 256; ++i) { a2[i] = std::sin(0.1 * i); b2[i] = std::cos(0.3 * i); if (i < 64) { a1[i] = a2[i] * 2; b1[i] = b2[i] - 1; } }
            FFT_2::TwoFFTProcessor().with_optimized(false).with_threshold(8).process_batch({{&a1, &b1, &f1, &g1}, {&a2, &b2, &f2, &g2}});
            FFT_2::twofft_optimized(a1, b1, r1, s1);
            CHECK(r1 == f1 && s1 == g1);
            auto ri = FFT_2::extract_real_imag(f2);
            CHECK(FFT_2::combine_real_imag(ri.first, ri.second) == f2);
            std::vector<double> x(64), y(64), x0, y0;
            for (std::size_t i = 0; i < 64; ++i) { x[i] = std::sin(0.2 * i); y[i] = 1.0 / (1 + i); }
            x0 = x; y0 = y;
            Real_FT::RealFTProcessor().process_batch({{&x, 64, 1}, {&y, 64, 1}});
            Real_FT::realft_optimized(x0, 64, 1);
            CHECK(x == x0);
            Real_FT::RealFTProcessor().process_batch({{&x, 64, -1}, {&y, 64, 7}});      // anything but 1 is the inverse (Real_FT.rs:15)
            for (std::size_t i = 0; i < 64; ++i) CHECK(std::fabs(y[i] / 32.0 - y0[i]) < 1e-12);
        }
        // Cos_FT2.rs:248-264 round trip (true factor n/2); cosft1 of a constant: F_0 = n c, F_k = 0 for even k > 0
        {
            const std::size_t m = 16;
            std::vector<double> y(m + 1), y0;
            for (std::size_t i = 0; i <= m; ++i) y[i] = i == 0 ? 0.0 : std::sin((double)i);
            y0 = y;
            Cos_FT2::cosft2(y, m, 1);
            Cos_FT2::cosft2(y, m, -1);
            for (std::size_t i = 1; i <= m; ++i) CHECK(std::fabs(y[i] * 2.0 / m - y0[i]) < 1e-10);
            std::vector<double> c(m + 2, 1.0);
            Cos_FT::cosft1(c, m);
            CHECK(std::fabs(c[1] - (double)m) < 1e-10 && std::fabs(c[3]) < 1e-10 && std::fabs(c[2]) < 1e-10);
        }
        // N1: device-resident chain rlft3 -> identity-kernel product -> rlft3^-1 (kernel = unit impulse)
        {
            const std::size_t dims[3] = {8, 8, 8};
            nrb_plan_t plan = nullptr;
            CHECK(nrb_plan_create(NRB_KIND_RLFT3, dims, 3, 1, &plan) == NRB_OK);
            std::vector<double> v(512), imp(512, 0.0);
            for (int i = 0; i < 512; ++i) v[i] = std::cos(0.37 * i);
            imp[0] = 1.0;
            DeviceBuffer dv(v), dk(imp), sv(128), sk(128);
            CHECK(nrb_plan_exec(plan, dk.data(), sk.data(), nullptr, 1, 0, nullptr) == NRB_OK);
            CHECK(nrb_plan_exec(plan, dv.data(), sv.data(), nullptr, 1, 0, nullptr) == NRB_OK);
            CHECK(nrb_complex_multiply_device(dv.data(), dk.data(), 256, 0, 2.0 / 512, nullptr) == NRB_OK);
            CHECK(nrb_complex_multiply_device(sv.data(), sk.data(), 64, 0, 2.0 / 512, nullptr) == NRB_OK);
            CHECK(nrb_plan_exec(plan, dv.data(), sv.data(), nullptr, -1, 0, nullptr) == NRB_OK);
            auto back = dv.download();
            for (int i = 0; i < 512; ++i) CHECK(std::fabs(back[i] - v[i]) < 1e-12);
            nrb_plan_destroy(plan);
        }
        // Real_FT3.rs:268-311 with the true factor N/2
        std::vector<double> d(512), s(128, 0.0), o(512);
        for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) for (int k = 0; k < 8; ++k) o[(i * 8 + j) * 8 + k] = d[(i * 8 + j) * 8 + k] = i + j + k;
        Real_FT3::rlft3(d, s, 8, 8, 8, 1);
        Real_FT3::rlft3(d, s, 8, 8, 8, -1);
        for (int i = 0; i < 512; ++i) CHECK(std::fabs(d[i] / 256.0 - o[i]) < 1e-10);
        // realft round trip = (n/2) x
        std::vector<double> x(256), x0;
        for (int i = 0; i < 256; ++i) x[i] = std::sin(0.1 * i);
        x0 = x;
        Real_FT::realft(x, 256, 1);
        Real_FT::realft(x, 256, -1);
        for (int i = 0; i < 256; ++i) CHECK(std::fabs(x[i] / 128.0 - x0[i]) < 1e-10);
    } else {
        // no device: compute calls must fail loudly, never fall back
        if (nrb_device_count() == 0) {
            std::vector<double> c(16, 1.0);
            CHECK(throws<Panic>([&] { FFT_1::four1(c, 8, 1); }));
        }
    }
    std::printf(failures ? "host mirror: %d failures\n" : "host mirror: ok\n", failures);
    return failures ? 1 : 0;
}
