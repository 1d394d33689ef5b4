// num_rs.hpp -- C++ host-side mirror of the reference's public interface for the FFT hot path.
//
// The reference is a Rust crate (num_rs); its toolchain is not available in this image, so the
// host side above the C ABI is provided in C++ (the Rust shim a maintainer would add is in
// rust/ and INTEGRATION.md).  Names, argument meaning and error behaviour follow
// /root/reference/src/{FFT_1,Fourn,Real_FT,Real_FT3,Convolve,Correlation}.rs:
//   * functions returning () in Rust and panicking on misuse (four1, realft, rlft3) throw
//     num_rs::Panic here;
//   * functions returning Result<_, ConvlvError | CorrelError | io::Error> throw the matching
//     exception type carrying the same variant.
// Header-only; link with -lnumrs_b200.
#pragma once
#include <cstddef>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/numrs_b200.h"

namespace num_rs {

struct Panic : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline void panic_on(int rc)
{
    if (rc != NRB_OK) throw Panic(std::string("numrs_b200: ") + nrb_last_error());
}

// ------------------------------------------------------------------ FFT_1.rs
namespace FFT_1 {

// FFT_1.rs:5
inline void four1(double *data, std::size_t data_len, std::size_t nn, int isign)
{
    if (data_len < 2 * nn) throw Panic("index out of bounds: data.len() < 2*nn");
    panic_on(nrb_four1(data, nn, isign));
}
inline void four1(std::vector<double> &data, std::size_t nn, int isign) { four1(data.data(), data.size(), nn, isign); }
// FFT_1.rs:110
inline void four1_optimized(std::vector<double> &data, std::size_t nn, int isign) { four1(data, nn, isign); }

// FFT_1.rs:143-190
class FFTProcessor {
  public:
    FFTProcessor() : max_threads_(1), use_optimized_(true) {}
    FFTProcessor &with_threads(std::size_t t) { max_threads_ = t; return *this; }
    FFTProcessor &with_optimized(bool b) { use_optimized_ = b; return *this; }
    void fft(std::vector<double> &data, int isign) const { four1(data, data.size() / 2, isign); }
    // batches: mutable slices (pointer, length in doubles)
    void fft_batch(const std::vector<std::pair<double *, std::size_t>> &batches, int isign) const
    {
        std::vector<double *> ptrs;
        std::vector<std::size_t> nn;
        for (const auto &b : batches) { ptrs.push_back(b.first); nn.push_back(b.second / 2); }
        panic_on(nrb_four1_batch(ptrs.data(), nn.data(), ptrs.size(), isign));
    }

  private:
    std::size_t max_threads_;
    bool use_optimized_;
};

// FFT_1.rs:193-228 (host-side helpers)
inline std::vector<double> real_to_complex(const std::vector<double> &re)
{
    std::vector<double> c(2 * re.size(), 0.0);
    for (std::size_t i = 0; i < re.size(); ++i) c[2 * i] = re[i];
    return c;
}
inline std::vector<double> complex_to_real(const std::vector<double> &c)
{
    std::vector<double> r((c.size() + 1) / 2);
    for (std::size_t i = 0; i < r.size(); ++i) r[i] = c[2 * i];
    return r;
}
inline std::vector<double> power_spectrum(const std::vector<double> &c)
{
    std::vector<double> p((c.size() + 1) / 2, 0.0);
    for (std::size_t i = 0; 2 * i + 1 < c.size(); ++i) p[i] = c[2 * i] * c[2 * i] + c[2 * i + 1] * c[2 * i + 1];
    return p;
}

// device versions of the spectra: same result, computed by the CUDA library
inline std::vector<double> power_spectrum_device(const std::vector<double> &c, bool take_sqrt = false)
{
    std::vector<double> p(c.size() / 2);
    panic_on(nrb_power_spectrum(c.data(), p.size(), take_sqrt ? 1 : 0, p.data()));
    return p;
}

} // namespace FFT_1

// ------------------------------------------------------------------ FFT_2.rs
namespace FFT_2 {
// FFT_2.rs:3 twofft(data1, data2, fft1, fft2); asserts FFT_2.rs:5-7
inline void twofft(const std::vector<double> &data1, const std::vector<double> &data2, std::vector<double> &fft1,
                   std::vector<double> &fft2)
{
    const std::size_t n = data1.size();
    if (data2.size() != n) throw Panic("data2 length must equal data1 length");
    if (fft1.size() != 2 * n + 2) throw Panic("fft1 must have length 2*n + 2");
    if (fft2.size() != 2 * n + 2) throw Panic("fft2 must have length 2*n + 2");
    panic_on(nrb_twofft(data1.data(), data2.data(), n, fft1.data(), fft2.data()));
}
inline void twofft_optimized(const std::vector<double> &data1, const std::vector<double> &data2, std::vector<double> &fft1,
                             std::vector<double> &fft2) { twofft(data1, data2, fft1, fft2); }   // FFT_2.rs:135

// FFT_2.rs:222-265.  The builder flags select CPU code paths in the reference; kept and ignored (one device path).
class TwoFFTProcessor {
public:
    struct Item { const std::vector<double> *data1, *data2; std::vector<double> *fft1, *fft2; };
    TwoFFTProcessor() : use_optimized_(true), parallel_threshold_(1024) {}
    TwoFFTProcessor &with_optimized(bool v) { use_optimized_ = v; return *this; }
    TwoFFTProcessor &with_threshold(std::size_t t) { parallel_threshold_ = t; return *this; }
    void process(const std::vector<double> &data1, const std::vector<double> &data2, std::vector<double> &fft1,
                 std::vector<double> &fft2) const { twofft(data1, data2, fft1, fft2); }
    // FFT_2.rs:258 process_batch: runs of tuples with equal length go to the device as one batched plan
    void process_batch(const std::vector<Item> &batches) const
    {
        std::size_t i = 0;
        while (i < batches.size()) {
            const std::size_t n = batches[i].data1->size();
            std::vector<const double *> a, b;
            std::vector<double *> f1, f2;
            std::size_t j = i;
            for (; j < batches.size() && batches[j].data1->size() == n; ++j) {
                const Item &it = batches[j];
                if (it.data2->size() != n) throw Panic("data2 length must equal data1 length");
                if (it.fft1->size() != 2 * n + 2) throw Panic("fft1 must have length 2*n + 2");
                if (it.fft2->size() != 2 * n + 2) throw Panic("fft2 must have length 2*n + 2");
                a.push_back(it.data1->data()); b.push_back(it.data2->data());
                f1.push_back(it.fft1->data()); f2.push_back(it.fft2->data());
            }
            panic_on(nrb_twofft_batch(a.data(), b.data(), a.size(), n, f1.data(), f2.data()));
            i = j;
        }
    }
private:
    bool use_optimized_;
    std::size_t parallel_threshold_;
};

// FFT_2.rs:361 / :375
inline std::pair<std::vector<double>, std::vector<double>> extract_real_imag(const std::vector<double> &fft)
{
    const std::size_t n = fft.size() / 2;
    std::vector<double> re(n), im(n);
    for (std::size_t i = 0; i < n; ++i) { re[i] = fft[2 * i]; im[i] = fft[2 * i + 1]; }
    return {re, im};
}
inline std::vector<double> combine_real_imag(const std::vector<double> &re, const std::vector<double> &im)
{
    if (re.size() != im.size()) throw Panic("Real and imaginary parts must have same length");
    std::vector<double> c(2 * re.size());
    for (std::size_t i = 0; i < re.size(); ++i) { c[2 * i] = re[i]; c[2 * i + 1] = im[i]; }
    return c;
}
} // namespace FFT_2

// ------------------------------------------------------------------ Cos_FT.rs / Cos_FT2.rs / sinft (README.md:72)
namespace Cos_FT {
// Cos_FT.rs:7 cosft1(y, n): 1-based array, y[0] unused, data y[1..=n+1] (an out-of-range index panics in Rust)
inline void cosft1(std::vector<double> &y, std::size_t n)
{
    if (y.size() < n + 2) throw Panic("index out of bounds: y must hold n + 2 elements");
    panic_on(nrb_cosft1(y.data(), n));
}
inline void cosft1_optimized(std::vector<double> &y, std::size_t n) { cosft1(y, n); }   // Cos_FT.rs:77
} // namespace Cos_FT
namespace Cos_FT2 {
// Cos_FT2.rs:7 cosft2(y, n, isign): panics on isign outside {1, -1} (Cos_FT2.rs:11)
inline void cosft2(std::vector<double> &y, std::size_t n, int isign)
{
    if (isign != 1 && isign != -1) throw Panic("Invalid isign value. Must be 1 or -1");
    if (y.size() < n + 1) throw Panic("index out of bounds: y must hold n + 1 elements");
    panic_on(nrb_cosft2(y.data(), n, isign));
}
inline void cosft2_simd(std::vector<double> &y, std::size_t n, int isign) { cosft2(y, n, isign); }   // Cos_FT2.rs:202
} // namespace Cos_FT2
namespace Sin_FT {
inline void sinft(std::vector<double> &y, std::size_t n)
{
    if (y.size() < n + 1) throw Panic("index out of bounds: y must hold n + 1 elements");
    panic_on(nrb_sinft(y.data(), n));
}
} // namespace Sin_FT

// ------------------------------------------------------------------ Fourn.rs / Real_FT3.rs:35
namespace Fourn {

struct InvalidInput : std::invalid_argument {   // io::ErrorKind::InvalidInput, Fourn.rs:367-378
    using std::invalid_argument::invalid_argument;
};

// the in-memory call shape of Real_FT3.rs:35: Fourn(&mut flat, &nn, ndim, isign)
inline void fourn(std::vector<double> &data, const std::vector<std::size_t> &nn, std::size_t ndim, int isign)
{
    if (ndim == 0 || ndim > nn.size()) throw InvalidInput("Invalid dimensions");
    std::size_t total = 1;
    for (std::size_t d = 0; d < ndim; ++d) total *= nn[d];
    const int rc = nrb_fourn(data.data(), nn.data(), ndim, isign);
    if (rc == NRB_ERR_INVALID_DIMS || rc == NRB_ERR_INVALID_ISIGN) throw InvalidInput(nrb_last_error());
    if (rc == NRB_OK && data.size() < 2 * total) throw Panic("data.len() < 2*prod(nn)");
    panic_on(rc);
}

} // namespace Fourn

// ------------------------------------------------------------------ Real_FT.rs
namespace Real_FT {

// Real_FT.rs:4 (asserts :5-6)
inline void realft(std::vector<double> &data, std::size_t n, int isign)
{
    if (n % 2 != 0) throw Panic("n must be even");
    if (data.size() < n) throw Panic("data length must be at least n");
    panic_on(nrb_realft(data.data(), n, isign));
}
inline void realft_optimized(std::vector<double> &data, std::size_t n, int isign) { realft(data, n, isign); }

// Real_FT.rs:332-370 (builder flags kept and ignored: one device path)
class RealFTProcessor {
public:
    struct Item { std::vector<double> *data; std::size_t n; int isign; };
    RealFTProcessor() : use_optimized_(true), parallel_threshold_(1024) {}
    RealFTProcessor &with_optimized(bool v) { use_optimized_ = v; return *this; }
    RealFTProcessor &with_threshold(std::size_t t) { parallel_threshold_ = t; return *this; }
    void process(std::vector<double> &data, std::size_t n, int isign) const { realft(data, n, isign); }
    // Real_FT.rs:365 process_batch: runs of equal (n, direction) go to the device as one batch
    void process_batch(const std::vector<Item> &batches) const
    {
        std::size_t i = 0;
        while (i < batches.size()) {
            const std::size_t n = batches[i].n;
            const int dir = batches[i].isign == 1 ? 1 : -1;      // Real_FT.rs:10,15: anything but 1 is the inverse
            std::vector<double *> ptrs;
            std::size_t j = i;
            for (; j < batches.size() && batches[j].n == n && (batches[j].isign == 1 ? 1 : -1) == dir; ++j) {
                if (n % 2 != 0) throw Panic("n must be even");
                if (batches[j].data->size() < n) throw Panic("data length must be at least n");
                ptrs.push_back(batches[j].data->data());
            }
            panic_on(nrb_realft_batch(ptrs.data(), n, ptrs.size(), dir));
            i = j;
        }
    }
private:
    bool use_optimized_;
    std::size_t parallel_threshold_;
};

} // namespace Real_FT

// ------------------------------------------------------------------ Real_FT3.rs
namespace Real_FT3 {

// Real_FT3.rs:8; data is [nn1][nn2][nn3] row-major, speq is [nn1][2*nn2] (asserts :17-19)
inline void rlft3(std::vector<double> &data, std::vector<double> &speq, std::size_t nn1, std::size_t nn2,
                  std::size_t nn3, int isign)
{
    if (isign != 1 && isign != -1) throw Panic("isign must be 1 or -1");
    if (data.size() != nn1 * nn2 * nn3) throw Panic("data dimensions mismatch");
    if (speq.size() != nn1 * 2 * nn2) throw Panic("speq dimensions mismatch");
    panic_on(nrb_rlft3(data.data(), speq.data(), nn1, nn2, nn3, isign));
}
// Real_FT3.rs:145 rlft3_optimized: flat slices, same transform
inline void rlft3_optimized(std::vector<double> &data, std::vector<double> &speq, std::size_t nn1, std::size_t nn2,
                            std::size_t nn3, int isign) { rlft3(data, speq, nn1, nn2, nn3, isign); }

} // namespace Real_FT3

// ------------------------------------------------------------------ Convolve.rs
namespace Convolve {

struct ConvlvError : std::runtime_error {   // Convolve.rs:226-238
    enum Kind { EmptyInput, ResponseTooLong, InvalidIsign, DivisionByZero, FftError } kind;
    ConvlvError(Kind k, const std::string &m) : std::runtime_error(m), kind(k) {}
};

inline void raise(int rc)
{
    switch (rc) {
    case NRB_OK: return;
    case NRB_ERR_EMPTY_INPUT: throw ConvlvError(ConvlvError::EmptyInput, "Input arrays cannot be empty");
    case NRB_ERR_RESPONSE_TOO_LONG: throw ConvlvError(ConvlvError::ResponseTooLong, "Response function longer than data");
    case NRB_ERR_INVALID_ISIGN: throw ConvlvError(ConvlvError::InvalidIsign, "isign must be 1 (convolution) or -1 (deconvolution)");
    default: throw ConvlvError(ConvlvError::FftError, std::string("FFT computation error: ") + nrb_last_error());
    }
}

// Convolve.rs:8
inline std::vector<double> convlv(const std::vector<double> &data, const std::vector<double> &respns, int isign,
                                  int pad_mode = NRB_PAD_LITERAL)
{
    std::vector<double> ans(data.size());
    raise(nrb_convlv(data.data(), data.size(), respns.data(), respns.size(), isign, pad_mode, ans.data()));
    return ans;
}

// Convolve.rs:241
inline std::vector<std::vector<double>> convlv_batch(const std::vector<std::vector<double>> &batch,
                                                     const std::vector<double> &respns, int isign)
{
    std::vector<std::vector<double>> out;
    if (batch.empty()) return out;
    const std::size_t n = batch[0].size();
    bool uniform = true;
    for (const auto &b : batch) uniform = uniform && b.size() == n;
    if (!uniform) {
        for (const auto &b : batch) out.push_back(convlv(b, respns, isign));
        return out;
    }
    std::vector<const double *> in;
    std::vector<double *> o;
    out.assign(batch.size(), std::vector<double>(n));
    for (std::size_t i = 0; i < batch.size(); ++i) { in.push_back(batch[i].data()); o.push_back(out[i].data()); }
    raise(nrb_convlv_batch(in.data(), in.size(), n, respns.data(), respns.size(), isign, NRB_PAD_LITERAL, o.data()));
    return out;
}

// Convolve.rs:253-339
class ConvlvProcessor {
  public:
    ConvlvProcessor() : use_optimized_(true), threshold_(1024) {}
    ConvlvProcessor &with_optimized(bool b) { use_optimized_ = b; return *this; }
    ConvlvProcessor &with_threshold(std::size_t t) { threshold_ = t; return *this; }
    std::vector<double> process(const std::vector<double> &d, const std::vector<double> &r, int isign) const { return convlv(d, r, isign); }

  private:
    bool use_optimized_;
    std::size_t threshold_;
};

} // namespace Convolve

// ------------------------------------------------------------------ Correlation.rs
namespace Correlation {

struct CorrelError : std::runtime_error {   // Correlation.rs:389-399
    enum Kind { EmptyInput, LengthMismatch, FftError, ZeroStdDev } kind;
    CorrelError(Kind k, const std::string &m) : std::runtime_error(m), kind(k) {}
};

inline void raise(int rc)
{
    switch (rc) {
    case NRB_OK: return;
    case NRB_ERR_EMPTY_INPUT: throw CorrelError(CorrelError::EmptyInput, "Input arrays cannot be empty");
    case NRB_ERR_LENGTH_MISMATCH: throw CorrelError(CorrelError::LengthMismatch, "Input arrays must have the same length");
    case NRB_ERR_ZERO_STDDEV: throw CorrelError(CorrelError::ZeroStdDev, "Normalization error: standard deviation is zero");
    default: throw CorrelError(CorrelError::FftError, std::string("FFT computation error: ") + nrb_last_error());
    }
}

// Correlation.rs:8
inline std::vector<double> correl(const std::vector<double> &a, const std::vector<double> &b)
{
    std::vector<double> ans(a.size());
    raise(nrb_correl(a.data(), a.size(), b.data(), b.size(), ans.data()));
    return ans;
}
// Correlation.rs:281
inline std::vector<double> autocorrel(const std::vector<double> &a) { return correl(a, a); }

// Correlation.rs:273
inline std::vector<std::vector<double>> correl_batch(const std::vector<std::pair<std::vector<double>, std::vector<double>>> &pairs)
{
    std::vector<std::vector<double>> out;
    if (pairs.empty()) return out;
    const std::size_t n = pairs[0].first.size();
    bool uniform = n > 0;
    for (const auto &p : pairs) uniform = uniform && p.first.size() == n && p.second.size() == n;
    if (!uniform) {      // mixed lengths (or an error to report per pair): one call each, as the reference's par_iter does
        for (const auto &p : pairs) out.push_back(correl(p.first, p.second));
        return out;
    }
    // one device batch (sharded over the GPUs when the option num_devices says so)
    std::vector<const double *> a, b;
    std::vector<double *> o;
    out.assign(pairs.size(), std::vector<double>(n));
    for (std::size_t i = 0; i < pairs.size(); ++i) { a.push_back(pairs[i].first.data()); b.push_back(pairs[i].second.data()); o.push_back(out[i].data()); }
    raise(nrb_correl_batch(a.data(), b.data(), pairs.size(), n, o.data()));
    return out;
}

// Correlation.rs:189 / :226 / :286
inline std::vector<double> correl_normalized(const std::vector<double> &a, const std::vector<double> &b)
{
    std::vector<double> ans(a.size());
    raise(nrb_correl_normalized(a.data(), a.size(), b.data(), b.size(), 0, ans.data()));
    return ans;
}
inline std::vector<double> correl_normalized_fast(const std::vector<double> &a, const std::vector<double> &b)
{
    std::vector<double> ans(a.size());
    raise(nrb_correl_normalized(a.data(), a.size(), b.data(), b.size(), 1, ans.data()));
    return ans;
}
inline std::vector<double> autocorrel_fast(const std::vector<double> &a)
{
    std::vector<double> ans(a.size());
    raise(nrb_autocorrel_fast(a.data(), a.size(), ans.data()));
    return ans;
}

} // namespace Correlation

// ------------------------------------------------------------------ device-resident buffers
// RAII owner of device memory; upload / download are synchronous here (NULL stream + synchronise).
class DeviceBuffer {
  public:
    explicit DeviceBuffer(std::size_t doubles) : n_(doubles), p_(nullptr) { panic_on(nrb_device_alloc(8 * (doubles ? doubles : 1), &p_)); }
    explicit DeviceBuffer(const std::vector<double> &host) : DeviceBuffer(host.size()) { upload(host); }
    ~DeviceBuffer() { if (p_) nrb_device_free(p_); }
    DeviceBuffer(const DeviceBuffer &) = delete;
    DeviceBuffer &operator=(const DeviceBuffer &) = delete;
    void upload(const std::vector<double> &host)
    {
        if (host.size() > n_) throw Panic("DeviceBuffer::upload: source larger than the buffer");
        panic_on(nrb_upload(p_, host.data(), 8 * host.size(), nullptr));
        panic_on(nrb_stream_synchronize(nullptr));
    }
    std::vector<double> download() const
    {
        std::vector<double> host(n_);
        panic_on(nrb_download(host.data(), p_, 8 * n_, nullptr));
        panic_on(nrb_stream_synchronize(nullptr));
        return host;
    }
    double *data() const { return static_cast<double *>(p_); }
    std::size_t size() const { return n_; }

  private:
    std::size_t n_;
    void *p_;
};

} // namespace num_rs
