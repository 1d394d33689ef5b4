/*
 * nr_oracle.c -- CPU restatement of the numrs FFT hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for numrs_b200.  It is NOT part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product library (numrs_b200/libnumrs_b200.so) never links or calls it and
 * has no CPU fallback.
 *
 * The reference (SciRustaceans/numrs, Rust) cannot be compiled in this environment (no Rust
 * toolchain; the FFT modules are commented out of the crate and do not type-check, see
 * SURVEY.md section 0).  Each routine below therefore restates the reference source line by
 * line where the reference is correct ("literal"), and restates the Numerical Recipes routine
 * the reference file transliterates where the reference panics / is a placeholder
 * ("NR intent").  Every departure is listed in SURVEY.md section 8c (deviation ledger D1..D8).
 *
 * Pinning status:
 *   four1                 pinned by FFT_1.rs:246-267 (round trip) + numpy cross-check
 *   convlv / correl small pinned by Convolve.rs:355-357,452 and Correlation.rs:414,474-475,487,500-501
 *   realft, fourn, rlft3, large-n convlv/correl: "parity unpinned" by the reference's own tests
 *                         (none of them can run); pinned here against numpy (pocketfft) and an
 *                         mpmath 50-digit DFT in tests/test_oracle.py.
 *   next rows, at the end of this file, ledger D9..D11:
 *     correl_normalized_fast, autocorrel (small n), power/magnitude spectrum: pinned by Correlation.rs:505-512,
 *                         :481-491 (lag 0) and FFT_1.rs:305-321
 *     twofft, correl_normalized, autocorrel_fast (n > 32), cosft1, cosft2, sinft: "parity unpinned" by the
 *                         reference (twofft's mirror is off by one, FFT_2.rs:74; the cosine transforms stop at
 *                         `unimplemented!()`, Cos_FT.rs:70-74; sinft has no source); pinned here against numpy
 *                         and scipy.fft dct / dst in tests/test_oracle.py.
 *
 * All file:line citations are relative to /root/reference/src.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PI 3.14159265358979323846264338327950288

/* ------------------------------------------------------------------------------------------
 * Synthetic input generator: counter-based, reproducible in
 * C / CUDA / Python.  u(i) = splitmix64(seed * 0x9E3779B97F4A7C15 + i),
 * x_i = (u >> 11) * 2^-52 - 1  in [-1, 1).
 * ---------------------------------------------------------------------------------------- */
static inline uint64_t orc_splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

void orc_fill_uniform(uint64_t seed, uint64_t offset, size_t count, double *out)
{
    const uint64_t base = seed * 0x9E3779B97F4A7C15ULL + offset;
    size_t i;
#pragma omp parallel for schedule(static)
    for (i = 0; i < count; ++i) {
        uint64_t u = orc_splitmix64(base + (uint64_t)i);
        out[i] = (double)(u >> 11) * (1.0 / 4503599627370496.0) - 1.0;
    }
}

/* ------------------------------------------------------------------------------------------
 * four1 -- FFT_1.rs:5-44 (literal).
 * Bit reversal FFT_1.rs:8-22; Danielson-Lanczos stages FFT_1.rs:25-43; stages with
 * mmax < 1024 use the trigonometric recurrence (process_butterflies_sequential,
 * FFT_1.rs:47-72), stages with mmax >= 1024 use an exact cos/sin table
 * (process_butterflies_parallel, FFT_1.rs:75-107).  `mmax` counts doubles, as in the
 * reference (n = 2*nn doubles).
 * ---------------------------------------------------------------------------------------- */
static void orc_bitrev(double *data, size_t nn)
{
    /* FFT_1.rs:8-22 (1-based j/i on a 0-based slice: swap(j-1,i-1), swap(j,i)) */
    const size_t n = nn * 2;
    size_t j = 1, i;
    for (i = 1; i < n; i += 2) {
        if (j > i) {
            double t;
            t = data[j - 1]; data[j - 1] = data[i - 1]; data[i - 1] = t;
            t = data[j];     data[j]     = data[i];     data[i]     = t;
        }
        size_t m = nn;
        while (m >= 2 && j > m) { j -= m; m >>= 1; }
        j += m;
    }
}

static void orc_stage_recurrence(double *data, size_t n, size_t mmax, size_t istep,
                                 double wpr, double wpi)
{
    /* FFT_1.rs:47-72: twiddle index outer, stride-istep inner, recurrence update */
    double wr = 1.0, wi = 0.0;
    size_t m, i;
    for (m = 0; m < mmax; m += 2) {
        for (i = m; i < n; i += istep) {
            size_t j = i + mmax;
            if (j >= n) continue;
            double tempr = wr * data[j] - wi * data[j + 1];
            double tempi = wr * data[j + 1] + wi * data[j];
            data[j] = data[i] - tempr;
            data[j + 1] = data[i + 1] - tempi;
            data[i] += tempr;
            data[i + 1] += tempi;
        }
        double wtemp = wr;
        wr = wtemp * wpr - wi * wpi + wr;
        wi = wi * wpr + wtemp * wpi + wi;
    }
}

static void orc_stage_table(double *data, size_t n, size_t mmax, size_t istep, int isign,
                            int threaded)
{
    /* FFT_1.rs:75-107: exact (cos, sin) per twiddle, chunks of istep doubles processed
     * independently (par_chunks_mut(istep), FFT_1.rs:85). */
    const size_t half = mmax / 2;
    double *rot = (double *)malloc(sizeof(double) * 2 * half);
    size_t m;
    for (m = 0; m < half; ++m) {
        double angle = (double)isign * 2.0 * ORC_PI * (double)m / (double)mmax;
        rot[2 * m] = cos(angle);
        rot[2 * m + 1] = sin(angle);
    }
    const ptrdiff_t nchunks = (ptrdiff_t)(n / istep);
    ptrdiff_t c;
    (void)threaded;
#pragma omp parallel for schedule(static) if (threaded && nchunks > 1)
    for (c = 0; c < nchunks; ++c) {
        double *chunk = data + (size_t)c * istep;
        size_t mm;
        for (mm = 0; mm < half; ++mm) {
            const double wr = rot[2 * mm], wi = rot[2 * mm + 1];
            const size_t i = mm * 2, j = i + mmax;
            double tempr = wr * chunk[j] - wi * chunk[j + 1];
            double tempi = wr * chunk[j + 1] + wi * chunk[j];
            chunk[j] = chunk[i] - tempr;
            chunk[j + 1] = chunk[i + 1] - tempi;
            chunk[i] += tempr;
            chunk[i + 1] += tempi;
        }
    }
    free(rot);
}

static void orc_four1_impl(double *data, size_t nn, int isign, int hybrid, int threaded)
{
    const size_t n = nn * 2;
    if (nn < 2) return;
    orc_bitrev(data, nn);
    size_t mmax = 2;
    while (n > mmax) {
        const size_t istep = mmax << 1;
        const double theta = (double)isign * (2.0 * ORC_PI / (double)mmax);
        const double wtemp = sin(0.5 * theta);
        const double wpr = -2.0 * wtemp * wtemp;
        const double wpi = sin(theta);
        if (hybrid && mmax >= 1024)
            orc_stage_table(data, n, mmax, istep, isign, threaded); /* FFT_1.rs:34-36 */
        else
            orc_stage_recurrence(data, n, mmax, istep, wpr, wpi);   /* FFT_1.rs:37-40 */
        mmax = istep;
    }
}

/* FFT_1.rs:5 `pub fn four1(data, nn, isign)`; single-threaded */
void orc_four1(double *data, size_t nn, int isign) { orc_four1_impl(data, nn, isign, 1, 0); }
/* same, with the reference's per-stage chunk parallelism (par_chunks_mut, FFT_1.rs:85) */
void orc_four1_mt(double *data, size_t nn, int isign) { orc_four1_impl(data, nn, isign, 1, 1); }
/* FFT_1.rs:110-140 `four1_optimized`: recurrence twiddles on every stage; this is also the
 * private copy Real_FT.rs:430-476 that realft calls. */
void orc_four1_optimized(double *data, size_t nn, int isign) { orc_four1_impl(data, nn, isign, 0, 0); }

/* FFT_1.rs:166-182 FFTProcessor::fft dispatch (use_optimized defaults to true) and
 * FFT_1.rs:185-189 fft_batch = one transform per rayon task. */
void orc_fft_batch(double *const *ptrs, const size_t *nn, size_t count, int isign, int threaded)
{
    ptrdiff_t b;
#pragma omp parallel for schedule(dynamic) if (threaded)
    for (b = 0; b < (ptrdiff_t)count; ++b) {
        if (nn[b] >= 512) orc_four1_optimized(ptrs[b], nn[b], isign);
        else orc_four1(ptrs[b], nn[b], isign);
    }
}

/* ------------------------------------------------------------------------------------------
 * fourn -- NR intent (deviation D3).  The reference has no in-memory N-dimensional FFT:
 * Real_FT3.rs:35,133,168,256 call `Fourn(&mut flat, &nn, 3, isign)` which exists nowhere
 * (Fourn.rs is a file-based skeleton).  This is NR `fourn`: in-place, row-major, nn[0]
 * slowest, last index fastest, F = sum x * exp(isign*2*pi*i*sum_d k_d j_d / nn_d),
 * unnormalised.  The independent inner lines (i1, i3 loops) carry no dependency and are
 * run in parallel when `threaded`.
 * ---------------------------------------------------------------------------------------- */
static void orc_fourn_impl(double *data0, const size_t *nn, int ndim, int isign, int threaded)
{
    double *data = data0 - 1; /* NR 1-based view */
    size_t ntot = 1, nprev = 1;
    int idim;
    (void)threaded;
    for (idim = 0; idim < ndim; ++idim) ntot *= nn[idim];
    for (idim = ndim - 1; idim >= 0; --idim) {
        const size_t n = nn[idim];
        const size_t nrem = ntot / (n * nprev);
        const size_t ip1 = nprev << 1;
        const size_t ip2 = ip1 * n;
        const size_t ip3 = ip2 * nrem;
        size_t i2rev = 1, i2;
        for (i2 = 1; i2 <= ip2; i2 += ip1) {
            if (i2 < i2rev) {
                const ptrdiff_t cnt1 = (ptrdiff_t)(ip1 / 2);
                ptrdiff_t c1;
#pragma omp parallel for schedule(static) if (threaded && ip3 > 65536)
                for (c1 = 0; c1 < cnt1; ++c1) {
                    size_t i1 = i2 + 2 * (size_t)c1, i3;
                    for (i3 = i1; i3 <= ip3; i3 += ip2) {
                        size_t i3rev = i2rev + i3 - i2;
                        double t;
                        t = data[i3]; data[i3] = data[i3rev]; data[i3rev] = t;
                        t = data[i3 + 1]; data[i3 + 1] = data[i3rev + 1]; data[i3rev + 1] = t;
                    }
                }
            }
            size_t ibit = ip2 >> 1;
            while (ibit >= ip1 && i2rev > ibit) { i2rev -= ibit; ibit >>= 1; }
            i2rev += ibit;
        }
        size_t ifp1 = ip1;
        while (ifp1 < ip2) {
            const size_t ifp2 = ifp1 << 1;
            const double theta = (double)isign * (2.0 * ORC_PI) / (double)(ifp2 / ip1);
            const double wtemp0 = sin(0.5 * theta);
            const double wpr = -2.0 * wtemp0 * wtemp0;
            const double wpi = sin(theta);
            double wr = 1.0, wi = 0.0;
            size_t i3;
            for (i3 = 1; i3 <= ifp1; i3 += ip1) {
                const ptrdiff_t cnt1 = (ptrdiff_t)(ip1 / 2);
                const ptrdiff_t cnt2 = (ptrdiff_t)((ip3 - i3) / ifp2 + 1);
                const ptrdiff_t total = cnt1 * cnt2;
                ptrdiff_t t;
#pragma omp parallel for schedule(static) if (threaded && total > 16384)
                for (t = 0; t < total; ++t) {
                    const size_t c2 = (size_t)(t / cnt1), c1 = (size_t)(t % cnt1);
                    const size_t k1 = i3 + 2 * c1 + c2 * ifp2;
                    const size_t k2 = k1 + ifp1;
                    const double tempr = wr * data[k2] - wi * data[k2 + 1];
                    const double tempi = wr * data[k2 + 1] + wi * data[k2];
                    data[k2] = data[k1] - tempr;
                    data[k2 + 1] = data[k1 + 1] - tempi;
                    data[k1] += tempr;
                    data[k1 + 1] += tempi;
                }
                const double wtemp = wr;
                wr = wtemp * wpr - wi * wpi + wr;
                wi = wi * wpr + wtemp * wpi + wi;
            }
            ifp1 = ifp2;
        }
        nprev *= n;
    }
}

/* Validation rules of Fourn.rs:367-378 / :57-62.  Returns 0 ok, -5 invalid dims, -3 isign. */
int orc_fourn_validate(const size_t *nn, size_t nn_len, size_t ndim, int isign)
{
    size_t d;
    if (ndim == 0 || ndim > nn_len) return -5;
    if (isign != 1 && isign != -1) return -3;
    for (d = 0; d < ndim; ++d) if (nn[d] <= 1) return -5;
    return 0;
}

void orc_fourn(double *data, const size_t *nn, int ndim, int isign) { orc_fourn_impl(data, nn, ndim, isign, 0); }
void orc_fourn_mt(double *data, const size_t *nn, int ndim, int isign) { orc_fourn_impl(data, nn, ndim, isign, 1); }

/* ------------------------------------------------------------------------------------------
 * realft -- Real_FT.rs:4-21 structure, NR intent for the index formulas (deviation D1, D2).
 * Constants c1 = 0.5, c2 = -/+0.5, theta = +/-pi/(n/2), recurrence start wr = 1+wpr,
 * wi = wpi (Real_FT.rs:24-33,120-129); DC/Nyquist handling Real_FT.rs:43-45,133-135.
 * The literal indices i1=2i-1.. (Real_FT.rs:61-64) are NR's 1-based formulas applied to a
 * 0-based slice (data[n] out of bounds for n >= 8); the 0-based pairs are
 * i1=2i, i2=2i+1, i3=n-2i, i4=n-2i+1 for i = 1..n/4-1.  The n >= 1024 table branch
 * (Real_FT.rs:83-117) is one twiddle step behind its own sequential branch (D2); both
 * branches are meant to agree (Real_FT.rs:524-545), so the recurrence is used throughout.
 * ---------------------------------------------------------------------------------------- */
static void orc_realft_untangle(double *data, size_t n, double c2, double theta)
{
    const double c1 = 0.5;
    const double wtemp0 = sin(0.5 * theta);
    const double wpr = -2.0 * wtemp0 * wtemp0;
    const double wpi = sin(theta);
    double wr = 1.0 + wpr, wi = wpi;
    size_t i;
    for (i = 1; i < n / 4; ++i) {
        const size_t i1 = 2 * i, i2 = i1 + 1, i3 = n - 2 * i, i4 = i3 + 1;
        const double h1r = c1 * (data[i1] + data[i3]);
        const double h1i = c1 * (data[i2] - data[i4]);
        const double h2r = -c2 * (data[i2] + data[i4]);
        const double h2i = c2 * (data[i1] - data[i3]);
        data[i1] = h1r + wr * h2r - wi * h2i;
        data[i2] = h1i + wr * h2i + wi * h2r;
        data[i3] = h1r - wr * h2r + wi * h2i;
        data[i4] = -h1i + wr * h2i + wi * h2r;
        const double wtemp = wr;
        wr = wtemp * wpr - wi * wpi + wr;
        wi = wi * wpr + wtemp * wpi + wi;
    }
}

void orc_realft(double *data, size_t n, int isign)
{
    const size_t half_n = n / 2;
    if (isign == 1) {                       /* Real_FT.rs:10-14 */
        orc_four1_optimized(data, half_n, 1);
        orc_realft_untangle(data, n, -0.5, ORC_PI / (double)half_n);
        const double h1r = data[0];         /* Real_FT.rs:43-45 */
        data[0] = h1r + data[1];
        data[1] = h1r - data[1];
    } else {                                /* Real_FT.rs:15-20: anything but 1 is inverse */
        const double h1r = data[0];         /* Real_FT.rs:133-135 */
        data[0] = 0.5 * (h1r + data[1]);
        data[1] = 0.5 * (h1r - data[1]);
        orc_realft_untangle(data, n, 0.5, -ORC_PI / (double)half_n);
        orc_four1_optimized(data, half_n, -1);
    }
}

/* ------------------------------------------------------------------------------------------
 * rlft3 -- Real_FT3.rs:8-141 structure (constants :21-29, loop nest :60-127), NR intent
 * for the two places the reference breaks: the missing `Fourn` call (D3) and the `speq`
 * store, which the reference writes at the mirrored column (Real_FT3.rs:47) where NR writes
 * it straight (D4).  data is [nn1][nn2][nn3] row-major real, speq is [nn1][2*nn2].
 * ---------------------------------------------------------------------------------------- */
static void orc_rlft3_impl(double *data, double *speq, size_t nn1, size_t nn2, size_t nn3,
                           int isign, int threaded)
{
    const double c1 = 0.5, c2 = -0.5 * (double)isign;
    const double theta = (double)isign * (2.0 * ORC_PI) / (double)nn3;
    const double wtemp0 = sin(0.5 * theta);
    const double wpr = -2.0 * wtemp0 * wtemp0;
    const double wpi = sin(theta);
    const size_t nn[3] = { nn1, nn2, nn3 >> 1 };
#define D(a, b, c) data[(((a) - 1) * nn2 + ((b) - 1)) * nn3 + ((c) - 1)]
#define S(a, b) speq[((a) - 1) * 2 * nn2 + ((b) - 1)]
    size_t i1, i2, i3;
    if (isign == 1) {
        orc_fourn_impl(data, nn, 3, isign, threaded);
        for (i1 = 1; i1 <= nn1; ++i1) {
            size_t j2 = 0;
            for (i2 = 1; i2 <= nn2; ++i2) {
                S(i1, ++j2) = D(i1, i2, 1);
                S(i1, ++j2) = D(i1, i2, 2);
            }
        }
    }
    for (i1 = 1; i1 <= nn1; ++i1) {
        /* rows i1 and j1 are touched together; sequential over i1 as in Real_FT3.rs:60 */
        const size_t j1 = (i1 != 1 ? nn1 - i1 + 2 : 1);
        double wr = 1.0, wi = 0.0;
        size_t ii3 = 1;
        for (i3 = 1; i3 <= (nn3 >> 2) + 1; ++i3, ii3 += 2) {
            for (i2 = 1; i2 <= nn2; ++i2) {
                if (i3 == 1) {
                    const size_t j2 = (i2 != 1 ? ((nn2 - i2) << 1) + 3 : 1);
                    const double h1r = c1 * (D(i1, i2, 1) + S(j1, j2));
                    const double h1i = c1 * (D(i1, i2, 2) - S(j1, j2 + 1));
                    const double h2i = c2 * (D(i1, i2, 1) - S(j1, j2));
                    const double h2r = -c2 * (D(i1, i2, 2) + S(j1, j2 + 1));
                    D(i1, i2, 1) = h1r + h2r;
                    D(i1, i2, 2) = h1i + h2i;
                    S(j1, j2) = h1r - h2r;
                    S(j1, j2 + 1) = h2i - h1i;
                } else {
                    const size_t j2 = (i2 != 1 ? nn2 - i2 + 2 : 1);
                    const size_t j3 = nn3 + 3 - (i3 << 1);
                    const double h1r = c1 * (D(i1, i2, ii3) + D(j1, j2, j3));
                    const double h1i = c1 * (D(i1, i2, ii3 + 1) - D(j1, j2, j3 + 1));
                    const double h2i = c2 * (D(i1, i2, ii3) - D(j1, j2, j3));
                    const double h2r = -c2 * (D(i1, i2, ii3 + 1) + D(j1, j2, j3 + 1));
                    D(i1, i2, ii3) = h1r + wr * h2r - wi * h2i;
                    D(i1, i2, ii3 + 1) = h1i + wr * h2i + wi * h2r;
                    D(j1, j2, j3) = h1r - wr * h2r + wi * h2i;
                    D(j1, j2, j3 + 1) = -h1i + wr * h2i + wi * h2r;
                }
            }
            const double wtemp = wr;
            wr = wtemp * wpr - wi * wpi + wr;
            wi = wi * wpr + wtemp * wpi + wi;
        }
    }
    if (isign == -1) orc_fourn_impl(data, nn, 3, isign, threaded);
#undef D
#undef S
}

void orc_rlft3(double *data, double *speq, size_t nn1, size_t nn2, size_t nn3, int isign)
{ orc_rlft3_impl(data, speq, nn1, nn2, nn3, isign, 0); }
void orc_rlft3_mt(double *data, double *speq, size_t nn1, size_t nn2, size_t nn3, int isign)
{ orc_rlft3_impl(data, speq, nn1, nn2, nn3, isign, 1); }

/* ------------------------------------------------------------------------------------------
 * convlv -- Convolve.rs:8-38.
 * Check order Convolve.rs:13-21 (literal): empty -> -1, m > n -> -2, isign -> -3.
 * Response padding Convolve.rs:41-63 literal (pad_mode 0, ledger L1) or NR (pad_mode 1).
 * Spectral step: NR intent on the packed realft layout (deviation D6; literal
 * Convolve.rs:106 skips complex pair 0 and so zeroes DC and Nyquist, contradicting the
 * reference's own known answers Convolve.rs:355-357).  Deconvolution guard mag2 < 1e-12 -> 0
 * is literal (Convolve.rs:118-122, ledger L2); DC and Nyquist are treated as complex bins
 * with zero imaginary part.  n must be a power of two (NR; the reference only asserts n even
 * in realft, Real_FT.rs:5): returns -6 otherwise.
 * ---------------------------------------------------------------------------------------- */
void orc_pad_response(const double *respns, size_t m, size_t n, int pad_mode, double *p)
{
    size_t i;
    if (pad_mode == 0) {
        const size_t mid = (m - 1) / 2;                 /* Convolve.rs:42 */
        for (i = 0; i < n; ++i) {
            if (i < mid) p[i] = respns[m - mid + i];    /* Convolve.rs:49-51 */
            else if (i < m) p[i] = respns[i];           /* :52-54 */
            else if (i < n - mid) p[i] = 0.0;           /* :55-57 */
            else p[i] = respns[i - (n - mid)];          /* :58-60 */
        }
    } else {
        /* NR convlv: respns[n-k] = respns[m-k] for k = 1..(m-1)/2, zero in between */
        const size_t half = (m - 1) / 2;
        for (i = 0; i < n; ++i) p[i] = 0.0;
        for (i = 0; i < (m + 1) / 2; ++i) p[i] = respns[i];
        for (i = 1; i <= half; ++i) p[n - i] = respns[m - i];
    }
}

static int orc_is_pow2(size_t n) { return n && !(n & (n - 1)); }

static void orc_spectral_convlv(const double *d, const double *r, size_t n, int isign, double *a)
{
    const double no2 = (double)(n >> 1);
    size_t k;
    if (isign == 1) {
        a[0] = d[0] * r[0] / no2;
        a[1] = d[1] * r[1] / no2;
        for (k = 1; k < n / 2; ++k) {      /* Convolve.rs:112-115 */
            const double dr = d[2 * k], di = d[2 * k + 1], rr = r[2 * k], ri = r[2 * k + 1];
            a[2 * k] = (dr * rr - di * ri) / no2;
            a[2 * k + 1] = (dr * ri + di * rr) / no2;
        }
    } else {
        for (k = 0; k < 2; ++k) {
            const double mag2 = r[k] * r[k];
            a[k] = (mag2 < 1e-12) ? 0.0 : (d[k] * r[k]) / mag2 / no2;
        }
        for (k = 1; k < n / 2; ++k) {      /* Convolve.rs:116-128 */
            const double dr = d[2 * k], di = d[2 * k + 1], rr = r[2 * k], ri = r[2 * k + 1];
            const double mag2 = rr * rr + ri * ri;
            if (mag2 < 1e-12) { a[2 * k] = 0.0; a[2 * k + 1] = 0.0; }
            else {
                a[2 * k] = (dr * rr + di * ri) / mag2 / no2;
                a[2 * k + 1] = (di * rr - dr * ri) / mag2 / no2;
            }
        }
    }
}

int orc_convlv(const double *data, size_t n, const double *respns, size_t m, int isign,
               int pad_mode, double *ans)
{
    if (n == 0 || m == 0) return -1;
    if (m > n) return -2;
    if (isign != 1 && isign != -1) return -3;
    if (!orc_is_pow2(n) || n < 2) return -6;
    double *d = (double *)malloc(sizeof(double) * n);
    double *r = (double *)malloc(sizeof(double) * n);
    memcpy(d, data, sizeof(double) * n);
    orc_pad_response(respns, m, n, pad_mode, r);
    orc_realft(d, n, 1);                     /* Convolve.rs:66-87 */
    orc_realft(r, n, 1);
    orc_spectral_convlv(d, r, n, isign, ans);
    orc_realft(ans, n, -1);                  /* Convolve.rs:34 */
    free(d); free(r);
    return 0;
}

int orc_convlv_batch(const double *const *data, size_t count, size_t n, const double *respns,
                     size_t m, int isign, int pad_mode, double *const *ans, int threaded)
{
    /* Convolve.rs:241-250: par_iter().map(convlv) -- the response FFT is recomputed per signal */
    int rc = 0;
    ptrdiff_t b;
#pragma omp parallel for schedule(dynamic) if (threaded)
    for (b = 0; b < (ptrdiff_t)count; ++b) {
        int r = orc_convlv(data[b], n, respns, m, isign, pad_mode, ans[b]);
        if (r != 0) {
#pragma omp critical
            rc = r;
        }
    }
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * correl -- Correlation.rs:8-34.
 * Checks Correlation.rs:11-16 (literal; the caller passes both lengths): -1 empty, -4 mismatch.
 * n <= 32: correl_direct, linear lags, Correlation.rs:37-50 (literal, ledger L3).
 * n > 32: NR `correl` via realft (deviation D7: the reference's own `realft` here is an O(n^2)
 * "Simple DFT implementation for demonstration", Correlation.rs:325-386):
 *   ans = F1 * conj(F2) / no2 on the packed layout incl. DC and Nyquist as real products
 *   (Correlation.rs:73-74,91-92), inverse realft -> r[j] = sum_k d1[(k+j) mod n] * d2[k].
 * ---------------------------------------------------------------------------------------- */
int orc_correl(const double *d1, size_t n1, const double *d2, size_t n2, double *ans)
{
    const size_t n = n1;
    size_t k;
    if (n == 0) return -1;
    if (n2 != n) return -4;
    if (n <= 32) {
        size_t lag, i;
        for (lag = 0; lag < n; ++lag) {
            double sum = 0.0;
            for (i = 0; i < n - lag; ++i) sum += d1[i + lag] * d2[i];
            ans[lag] = sum;
        }
        return 0;
    }
    if (!orc_is_pow2(n)) return -6;
    double *a = (double *)malloc(sizeof(double) * n);
    double *b = (double *)malloc(sizeof(double) * n);
    memcpy(a, d1, sizeof(double) * n);
    memcpy(b, d2, sizeof(double) * n);
    orc_realft(a, n, 1);
    orc_realft(b, n, 1);
    const double no2 = (double)(n >> 1);
    ans[0] = a[0] * b[0] / no2;
    ans[1] = a[1] * b[1] / no2;
    for (k = 1; k < n / 2; ++k) {
        const double ar = a[2 * k], ai = a[2 * k + 1], br = b[2 * k], bi = b[2 * k + 1];
        ans[2 * k] = (ar * br + ai * bi) / no2;       /* Correlation.rs:91 */
        ans[2 * k + 1] = (ai * br - ar * bi) / no2;   /* Correlation.rs:92 */
    }
    orc_realft(ans, n, -1);
    free(a); free(b);
    return 0;
}

int orc_correl_batch(const double *const *d1, const double *const *d2, size_t count, size_t n,
                     double *const *ans, int threaded)
{
    /* Correlation.rs:273-278 */
    int rc = 0;
    ptrdiff_t b;
#pragma omp parallel for schedule(dynamic) if (threaded)
    for (b = 0; b < (ptrdiff_t)count; ++b) {
        int r = orc_correl(d1[b], n, d2[b], n, ans[b]);
        if (r != 0) {
#pragma omp critical
            rc = r;
        }
    }
    return rc;
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py's reference arm: launchers such as torchrun export OMP_NUM_THREADS=1 to their children, which
 * would silently time the "all host cores" baseline on one core; the caller states the count instead. */
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ==========================================================================================
 * "Next" rows of SURVEY.md section 8(f): callers on either side of the hot path.
 * ======================================================================================== */

/* twofft -- FFT_2.rs:3-17: spectra of two real signals from one complex transform.
 * NR intent (deviation D9): the reference applies NR's 1-based mirror `nn2 - j` (nn2 = 2n+2) to a
 * 0-based array (FFT_2.rs:74,111,194), i.e. pairs bin k with bin n-k+1 (measured 84 % / 142 % error,
 * SURVEY.md appendix A); the 0-based partner of bin k is n-k.  fft1 / fft2 receive the n complex
 * bins F1[k], F2[k] = sum_j d{1,2}[j] exp(+2 pi i jk/n) in their first 2n doubles; the reference
 * demands length 2n+2 (FFT_2.rs:6-7), the two extra doubles are set to 0. */
void orc_twofft(const double *d1, const double *d2, size_t n, double *fft1, double *fft2)
{
    size_t j, k;
    for (j = 0; j < n; ++j) { fft1[2 * j] = d1[j]; fft1[2 * j + 1] = d2[j]; }   /* FFT_2.rs:33-37 */
    fft1[2 * n] = fft1[2 * n + 1] = 0.0;
    fft2[2 * n] = fft2[2 * n + 1] = 0.0;
    orc_four1(fft1, n, 1);                                                       /* FFT_2.rs:13 */
    fft2[0] = fft1[1]; fft1[1] = 0.0; fft2[1] = 0.0;                             /* FFT_2.rs:60-62 */
    for (k = 1; k <= n / 2; ++k) {
        const size_t m = n - k;
        const double rep = 0.5 * (fft1[2 * k] + fft1[2 * m]);                    /* FFT_2.rs:77-80 */
        const double rem = 0.5 * (fft1[2 * k] - fft1[2 * m]);
        const double aip = 0.5 * (fft1[2 * k + 1] + fft1[2 * m + 1]);
        const double aim = 0.5 * (fft1[2 * k + 1] - fft1[2 * m + 1]);
        fft1[2 * k] = rep; fft1[2 * k + 1] = aim;                                /* FFT_2.rs:82-90 */
        fft1[2 * m] = rep; fft1[2 * m + 1] = -aim;
        fft2[2 * k] = aip; fft2[2 * k + 1] = -rem;
        fft2[2 * m] = aip; fft2[2 * m + 1] = rem;
    }
}

/* power / magnitude spectrum -- FFT_1.rs:206-228 (literal) */
void orc_power_spectrum(const double *c, size_t npoints, int take_sqrt, double *out)
{
    size_t i;
    for (i = 0; i < npoints; ++i) {
        const double p = c[2 * i] * c[2 * i] + c[2 * i + 1] * c[2 * i + 1];
        out[i] = take_sqrt ? sqrt(p) : p;
    }
}

/* correl_normalized -- Correlation.rs:189-223 (fast = 0) and correl_normalized_fast :226-270 (fast = 1).
 * Literal: population statistics, -8 when a standard deviation is zero (CorrelError::ZeroStdDev), the
 * signals are centred and scaled, then correl().  The fast variant computes std from the single-pass
 * sums (Correlation.rs:236-244) and has its own n <= 32 branch with the extra 1/(s1 s2 n) factor
 * (Correlation.rs:251-263); for n > 32 neither variant divides by n (literal). */
int orc_correl_normalized(const double *d1, size_t n1, const double *d2, size_t n2, int fast, double *ans)
{
    const size_t n = n1;
    size_t i, lag;
    if (n == 0) return -1;
    if (n2 != n) return -4;
    double s1 = 0, s2 = 0, q1 = 0, q2 = 0;
    for (i = 0; i < n; ++i) { s1 += d1[i]; s2 += d2[i]; }
    const double m1 = s1 / (double)n, m2 = s2 / (double)n;
    double sd1, sd2;
    if (fast) {
        for (i = 0; i < n; ++i) { q1 += d1[i] * d1[i]; q2 += d2[i] * d2[i]; }
        sd1 = sqrt(q1 / (double)n - m1 * m1);
        sd2 = sqrt(q2 / (double)n - m2 * m2);
    } else {
        for (i = 0; i < n; ++i) { q1 += (d1[i] - m1) * (d1[i] - m1); q2 += (d2[i] - m2) * (d2[i] - m2); }
        sd1 = sqrt(q1 / (double)n);
        sd2 = sqrt(q2 / (double)n);
    }
    if (sd1 == 0.0 || sd2 == 0.0) return -8;      /* Correlation.rs:214-216, :246-248 (a NaN std passes, literal) */
    if (fast && n <= 32) {
        const double f = 1.0 / (sd1 * sd2 * (double)n);
        for (lag = 0; lag < n; ++lag) {
            double sum = 0.0;
            for (i = 0; i < n - lag; ++i) sum += (d1[i + lag] - m1) * (d2[i] - m2);
            ans[lag] = sum * f;
        }
        return 0;
    }
    double *a = (double *)malloc(sizeof(double) * n), *b = (double *)malloc(sizeof(double) * n);
    for (i = 0; i < n; ++i) { a[i] = (d1[i] - m1) / sd1; b[i] = (d2[i] - m2) / sd2; }
    const int rc = orc_correl(a, n, b, n, ans);
    free(a); free(b);
    return rc;
}

/* autocorrel_fast -- Correlation.rs:286-323.  n <= 32: the direct loop (literal).  n > 32: NR intent
 * (deviation D10): the literal code feeds |F|^2 to the placeholder inverse DFT in that placeholder's own
 * layout and without the 1/no2 factor; the intent is the autocorrelation, i.e. correl(data, data). */
int orc_autocorrel_fast(const double *d, size_t n, double *ans) { return orc_correl(d, n, d, n, ans); }

/* ==========================================================================================
 * N3 of SURVEY.md 8f: cosft1 / cosft2 / sinft around realft.  The reference's bodies are NR
 * transliterations that end in `unimplemented!()` (Cos_FT.rs:70-74, Cos_FT2.rs realft stub), so these are
 * NR intent (deviation D11), with the reference's 1-based calling convention kept: y[0] is unused,
 * the data are y[1..=n+1] (cosft1, Cos_FT.rs:7-67: `y[n + 1]` is read and written) resp. y[1..=n]
 * (cosft2, Cos_FT2.rs:7-13; sinft, README.md:72 -- listed, no source file).  Twiddles by NR's recurrences, as
 * the reference (Cos_FT.rs:9-12,27-34; Cos_FT2.rs:17-21,40-44).  Places where the reference departs from NR:
 *   cosft1: `y[2] = sum` after `y[n+1] = y[2]` is missing (Cos_FT.rs:60-61);
 *   cosft2: the first loop uses a constant wi1 (Cos_FT2.rs:26-36 updates it only afterwards, :39-44), the
 *           second loop's scan adds wr1 / wi1 instead of the state (Cos_FT2.rs:58-59), the inverse rotation has
 *           both outputs built from y[i+1]*wr (Cos_FT2.rs:125-126) and zero increments (:112-113).
 * ======================================================================================== */
void orc_cosft1(double *y, size_t n)
{
    const double theta = ORC_PI / (double)n;
    double wtemp = sin(0.5 * theta);
    const double wpr = -2.0 * wtemp * wtemp, wpi = sin(theta);
    double wr = 1.0, wi = 0.0;
    const size_t n2 = n + 2;
    size_t j;
    double sum = 0.5 * (y[1] - y[n + 1]);                        /* Cos_FT.rs:17-18 */
    y[1] = 0.5 * (y[1] + y[n + 1]);
    for (j = 2; j <= (n >> 1); ++j) {                            /* Cos_FT.rs:36-52 */
        wtemp = wr;
        wr = wr * wpr - wi * wpi + wr;
        wi = wi * wpr + wtemp * wpi + wi;
        const double y1 = 0.5 * (y[j] + y[n2 - j]);
        const double y2 = y[j] - y[n2 - j];
        y[j] = y1 - wi * y2;
        y[n2 - j] = y1 + wi * y2;
        sum += wr * y2;
    }
    orc_realft(y + 1, n, 1);                                     /* Cos_FT.rs:58 */
    y[n + 1] = y[2];                                             /* Cos_FT.rs:61 */
    y[2] = sum;                                                  /* NR; missing in the reference */
    for (j = 4; j <= n; j += 2) { sum += y[j]; y[j] = sum; }     /* Cos_FT.rs:64-67 */
}

int orc_cosft2(double *y, size_t n, int isign)
{
    if (isign != 1 && isign != -1) return -3;                    /* Cos_FT2.rs:11 panics */
    const double theta = 0.5 * ORC_PI / (double)n;
    double wr1 = cos(theta), wi1 = sin(theta), wr = 1.0, wi = 0.0, wtemp;
    const double wpr = -2.0 * wi1 * wi1, wpi = sin(2.0 * theta);
    size_t i;
    if (isign == 1) {
        for (i = 1; i <= n / 2; ++i) {                           /* Cos_FT2.rs:26-36 (+ NR's update inside) */
            const double y1 = 0.5 * (y[i] + y[n - i + 1]);
            const double y2 = wi1 * (y[i] - y[n - i + 1]);
            y[i] = y1 + y2;
            y[n - i + 1] = y1 - y2;
            wtemp = wr1;
            wr1 = wr1 * wpr - wi1 * wpi + wr1;
            wi1 = wi1 * wpr + wtemp * wpi + wi1;
        }
        orc_realft(y + 1, n, 1);                                 /* Cos_FT2.rs:48 */
        for (i = 3; i <= n; i += 2) {                            /* Cos_FT2.rs:54-77 */
            wtemp = wr;
            wr = wr * wpr - wi * wpi + wr;
            wi = wi * wpr + wtemp * wpi + wi;
            const double y1 = y[i] * wr - y[i + 1] * wi;
            const double y2 = y[i + 1] * wr + y[i] * wi;
            y[i] = y1;
            y[i + 1] = y2;
        }
        double sum = 0.5 * y[2];                                 /* Cos_FT2.rs:80-85 */
        for (i = n; i >= 2; i -= 2) {
            const double sum1 = sum;
            sum += y[i];
            y[i] = sum1;
        }
    } else {
        const double ytemp = y[n];                               /* Cos_FT2.rs:97-101 */
        for (i = n; i >= 4; i -= 2) y[i] = y[i - 2] - y[i];
        y[2] = 2.0 * ytemp;
        for (i = 3; i <= n; i += 2) {                            /* Cos_FT2.rs:108-131 (NR rotation) */
            wtemp = wr;
            wr = wr * wpr - wi * wpi + wr;
            wi = wi * wpr + wtemp * wpi + wi;
            const double y1 = y[i] * wr + y[i + 1] * wi;
            const double y2 = y[i + 1] * wr - y[i] * wi;
            y[i] = y1;
            y[i + 1] = y2;
        }
        orc_realft(y + 1, n, -1);                                /* Cos_FT2.rs:134 */
        for (i = 1; i <= n / 2; ++i) {                           /* Cos_FT2.rs:142-162 */
            const double y1 = y[i] + y[n - i + 1];
            const double y2 = (0.5 / wi1) * (y[i] - y[n - i + 1]);
            y[i] = 0.5 * (y1 + y2);
            y[n - i + 1] = 0.5 * (y1 - y2);
            wtemp = wr1;
            wr1 = wr1 * wpr - wi1 * wpi + wr1;
            wi1 = wi1 * wpr + wtemp * wpi + wi1;
        }
    }
    return 0;
}

/* NR sinft; the reference lists it (README.md:72) but has no source file for it */
void orc_sinft(double *y, size_t n)
{
    const double theta = ORC_PI / (double)n;
    double wtemp = sin(0.5 * theta);
    const double wpr = -2.0 * wtemp * wtemp, wpi = sin(theta);
    double wr = 1.0, wi = 0.0, sum;
    const size_t n2 = n + 2;
    size_t j;
    y[1] = 0.0;
    for (j = 2; j <= (n >> 1) + 1; ++j) {
        wtemp = wr;
        wr = wr * wpr - wi * wpi + wr;
        wi = wi * wpr + wtemp * wpi + wi;
        const double y1 = wi * (y[j] + y[n2 - j]);
        const double y2 = 0.5 * (y[j] - y[n2 - j]);
        y[j] = y1 + y2;
        y[n2 - j] = y1 - y2;
    }
    orc_realft(y + 1, n, 1);
    y[1] *= 0.5;
    sum = y[2] = 0.0;
    for (j = 1; j <= n - 1; j += 2) {
        sum += y[j];
        y[j] = y[j + 1];
        y[j + 1] = sum;
    }
}
