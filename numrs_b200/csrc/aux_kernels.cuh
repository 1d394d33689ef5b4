// aux_kernels.cuh -- elementwise kernels around the FFT passes (all HBM/L2-bound streaming).
//   untangle      : standalone NR realft untangling for lines longer than one CTA tile
//                   (Real_FT.rs:49-80,145-176) and for N == 1
//   spectral      : packed-spectrum multiply / divide / conj-multiply with the 1/no2 scale
//                   (Convolve.rs:96,112-129; Correlation.rs:73-74,91-92)
//   pad_response  : response placement (Convolve.rs:41-63 literal, or NR wrap-around)
//   correl_direct : linear-lag direct correlation for n <= 32 (Correlation.rs:37-50)
// Bodies are written grid-stride over a flat item index so the CUDA wrapper and the host
// emulation used by the tests share them.
#pragma once
#include <math.h>

#include "fft_pass.cuh"

namespace nrb {

NRB_DEV double2 two_level_tw(const double2 *lo, const double2 *hi, int h, u64 m)
{
    return cmul(NRB_LDG(lo + (m & ((1ull << h) - 1ull))), NRB_LDG(hi + (m >> h)));
}

// items: count * max(N/2, 1); item (line, k).  In/out may alias (each pair is owned by one item).
NRB_DEV void aux_untangle(const AuxParams &A, u64 gtid, u64 gthreads)
{
    const u64 N = A.n;
    const u64 half = N >= 2 ? N / 2 : 1;
    const u64 items = A.count * half;
    for (u64 it = gtid; it < items; it += gthreads) {
        const u64 k = it % half, line = it / half;
        const double2 *src = A.a + (i64)line * A.a_stride;
        double2 *dst = A.out + (i64)line * A.out_stride;
        if (k == 0) {
            const double2 g0 = src[0];
            const double2 mid = N >= 2 ? src[N / 2] : make_double2(0.0, 0.0);
            if (A.dir > 0) {
                const double f0 = g0.x + g0.y, fn = g0.x - g0.y;
                if (A.op == REAL_SPEQ) { dst[0] = make_double2(f0, 0.0); A.speq[line] = make_double2(fn, 0.0); }
                else dst[0] = make_double2(f0, fn);
            } else {
                if (A.op == REAL_SPEQ) dst[0] = dc_inverse_speq(g0, A.speq[line]);
                else dst[0] = make_double2(0.5 * (g0.x + g0.y), 0.5 * (g0.x - g0.y));
            }
            if (N >= 2) dst[N / 2] = mid;
        } else {
            const double2 a = NRB_LDS(src + k), b = NRB_LDS(src + (N - k));
            const double2 t = two_level_tw(A.rtw_lo, A.rtw_hi, A.rtw_h, k);
            double2 oa, ob;
            if (A.dir > 0) untangle_pair<1>(a, b, t, oa, ob);
            else untangle_pair<-1>(a, b, t, oa, ob);
            dst[k] = oa;
            dst[N - k] = ob;
        }
    }
}

// packed spectra of real length n: element 0 = (F_0, F_{n/2}) both real, element k = F_k.
// items: count * n/2.  out may alias a.
NRB_DEV void aux_spectral(const AuxParams &A, u64 gtid, u64 gthreads)
{
    const u64 half = A.n / 2;
    const u64 items = A.count * half;
    const double inv = 1.0 / (double)half;      // 1/no2; n is a power of two -> exact
    for (u64 it = gtid; it < items; it += gthreads) {
        const u64 k = it % half, sig = it / half;
        const double2 d = NRB_LDS(A.a + (i64)sig * A.a_stride + (i64)k);
        double2 r = d;
        if (A.op != SPEC_AUTOCORREL) r = A.b_stride ? NRB_LDS(A.b + (i64)sig * A.b_stride + (i64)k) : NRB_LDG(A.b + (i64)k);
        double2 o;
        if (A.op == SPEC_CONV_MUL) {
            if (k == 0) o = make_double2(d.x * r.x * inv, d.y * r.y * inv);
            else o = make_double2((d.x * r.x - d.y * r.y) * inv, (d.x * r.y + d.y * r.x) * inv);
        } else if (A.op == SPEC_AUTOCORREL) {   // |F|^2 (Correlation.rs:306-313): DC and Nyquist real squares
            if (k == 0) o = make_double2(d.x * d.x * inv, d.y * d.y * inv);
            else o = make_double2((d.x * d.x + d.y * d.y) * inv, 0.0);
        } else if (A.op == SPEC_CORREL) {
            if (k == 0) o = make_double2(d.x * r.x * inv, d.y * r.y * inv);
            else o = make_double2((d.x * r.x + d.y * r.y) * inv, (d.y * r.x - d.x * r.y) * inv);
        } else { // SPEC_CONV_DIV, guard mag2 < 1e-12 -> 0 (Convolve.rs:118-122)
            if (k == 0) {
                const double m0 = r.x * r.x, m1 = r.y * r.y;
                o.x = m0 < 1e-12 ? 0.0 : d.x * r.x / m0 * inv;
                o.y = m1 < 1e-12 ? 0.0 : d.y * r.y / m1 * inv;
            } else {
                const double mag2 = r.x * r.x + r.y * r.y;
                if (mag2 < 1e-12) o = make_double2(0.0, 0.0);
                else o = make_double2((d.x * r.x + d.y * r.y) / mag2 * inv, (d.y * r.x - d.x * r.y) / mag2 * inv);
            }
        }
        A.out[(i64)sig * A.out_stride + (i64)k] = o;
    }
}

// Fused spectral step for transforms whose real untangling is a separate pass (lines longer than a CTA
// tile): a = raw c2c output Z of the data (N = n/2 complex per signal), b = either the packed, already
// untangled response spectrum shared by all signals (b_stride == 0, convlv) or the raw c2c output of the
// second signal (b_stride != 0, correl).  Per pair (k, N-k): forward untangle (Real_FT.rs:49-80), the
// spectral op with the 1/no2 scale (Convolve.rs:112-129 / Correlation.rs:91-92), inverse untangle
// (Real_FT.rs:145-176) -- out is ready for the inverse c2c.  Saves two (convlv) / three (correl) full
// passes over the data.  items: count * max(N/2, 1).
NRB_DEV double2 spectral_op(int op, double2 d, double2 r, double inv)
{
    if (op == SPEC_CONV_MUL) return make_double2((d.x * r.x - d.y * r.y) * inv, (d.x * r.y + d.y * r.x) * inv);
    if (op == SPEC_CORREL) return make_double2((d.x * r.x + d.y * r.y) * inv, (d.y * r.x - d.x * r.y) * inv);
    const double mag2 = r.x * r.x + r.y * r.y;
    if (mag2 < 1e-12) return make_double2(0.0, 0.0);
    return make_double2((d.x * r.x + d.y * r.y) / mag2 * inv, (d.y * r.x - d.x * r.y) / mag2 * inv);
}
NRB_DEV double spectral_op_real(int op, double d, double r, double inv)
{
    if (op == SPEC_CONV_DIV) { const double m = r * r; return m < 1e-12 ? 0.0 : d * r / m * inv; }
    return d * r * inv;
}

NRB_DEV void aux_spectral_z(const AuxParams &A, u64 gtid, u64 gthreads)
{
    const u64 N = A.n / 2;
    const u64 half = N >= 2 ? N / 2 : 1;
    const u64 items = A.count * half;
    const double inv = 1.0 / (double)N;
    const bool self = A.op == SPEC_AUTOCORREL;      // second operand = the first one (|F|^2)
    const bool b_raw = A.b_stride != 0 && !self;
    const int op = self ? SPEC_CORREL : A.op;
    for (u64 it = gtid; it < items; it += gthreads) {
        const u64 k = it % half, sig = it / half;
        const double2 *za = A.a + (i64)sig * A.a_stride;
        const double2 *zb = self ? za : A.b + (i64)sig * A.b_stride;
        double2 *out = A.out + (i64)sig * A.out_stride;
        if (k == 0) {
            const double2 a0 = NRB_LDS(za);
            double2 r0 = self ? a0 : (b_raw ? NRB_LDS(zb) : NRB_LDG(zb));
            if (b_raw || self) r0 = make_double2(r0.x + r0.y, r0.x - r0.y);       // (B_0, B_N)
            const double g0 = spectral_op_real(op, a0.x + a0.y, r0.x, inv);
            const double gn = spectral_op_real(op, a0.x - a0.y, r0.y, inv);
            out[0] = make_double2(0.5 * (g0 + gn), 0.5 * (g0 - gn));
            if (N >= 2) {   // middle bin: untangling is the identity there
                const double2 am = NRB_LDS(za + N / 2);
                const double2 rm = self ? am : (b_raw ? NRB_LDS(zb + N / 2) : NRB_LDG(zb + N / 2));
                out[N / 2] = spectral_op(op, am, rm, inv);
            }
        } else {
            const double2 t = two_level_tw(A.rtw_lo, A.rtw_hi, A.rtw_h, k);
            double2 fa, fm, ra, rm;
            untangle_pair<1>(NRB_LDS(za + k), NRB_LDS(za + (N - k)), t, fa, fm);
            if (self) { ra = fa; rm = fm; }
            else if (b_raw) untangle_pair<1>(NRB_LDS(zb + k), NRB_LDS(zb + (N - k)), t, ra, rm);
            else { ra = NRB_LDG(zb + k); rm = NRB_LDG(zb + (N - k)); }
            const double2 ga = spectral_op(op, fa, ra, inv), gm = spectral_op(op, fm, rm, inv);
            double2 oa, ob;
            untangle_pair<-1>(ga, gm, t, oa, ob);
            out[k] = oa;
            out[N - k] = ob;
        }
    }
}

// The same fused step for spectra left in the TRANSPOSED order of a two-pass transform without transposition:
// N = F * REST complex points (F = 2^m), bin k = kf + F*kr sits at position kf*REST + kr (what "strided F-point
// pass + twiddle, then contiguous REST-point pass" produces from natural-order input, and what the mirrored
// inverse consumes).  A spectrum that is only an intermediate -- convlv / correl -- never needs natural order,
// which saves one HBM pass per transform over the three-factor natural-order plan.
// Partner of (kf, kr): (F - kf, REST - 1 - kr) for kf != 0, (0, (REST - kr) mod REST) for kf = 0; an item owns one
// pair.  Items are enumerated row by row, (kf in [0, F/2], kr in [0, REST)), so that a warp reads row kf forwards
// and row F - kf backwards: count * (F/2 + 1) * REST items, the surplus ones of rows 0 and F/2 are skipped.
// b: the raw transposed c2c output of the second operand -- per signal (b_stride != 0, correl) or one shared by all
// signals (b_stride == 0 and dir == 1: the response of convlv, kept raw so that its reads are coalesced like a's);
// b_stride == 0 and dir == 0: a packed, untangled spectrum in NATURAL order; SPEC_AUTOCORREL: the first operand.
NRB_DEV void aux_spectral_zt(const AuxParams &A, u64 gtid, u64 gthreads)
{
    const u64 N = A.n / 2;
    const int f = (int)A.m;
    const u64 F = 1ull << f, REST = N >> f;
    const u64 rows = F / 2 + 1, per = rows * REST, items = A.count * per;
    const double inv = 1.0 / (double)N;
    const bool self = A.op == SPEC_AUTOCORREL;
    const bool b_raw = (A.b_stride != 0 || A.dir == 1) && !self;
    const int op = self ? SPEC_CORREL : A.op;
    for (u64 it = gtid; it < items; it += gthreads) {
        const u64 sig = it / per, w = it % per, kf = w / REST, kr = w % REST;
        const double2 *za = A.a + (i64)sig * A.a_stride;
        const double2 *zb = self ? za : A.b + (i64)sig * A.b_stride;
        double2 *out = A.out + (i64)sig * A.out_stride;
        if (kf == 0 && kr == 0) {                           // k = 0: DC and Nyquist share the element
            const double2 a0 = NRB_LDS(za);
            double2 r0 = self ? a0 : (b_raw ? NRB_LDS(zb) : NRB_LDG(zb));
            if (b_raw || self) r0 = make_double2(r0.x + r0.y, r0.x - r0.y);
            const double g0 = spectral_op_real(op, a0.x + a0.y, r0.x, inv);
            const double gn = spectral_op_real(op, a0.x - a0.y, r0.y, inv);
            out[0] = make_double2(0.5 * (g0 + gn), 0.5 * (g0 - gn));
            continue;
        }
        if (kf == 0 && kr == REST / 2) {                    // k = N/2: untangling is the identity
            const double2 am = NRB_LDS(za + kr);
            const double2 rm = self ? am : (b_raw ? NRB_LDS(zb + kr) : NRB_LDG(zb + N / 2));
            out[kr] = spectral_op(op, am, rm, inv);
            continue;
        }
        u64 pkf, pkr;                                       // partner
        if (kf == 0) { if (kr > REST / 2) continue; pkf = 0; pkr = REST - kr; }
        else if (2 * kf == F) { if (kr >= REST / 2) continue; pkf = kf; pkr = REST - 1 - kr; }
        else { pkf = F - kf; pkr = REST - 1 - kr; }
        const u64 k = kf + F * kr;                          // any 1 <= k <= N-1 works in the pair formula
        const u64 pa = kf * REST + kr, pm = pkf * REST + pkr;
        const double2 t = two_level_tw(A.rtw_lo, A.rtw_hi, A.rtw_h, k);
        double2 fa, fm, ra, rm;
        untangle_pair<1>(NRB_LDS(za + pa), NRB_LDS(za + pm), t, fa, fm);
        if (self) { ra = fa; rm = fm; }
        else if (b_raw) untangle_pair<1>(NRB_LDS(zb + pa), NRB_LDS(zb + pm), t, ra, rm);
        else { ra = NRB_LDG(zb + k); rm = NRB_LDG(zb + (N - k)); }
        const double2 ga = spectral_op(op, fa, ra, inv), gm = spectral_op(op, fm, rm, inv);
        double2 oa, ob;
        untangle_pair<-1>(ga, gm, t, oa, ob);
        out[pa] = oa;
        out[pm] = ob;
    }
}

// a = response taps (m doubles), out = padded response (n doubles).  items: n.
NRB_DEV void aux_pad_response(const AuxParams &A, u64 gtid, u64 gthreads)
{
    const double *r = reinterpret_cast<const double *>(A.a);
    double *p = reinterpret_cast<double *>(A.out);
    const u64 n = A.n, m = A.m;
    for (u64 i = gtid; i < n; i += gthreads) {
        double v = 0.0;
        if (A.op == 0) {
            const u64 mid = (m - 1) / 2;
            if (i < mid) v = r[m - mid + i];
            else if (i < m) v = r[i];
            else if (i < n - mid) v = 0.0;
            else v = r[i - (n - mid)];
        } else {
            const u64 half = (m - 1) / 2;
            if (i < (m + 1) / 2) v = r[i];
            else if (i >= n - half) v = r[m - (n - i)];
        }
        p[i] = v;
    }
}

// a, b = signals (count x n doubles), out = count x n doubles; n <= 32.  items: count * n (lag).
NRB_DEV void aux_correl_direct(const AuxParams &A, u64 gtid, u64 gthreads)
{
    const double *d1 = reinterpret_cast<const double *>(A.a);
    const double *d2 = reinterpret_cast<const double *>(A.b) + A.m;   // m: extra offset of b in doubles
    double *out = reinterpret_cast<double *>(A.out);
    const u64 n = A.n, items = A.count * n;
    for (u64 it = gtid; it < items; it += gthreads) {
        const u64 lag = it % n, sig = it / n;
        const double *x = d1 + (i64)sig * A.a_stride, *y = d2 + (i64)sig * A.b_stride;
        double sum = 0.0;
        for (u64 i = 0; i + lag < n; ++i) sum += x[i + lag] * y[i];
        out[(i64)sig * A.out_stride + (i64)lag] = sum;
    }
}

// u(i) = splitmix64(seed * 0x9E3779B97F4A7C15 + offset + i), x = (u >> 11) * 2^-52 - 1 in [-1, 1)
NRB_DEV void aux_fill(const AuxParams &A, u64 gtid, u64 gthreads)
{
    double *out = reinterpret_cast<double *>(A.out);
    const u64 base = A.m * 0x9E3779B97F4A7C15ull + A.count;
    for (u64 i = gtid; i < A.n; i += gthreads) {
        u64 x = base + i + 0x9E3779B97F4A7C15ull;
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
        x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
        x = x ^ (x >> 31);
        out[i] = (double)(x >> 11) * (1.0 / 4503599627370496.0) - 1.0;
    }
}


// ------------------------------------------------------------------ "next" rows
// Signals of a two-input launch: s < count/2 comes from a, the rest from b (b == nullptr: all from a).
NRB_DEV const double *signal_src(const AuxParams &A, u64 s)
{
    const double *a = reinterpret_cast<const double *>(A.a), *b = reinterpret_cast<const double *>(A.b);
    if (!b) return a + (i64)s * A.a_stride;
    const u64 half = A.count / 2;
    return s < half ? a + (i64)s * A.a_stride : b + (i64)(s - half) * A.b_stride;
}

// Deterministic strided partial sums: accumulator (s, c), c < C = A.m, adds the terms i = c, c + C, ... of
// signal s, so consecutive threads read consecutive addresses.  Terms (op): RED_SUM_SQ (x, x^2) of raw data
// (Correlation.rs:236-239), RED_CENTERED_SQ ((x - mean)^2, 0) with the mean from the stats array in `speq`
// (Correlation.rs:207), RED_PARTIALS a previous level's double2 partials.  out[s*C + c].  items: count * C.
NRB_DEV void aux_reduce(const AuxParams &A, u64 gtid, u64 gthreads)
{
    const u64 C = A.m, items = A.count * C;
    for (u64 it = gtid; it < items; it += gthreads) {
        const u64 c = it % C, s = it / C;
        double sx = 0.0, sy = 0.0;
        if (A.op == RED_PARTIALS) {
            const double2 *p = A.a + (i64)(s * A.n);
            for (u64 i = c; i < A.n; i += C) { const double2 v = p[i]; sx += v.x; sy += v.y; }
        } else {
            const double *x = signal_src(A, s);
            const double mean = A.op == RED_CENTERED_SQ ? A.speq[s].x : 0.0;
            // 16-byte loads when every signal starts on a 16-byte boundary (n even and aligned bases), four
            // independent accumulator pairs for latency; same terms, fixed order -> deterministic
            const bool vec = (A.n % 2 == 0) && (reinterpret_cast<size_t>(x) % 16 == 0);
            if (vec) {
                const double2 *x2 = reinterpret_cast<const double2 *>(x);
                const u64 n2 = A.n / 2;
                double ax[4] = {0.0, 0.0, 0.0, 0.0}, ay[4] = {0.0, 0.0, 0.0, 0.0};
                u64 i = c;
                for (; i + 3 * C < n2; i += 4 * C) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const double2 v = NRB_LDS(x2 + i + (u64)u * C);
                        if (A.op == RED_SUM_SQ) { ax[u] += v.x + v.y; ay[u] += v.x * v.x + v.y * v.y; }
                        else { const double a = v.x - mean, b = v.y - mean; ax[u] += a * a + b * b; }
                    }
                }
                for (; i < n2; i += C) {
                    const double2 v = NRB_LDS(x2 + i);
                    if (A.op == RED_SUM_SQ) { ax[0] += v.x + v.y; ay[0] += v.x * v.x + v.y * v.y; }
                    else { const double a = v.x - mean, b = v.y - mean; ax[0] += a * a + b * b; }
                }
                sx = (ax[0] + ax[1]) + (ax[2] + ax[3]);
                sy = (ay[0] + ay[1]) + (ay[2] + ay[3]);
            } else if (A.op == RED_SUM_SQ) {
                for (u64 i = c; i < A.n; i += C) { const double v = x[i]; sx += v; sy += v * v; }
            } else {
                for (u64 i = c; i < A.n; i += C) { const double v = x[i] - mean; sx += v * v; }
            }
        }
        A.out[it] = make_double2(sx, sy);
    }
}

// stats[s] = (mean, std) from the m partials of signal s; n = signal length.  items: count.
NRB_DEV void aux_stats_final(const AuxParams &A, u64 gtid, u64 gthreads)
{
    for (u64 s = gtid; s < A.count; s += gthreads) {
        double sx = 0.0, sy = 0.0;
        for (u64 j = 0; j < A.m; ++j) { const double2 v = A.a[s * A.m + j]; sx += v.x; sy += v.y; }
        const double n = (double)A.n;
        double2 st = A.out[s];
        if (A.op == STATS_FAST) { st.x = sx / n; st.y = sqrt(sy / n - st.x * st.x); }   // Correlation.rs:241-244
        else if (A.op == STATS_MEAN) { st.x = sx / n; st.y = 0.0; }                     // Correlation.rs:200
        else if (A.op == STATS_SUM) { st.x = sx; st.y = sy; }
        else st.y = sqrt(sx / n);                                                       // Correlation.rs:207
        A.out[s] = st;
    }
}

// out[s][i] = (x[s][i] - mean_s) / std_s, stats in `speq`.  items: count * n/2 pairs when n is even and the
// lines are 16-byte aligned (op 1, chosen by the launcher), else count * n.
NRB_DEV void aux_normalize(const AuxParams &A, u64 gtid, u64 gthreads)
{
    double *out = reinterpret_cast<double *>(A.out);
    if (A.op == 1) {
        // (the kernel was bound by its two 64-bit integer divisions and two FP64 divisions per 32 bytes, not by HBM: a shift
        // when n is a power of two -- always on the FFT path --, and one reciprocal per item: <= 1 ulp from the quotient)
        const u64 n2 = A.n / 2, items = A.count * n2;
        int sh = -1;
        if ((n2 & (n2 - 1)) == 0) { sh = 0; while ((1ull << sh) < n2) ++sh; }
        for (u64 it = gtid; it < items; it += gthreads) {
            const u64 s = sh >= 0 ? it >> sh : it / n2, i = it - s * n2;
            const double2 st = A.speq[s];
            const double inv = 1.0 / st.y;
            const double2 v = NRB_LDS(reinterpret_cast<const double2 *>(signal_src(A, s)) + i);
            reinterpret_cast<double2 *>(out + (i64)s * A.out_stride)[i] = make_double2((v.x - st.x) * inv, (v.y - st.x) * inv);
        }
        return;
    }
    const u64 items = A.count * A.n;
    for (u64 it = gtid; it < items; it += gthreads) {
        const u64 i = it % A.n, s = it / A.n;
        const double2 st = A.speq[s];
        out[(i64)s * A.out_stride + (i64)i] = (signal_src(A, s)[i] - st.x) / st.y;
    }
}

// out[i] = |a[i]|^2 (op 0, FFT_1.rs:218-228) or |a[i]| (op 1, FFT_1.rs:206-216).  items: n complex points.
NRB_DEV void aux_power(const AuxParams &A, u64 gtid, u64 gthreads)
{
    double *out = reinterpret_cast<double *>(A.out);
    for (u64 i = gtid; i < A.n; i += gthreads) {
        const double2 z = NRB_LDS(A.a + i);
        const double p = z.x * z.x + z.y * z.y;
        out[i] = A.op ? sqrt(p) : p;
    }
}

// twofft packing (FFT_2.rs:33-37): out[s][j] = (d1[s][j], d2[s][j]).  Strides of a / b in doubles, of out in
// complex elements.  items: count * n.
NRB_DEV void aux_pack2(const AuxParams &A, u64 gtid, u64 gthreads)
{
    const double *d1 = reinterpret_cast<const double *>(A.a), *d2 = reinterpret_cast<const double *>(A.b);
    const u64 items = A.count * A.n;
    for (u64 it = gtid; it < items; it += gthreads) {
        const u64 j = it % A.n, s = it / A.n;
        A.out[(i64)s * A.out_stride + (i64)j] = make_double2(d1[(i64)s * A.a_stride + (i64)j], d2[(i64)s * A.b_stride + (i64)j]);
    }
}

// twofft separation (FFT_2.rs:53-90 with the 0-based mirror n - k, ledger D9): a = transform of the packed
// signal (n complex per signal, stride a_stride), out = fft1, speq = fft2 (both stride out_stride >= n + 1;
// element n is the reference's two extra doubles, FFT_2.rs:6-7, set to 0).  a may alias out when the strides
// agree: item k owns bins k and n - k.  items: count * (n/2 + 1).
NRB_DEV void aux_twofft_split(const AuxParams &A, u64 gtid, u64 gthreads)
{
    const u64 n = A.n, per = n / 2 + 1, items = A.count * per;
    for (u64 it = gtid; it < items; it += gthreads) {
        const u64 k = it % per, s = it / per;
        const double2 *f = A.a + (i64)s * A.a_stride;
        double2 *f1 = A.out + (i64)s * A.out_stride, *f2 = A.speq + (i64)s * A.out_stride;
        if (k == 0) {
            const double2 z = f[0];
            f1[0] = make_double2(z.x, 0.0);
            f2[0] = make_double2(z.y, 0.0);
            f1[n] = make_double2(0.0, 0.0);
            f2[n] = make_double2(0.0, 0.0);
        } else {
            const u64 m = n - k;
            const double2 a = f[k], b = f[m];
            const double rep = 0.5 * (a.x + b.x), rem = 0.5 * (a.x - b.x);
            const double aip = 0.5 * (a.y + b.y), aim = 0.5 * (a.y - b.y);
            f1[k] = make_double2(rep, aim);
            f2[k] = make_double2(aip, -rem);
            if (m != k) { f1[m] = make_double2(rep, -aim); f2[m] = make_double2(aip, rem); }
        }
    }
}

// out[i] *= 1/m (the 1/n of correl_normalized_fast's direct branch, Correlation.rs:252).  items: n doubles.
NRB_DEV void aux_scale(const AuxParams &A, u64 gtid, u64 gthreads)
{
    double *out = reinterpret_cast<double *>(A.out);
    const double f = 1.0 / (double)A.m;
    for (u64 i = gtid; i < A.n; i += gthreads) out[i] *= f;
}

// ------------------------------------------------------------------ cosft1 / cosft2 / sinft
// The reference's routines work on 1-based arrays (y[0] unused): line l of the io buffer starts at
// io + l*ld doubles and its data are f(i) = io[l*ld + 1 + i].  G is the realft work array (n/2 complex per line).
// exp(-i pi m / n) from the two-level table of M = 2n; COS2 modes use M = 4n, i.e. exp(-i pi m / (2n)).

// Pre-processing (and cosft2's inverse post-processing): m accumulators per line, accumulator c handles the
// pairs j = c, c + m, ... (consecutive threads -> consecutive addresses) and, for cosft1, adds up its share
// of sum = 0.5 (f_0 - f_n) + sum_j cos(j pi/n) (f_j - f_{n-j})  (Cos_FT.rs:17,36-55) into b[l*m + c].
// a = io (doubles, stride a_stride = ld), out = G, except COS2I_POST: a = G, out = io (stride out_stride = ld).
NRB_DEV void aux_cosft(const AuxParams &A, u64 gtid, u64 gthreads)
{
    const u64 n = A.n, N = n / 2, C = A.m, items = A.count * C;
    const int mode = A.dir;
    for (u64 it = gtid; it < items; it += gthreads) {
        const u64 c = it % C, l = it / C;
        double acc = 0.0;
        if (mode == COS2I_POST) {                                          // Cos_FT2.rs:142-162
            const double *g = reinterpret_cast<const double *>(A.a + (i64)(l * N));
            double *f = reinterpret_cast<double *>(A.out) + (i64)l * A.out_stride + 1;
            for (u64 i = c; i < N; i += C) {
                const double2 t = two_level_tw(A.rtw_lo, A.rtw_hi, A.rtw_h, 2 * i + 1);   // sin((2i+1) pi/(2n)) = -t.y
                const double gi = g[i], gm = g[n - 1 - i];
                const double y1 = gi + gm, y2 = (0.5 / -t.y) * (gi - gm);
                f[i] = 0.5 * (y1 + y2);
                f[n - 1 - i] = 0.5 * (y1 - y2);
            }
            continue;
        }
        const double *f = reinterpret_cast<const double *>(A.a) + (i64)l * A.a_stride + 1;
        double *g = reinterpret_cast<double *>(A.out + (i64)(l * N));
        if (mode == COS1) {
            for (u64 j = c; j <= N; j += C) {
                if (j == 0) { g[0] = 0.5 * (f[0] + f[n]); acc += 0.5 * (f[0] - f[n]); }
                else if (j == N) g[N] = f[N];
                else {
                    const double2 t = two_level_tw(A.rtw_lo, A.rtw_hi, A.rtw_h, j);      // (cos, -sin)(j pi/n)
                    const double y1 = 0.5 * (f[j] + f[n - j]), y2 = f[j] - f[n - j];
                    g[j] = y1 + t.y * y2;
                    g[n - j] = y1 - t.y * y2;
                    acc += t.x * y2;
                }
            }
            if (A.b) const_cast<double2 *>(A.b)[it] = make_double2(acc, 0.0);
        } else if (mode == COS2F) {                                        // Cos_FT2.rs:26-36
            for (u64 i = c; i < N; i += C) {
                const double2 t = two_level_tw(A.rtw_lo, A.rtw_hi, A.rtw_h, 2 * i + 1);
                const double y1 = 0.5 * (f[i] + f[n - 1 - i]), y2 = -t.y * (f[i] - f[n - 1 - i]);
                g[i] = y1 + y2;
                g[n - 1 - i] = y1 - y2;
            }
        } else if (mode == SINFT) {                                        // NR sinft, first loop
            for (u64 j = c; j <= N; j += C) {
                if (j == 0) { g[0] = 0.0; continue; }
                const double2 t = two_level_tw(A.rtw_lo, A.rtw_hi, A.rtw_h, j);
                const double y1 = -t.y * (f[j] + f[n - j]), y2 = 0.5 * (f[j] - f[n - j]);
                g[j] = y1 + y2;
                if (j != N) g[n - j] = y1 - y2;
            }
        } else {                                                           // COS2I_PRE, Cos_FT2.rs:97-131
            for (u64 k = c; k < N; k += C) {
                if (k == 0) { g[0] = f[0]; g[1] = 2.0 * f[n - 1]; continue; }
                const double2 t = two_level_tw(A.rtw_lo, A.rtw_hi, A.rtw_h, 2 * k);      // (cos, -sin)(k pi/n)
                const double re = f[2 * k], im = f[2 * k - 1] - f[2 * k + 1];
                g[2 * k] = re * t.x - im * t.y;
                g[2 * k + 1] = im * t.x + re * t.y;
            }
        }
    }
}

// The term of position pos (0 <= pos < N) in the running sum, and the packed-spectrum value that goes to the
// even output: COS1 forward over Im F_k (Cos_FT.rs:64-67), SINFT forward over Re F_k (NR sinft, last loop),
// COS2F backwards (pos = N-1-k) over Im(F_k e^{i k pi/n}) after the rotation (Cos_FT2.rs:54-85).
NRB_DEV void scan_term(const AuxParams &A, const double2 *G, u64 pos, int mode, double &term, double &even, u64 &k)
{
    const u64 N = A.n / 2;
    if (mode == COS1) { k = pos; const double2 z = G[k]; term = k ? z.y : 0.0; even = z.x; }
    else if (mode == SINFT) { k = pos; const double2 z = G[k]; term = k ? z.x : 0.5 * z.x; even = k ? z.y : 0.0; }
    else {
        k = N - 1 - pos;
        double2 z = G[k];
        if (k) { const double2 t = two_level_tw(A.rtw_lo, A.rtw_hi, A.rtw_h, 2 * k); z = make_double2(z.x * t.x + z.y * t.y, z.y * t.x - z.x * t.y); }
        term = z.y; even = z.x;
    }
}

// Three-phase running sum over chunks of kScanChunk positions.  Phases 0 and 2 are block-cooperative (one CTA of
// kScanThreads threads per chunk, all global accesses coalesced, aux_scan_cta); phase 1 is one thread per line
// (aux_scan, op 1).  op 0 = chunk sums into b[l*nch + c]; op 1 = chunk sums -> chunk prefixes (starting value:
// COS1 the sum of the pre-pass in speq[l].x, SINFT 0, COS2F 0.5 F_{n/2}); op 2 = every chunk turns its terms into
// running sums and writes the even and odd outputs.

NRB_DEV void aux_scan(const AuxParams &A, u64 gtid, u64 gthreads)
{
    const u64 N = A.n / 2, nch = (N + kScanChunk - 1) / kScanChunk;
    const int mode = A.dir;
    double2 *P = const_cast<double2 *>(A.b);
    for (u64 l = gtid; l < A.count; l += gthreads) {
        double run = mode == COS1 ? A.speq[l].x : mode == COS2F ? 0.5 * A.a[(i64)(l * N)].y : 0.0;
        for (u64 c = 0; c < nch; ++c) { const double s = P[l * nch + c].x; P[l * nch + c].x = run; run += s; }
    }
}

// block = l * nch + c; sm holds kScanSmem double2
NRB_DEV void aux_scan_cta(const AuxParams &A, double2 *sm, unsigned block, int tid)
{
    const u64 N = A.n / 2, nch = (N + kScanChunk - 1) / kScanChunk;
    const int mode = A.dir;
    const u64 c = block % nch, l = block / nch;
    const double2 *G = A.a + (i64)(l * N);
    double2 *P = const_cast<double2 *>(A.b);
    const u64 p0 = c * kScanChunk;
    double2 *ts = sm + kScanChunk + kScanChunk / 8;          // per-thread sums
    // coalesced pass over the chunk: (term, even) of position p0 + i*T + tid
    double part = 0.0;
#pragma unroll
    for (int i = 0; i < kScanPer; ++i) {
        const int j = i * kScanThreads + tid;
        const u64 pos = p0 + (u64)j;
        double term = 0.0, even = 0.0;
        u64 k = 0;
        if (pos < N) scan_term(A, G, pos, mode, term, even, k);
        part += term;
        if (A.op == 2) sm[j + (j >> 3)] = make_double2(term, even);
    }
    if (A.op == 0) {
        ts[tid] = make_double2(part, 0.0);
        NRB_SYNC();
        for (int s = kScanThreads / 2; s > 0; s >>= 1) {
            if (tid < s) ts[tid].x += ts[tid + s].x;
            NRB_SYNC();
        }
        if (tid == 0) P[block] = make_double2(ts[0].x, 0.0);
        return;
    }
    NRB_SYNC();
    // thread tid owns positions [tid*PER, tid*PER + PER): local sum, block-wide exclusive scan of the sums
    double loc = 0.0;
#pragma unroll
    for (int i = 0; i < kScanPer; ++i) { const int j = tid * kScanPer + i; loc += sm[j + (j >> 3)].x; }
    ts[tid] = make_double2(loc, 0.0);
    NRB_SYNC();
    for (int s = 1; s < kScanThreads; s <<= 1) {             // Hillis-Steele inclusive scan
        const double add = tid >= s ? ts[tid - s].x : 0.0;
        NRB_SYNC();
        ts[tid].x += add;
        NRB_SYNC();
    }
    double run = P[block].x + (ts[tid].x - loc);
#pragma unroll
    for (int i = 0; i < kScanPer; ++i) {                    // (term, even) -> (odd output, even output)
        const int j = tid * kScanPer + i;
        const double2 te = sm[j + (j >> 3)];
        if (mode == COS2F) { sm[j + (j >> 3)] = make_double2(run, te.y); run += te.x; }   // exclusive, from the top
        else { run += te.x; sm[j + (j >> 3)] = make_double2(run, te.y); }                 // inclusive
    }
    NRB_SYNC();
    double *f = reinterpret_cast<double *>(A.out) + (i64)l * A.out_stride + 1;
    if (mode == COS1 && c == 0 && tid == 0) f[A.n] = G[0].y;                              // Cos_FT.rs:61  y[n+1] = y[2]
#pragma unroll
    for (int i = 0; i < kScanPer; ++i) {
        const int j = i * kScanThreads + tid;
        const u64 pos = p0 + (u64)j;
        if (pos < N) {
            const u64 k = mode == COS2F ? N - 1 - pos : pos;
            const double2 oe = sm[j + (j >> 3)];
            f[2 * k] = oe.y;
            f[2 * k + 1] = oe.x;
        }
    }
}

// Pointwise product of two device-resident complex arrays (rlft3 -> multiply -> rlft3^-1 chains
// without PCIe): out[i] = a[i] * b[i] * scale (op 0) or a[i] * conj(b[i]) * scale (op 1).  items: n.
NRB_DEV void aux_cmul(const AuxParams &A, u64 gtid, u64 gthreads)
{
    union { u64 u; double d; } sc;
    sc.u = A.m;
    for (u64 i = gtid; i < A.n; i += gthreads) {
        const double2 a = NRB_LDS(A.a + i), b = NRB_LDS(A.b + i);
        const double2 p = A.op ? make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y)
                               : make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
        A.out[i] = make_double2(p.x * sc.d, p.y * sc.d);
    }
}

// Cross-GPU barrier of the fused slab exchange, without a collective: after its stage-0 kernels (whose
// peer stores are complete when the kernel ends) every rank publishes the call's epoch into slot `rank`
// of each peer's flag array; stage 1 starts after the local array shows the epoch in all slots.
#if defined(NRB_EMU)
NRB_DEV void flag_store(unsigned long long *p, unsigned long long v) { *(volatile unsigned long long *)p = v; }
NRB_DEV unsigned long long flag_load(const unsigned long long *p) { return *(const volatile unsigned long long *)p; }
NRB_DEV void flag_pause() {}
#else
NRB_DEV void flag_store(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
NRB_DEV unsigned long long flag_load(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
NRB_DEV void flag_pause() { __nanosleep(200); }
#endif

NRB_DEV void aux_signal(const AuxParams &A, u64 gtid, u64)
{
    if (gtid < A.count) {
        flag_store(A.peer_flags[gtid] + A.n, A.m);
        if (A.op == 1) {      // signal and wait in one launch: A.out = the local flag array
#if !defined(NRB_EMU)
            const unsigned long long *local = reinterpret_cast<const unsigned long long *>(A.out);
            while (flag_load(local + gtid) < A.m) flag_pause();
#endif
        }
    }
}
NRB_DEV void aux_wait(const AuxParams &A, u64 gtid, u64)
{
    if (gtid < A.count) {
#if !defined(NRB_EMU)
        while (flag_load(A.peer_flags[0] + gtid) < A.m) flag_pause();
#endif
    }
}

NRB_DEV void aux_body(const AuxParams &A, u64 gtid, u64 gthreads)
{
    switch (A.kind) {
    case AUX_UNTANGLE: aux_untangle(A, gtid, gthreads); break;
    case AUX_SPECTRAL: aux_spectral(A, gtid, gthreads); break;
    case AUX_PAD_RESPONSE: aux_pad_response(A, gtid, gthreads); break;
    case AUX_FILL: aux_fill(A, gtid, gthreads); break;
    case AUX_SPECTRAL_Z: aux_spectral_z(A, gtid, gthreads); break;
    case AUX_SIGNAL: aux_signal(A, gtid, gthreads); break;
    case AUX_WAIT: aux_wait(A, gtid, gthreads); break;
    case AUX_REDUCE: aux_reduce(A, gtid, gthreads); break;
    case AUX_STATS_FINAL: aux_stats_final(A, gtid, gthreads); break;
    case AUX_NORMALIZE: aux_normalize(A, gtid, gthreads); break;
    case AUX_POWER: aux_power(A, gtid, gthreads); break;
    case AUX_PACK2: aux_pack2(A, gtid, gthreads); break;
    case AUX_TWOFFT_SPLIT: aux_twofft_split(A, gtid, gthreads); break;
    case AUX_SCALE: aux_scale(A, gtid, gthreads); break;
    case AUX_COSFT: aux_cosft(A, gtid, gthreads); break;
    case AUX_SCAN: aux_scan(A, gtid, gthreads); break;
    case AUX_CMUL: aux_cmul(A, gtid, gthreads); break;
    case AUX_SPECTRAL_ZT: aux_spectral_zt(A, gtid, gthreads); break;
    default: aux_correl_direct(A, gtid, gthreads); break;
    }
}

} // namespace nrb
