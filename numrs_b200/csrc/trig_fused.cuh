// trig_fused.cuh -- cosft1 / cosft2 / sinft and twofft for lines that fit on chip: ONE kernel, ONE HBM pass.
//
// NR's cosine / sine transforms are O(n) pre- and post-processing around realft (Cos_FT.rs:7-74, Cos_FT2.rs:7-162,
// README.md:72; oracle ledger D11), twofft is a packing step, four1 and a separation step (FFT_2.rs:3-90, ledger D9).
// As separate launches that is 5-7 sweeps over HBM per call (aux_cosft, aux_reduce, the ROW-REAL pass, three scan
// phases; aux_pack2, the pass, aux_twofft_split): 0.10-0.15 (0.33 for twofft) of the measured copy bandwidth
// (profiles/r02_kernel_table_next.txt).  Here a CTA owns L whole lines (2048 complex points per tile, or one line of up to 8192) and
// does every step while the lines are in shared memory:
//
//   cosft1 / cosft2 forward / sinft:  load the line with aligned 16-byte accesses (the reference's arrays are 1-based,
//     so a line starts on an odd or even double: scalar head / tail) -> pre-processing into the realft work array
//     (cosft1 also accumulates its `sum`, reduced over the line's threads by warp shuffles) -> c2c stages (the Stockham
//     stages of fft_pass2.cuh / conv_mid.cuh on registers, exchanges through the padded buffer) -> NR realft untangling
//     in place -> running sum of the odd outputs (per-thread runs of 8 positions, segmented warp-shuffle scan over the
//     line's threads) -> aligned 16-byte stores.
//   cosft2 inverse: pre-rotation -> inverse untangling -> c2c -> post-processing -> store.
//   twofft: two real lines loaded as one complex line -> c2c (isign = +1) -> separation of the two spectra straight
//     into fft1 / fft2.
//
// Same arithmetic per element as aux_cosft / scan_term / aux_twofft_split and the ROW-REAL pass (the tests compare both
// paths); only the order of the additions inside cosft1's `sum` and the running sums differs.  Algorithmic bytes per
// line: 2 * 8 * (n + 1) (cosft1), 2 * 8 * n (cosft2, sinft), 16 n + 32 (n + 1) (twofft).  Bound: HBM.
#pragma once
#include "conv_mid.cuh"

namespace nrb {

// complex points per line: 2^kTrigMinLog2 .. 2^kTrigMaxLog2 (nrb_common.h)
// complex points per CTA tile (lines longer than that take one CTA each): 2^11 = 256 threads, 36 KiB of shared memory, four
// CTAs per SM.  NRB_TRIG_TILE_LOG2 / NRB_TRIG_MINB (resident CTAs per SM the kernel is compiled for) are A/B switches.
#ifndef NRB_TRIG_TILE_LOG2
#define NRB_TRIG_TILE_LOG2 11
#endif
constexpr int kTrigTileLog2 = NRB_TRIG_TILE_LOG2;
NRB_HD constexpr int trig_min_ctas(int log2n)
{
#ifdef NRB_TRIG_MINB
    return log2n <= kTrigTileLog2 ? NRB_TRIG_MINB : (log2n >= 13 ? 1 : log2n == 12 ? 2 : 4);
#else
    return (log2n > kTrigTileLog2 ? log2n : kTrigTileLog2) >= 13 ? 1 : (log2n > kTrigTileLog2 ? log2n : kTrigTileLog2) == 12 ? 2
         : (log2n > kTrigTileLog2 ? log2n : kTrigTileLog2) == 11 ? 4 : 8;
#endif
}

template <int LOG2N_> struct GeoT {
    static constexpr int LOG2N = LOG2N_, LAYOUT = LAYOUT_ROW, VARIANT = VAR_PLAIN;
    static constexpr int N = 1 << LOG2N;
    static constexpr int TL = LOG2N > kTrigTileLog2 ? LOG2N : kTrigTileLog2;
    static constexpr int TILE = 1 << TL;
    static constexpr int L = TILE / N;
    static constexpr int PPT = 8;
    static constexpr int NT = TILE / PPT;
    static constexpr int TPL = N / PPT;                       // threads per line (contiguous thread ids)
    static constexpr int NST = radix_plan(LOG2N).nst;
    static constexpr int LP = N + (N >> 3);
    static constexpr int WS = 32;                             // scratch behind the tile: warp totals of the segmented scan (doubles)
    static constexpr size_t SMEM_BYTES = (size_t)L * LP * sizeof(double2) + WS * sizeof(double);
    static_assert(N >= PPT, "trig_fused: lines shorter than the points per thread are not built");
    static_assert(NT % 32 == 0 && NT / 32 <= WS, "trig_fused: whole warps, at most 32 of them");
    NRB_DEVM static int phys(int l, int n) { return l * LP + n + (n >> 3); }
};

// Inclusive prefix sum of v over the TPL consecutive threads of a line (thread ids [ln * TPL, (ln + 1) * TPL)), and the
// line's total.  Up to 32 threads per line: shuffles inside the warp; longer lines: warp totals through `ws`.
// Every thread of the CTA must call this (it contains CTA-wide barriers when TPL > 32).
template <class G>
NRB_DEV double seg_scan(double v, int tid, double *ws, double &total)
{
    constexpr int TPL = G::TPL, W = TPL < 32 ? TPL : 32;
    const int lane = tid & 31, ls = tid & (W - 1);
#pragma unroll
    for (int d = 1; d < W; d <<= 1) {
        const double t = NRB_SHFL(v, (lane - d) & 31);
        if (ls >= d) v += t;
    }
    total = W > 1 ? NRB_SHFL(v, lane | (W - 1)) : v;
    if (TPL > 32) {
        constexpr int WPL = TPL / 32 > 0 ? TPL / 32 : 1;       // warps per line
        const int w = tid >> 5, w0 = (w / WPL) * WPL;
        if (lane == 31) ws[w] = v;
        NRB_SYNC();
        double before = 0.0, all = 0.0;
#pragma unroll 4
        for (int x = 0; x < WPL; ++x) {
            const double s = ws[w0 + x];
            if (w0 + x < w) before += s;
            all += s;
        }
        NRB_SYNC();                                            // ws may be written again by the next call
        v += before;
        total = all;
    }
    return v;
}

// ------------------------------------------------------------------ cosft1 / cosft2 / sinft

// real element r of line l in the padded tile (the realft work array g[0 .. n))
template <class G> NRB_DEVM int trig_ri(int l, int r) { return 2 * G::phys(l, r >> 1) + (r & 1); }

template <int LOG2N>
NRB_DEV void trig_cta(const TrigParams &T, double2 *E, unsigned tile, int tid)
{
    typedef GeoT<LOG2N> G;
    typedef Stage2<G, 0> S0;
    constexpr int N = G::N, n = 2 * N, TPL = G::TPL;
    double *Ed = reinterpret_cast<double *>(E);
    double *ws = Ed + 2 * G::L * G::LP;
    const int mode = T.mode;
    const int ln = tid / TPL, lt = tid & (TPL - 1);           // the line this thread works on in the per-line phases
    const u64 q_own = (u64)tile * G::L + (u64)ln;
    const int nd = mode == COS1 ? n + 1 : n;                   // doubles of a line that carry data

    // ---- load: aligned 16-byte chunks of every line into the tile (natural order, no padding: line l at 2 * l * LP doubles,
    // shifted by the line's alignment so that chunk c of the global line is chunk c of the staging area).  All of a thread's
    // loads are issued before the first of them is used (one DRAM latency per tile, not one per chunk).
    constexpr int CH = N + 1;                                  // chunks that cover a line on either alignment
    constexpr int ITER = (G::L * CH + G::NT - 1) / G::NT;
    {
        double2 buf[ITER];
#pragma unroll
        for (int i = 0; i < ITER; ++i) {
            const int it = tid + i * G::NT;
            const int l = it / CH, c = it - l * CH;
            const u64 q = (u64)tile * G::L + (u64)l;
            buf[i] = make_double2(0.0, 0.0);
            if (it < G::L * CH && q < T.count) {
                const double *gb = T.io + (i64)q * T.ld + 1;
                const int par = (int)((reinterpret_cast<size_t>(gb) >> 3) & 1);
                const int d0 = 2 * c - par;
                if (d0 >= 0 && d0 + 1 < nd) buf[i] = NRB_LDS(reinterpret_cast<const double2 *>(gb + d0));
                else {
                    if (d0 >= 0 && d0 < nd) buf[i].x = NRB_LDS(gb + d0);
                    if (d0 + 1 >= 0 && d0 + 1 < nd) buf[i].y = NRB_LDS(gb + d0 + 1);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < ITER; ++i) {
            const int it = tid + i * G::NT;
            const int l = it / CH, c = it - l * CH;
            if (it < G::L * CH) E[l * G::LP + c] = buf[i];
        }
    }
    NRB_SYNC();

    // ---- pre-processing: all of a thread's inputs into registers, barrier, then the work array in the padded layout ----
    const int par_own = (int)((reinterpret_cast<size_t>(T.io + (i64)q_own * T.ld + 1) >> 3) & 1);
    const double *S = Ed + 2 * ln * G::LP + par_own;
    // exp(-2 pi i m / M) of the thread's items: m advances by TPL (M = 2n: cosft1, sinft, the inverse's pre-rotation) or by
    // 2 TPL (M = 4n, m odd: cosft2), i.e. by the fixed angle pi / 16 -- one table look-up and the rotations exp(-i pi i / 16)
    // instead of a look-up per item (the kernel is L1-bound: ncu, profiles/r02_ncu_full_trig.md)
    const u64 m0 = mode == COS2F ? (u64)(2 * lt + 1) : mode == COS2I_PRE ? (u64)(2 * lt) : (u64)lt;
    const double2 t0 = two_level_tw(T.ctw_lo, T.ctw_hi, T.ctw_h, m0);
    double fa[G::PPT], fb[G::PPT], fc = 0.0;
#pragma unroll
    for (int i = 0; i < G::PPT; ++i) {
        const int j = lt + i * TPL;                            // 0 <= j < N
        if (mode == COS1 || mode == SINFT) { fa[i] = S[j]; fb[i] = mode == SINFT && j == 0 ? 0.0 : S[n - j]; if (j == 0) fc = S[N]; }
        else if (mode == COS2F) { fa[i] = S[j]; fb[i] = S[n - 1 - j]; }
        else {                                                 // inverse cosft2 (Cos_FT2.rs:97-131)
            if (j == 0) { fa[i] = S[0]; fb[i] = S[n - 1]; }
            else { fa[i] = S[2 * j]; fb[i] = S[2 * j - 1] - S[2 * j + 1]; }
        }
    }
    NRB_SYNC();
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < G::PPT; ++i) {
        const int j = lt + i * TPL;
        if (mode == COS1) {                                    // Cos_FT.rs:17,36-55
            if (j == 0) {
                Ed[trig_ri<G>(ln, 0)] = 0.5 * (fa[i] + fb[i]);
                acc += 0.5 * (fa[i] - fb[i]);
                Ed[trig_ri<G>(ln, N)] = fc;
            } else {
                const double2 t = i ? cmul(t0, unit_rot<16>(i)) : t0;                      // (cos, -sin)(j pi / n)
                const double y1 = 0.5 * (fa[i] + fb[i]), y2 = fa[i] - fb[i];
                Ed[trig_ri<G>(ln, j)] = y1 + t.y * y2;
                Ed[trig_ri<G>(ln, n - j)] = y1 - t.y * y2;
                acc += t.x * y2;
            }
        } else if (mode == COS2F) {                            // Cos_FT2.rs:26-36
            const double2 t = i ? cmul(t0, unit_rot<16>(i)) : t0;                          // exp(-i (2j+1) pi / (2n))
            const double y1 = 0.5 * (fa[i] + fb[i]), y2 = -t.y * (fa[i] - fb[i]);
            Ed[trig_ri<G>(ln, j)] = y1 + y2;
            Ed[trig_ri<G>(ln, n - 1 - j)] = y1 - y2;
        } else if (mode == SINFT) {                            // NR sinft, first loop
            if (j == 0) {
                Ed[trig_ri<G>(ln, 0)] = 0.0;
                const double2 t = two_level_tw(T.ctw_lo, T.ctw_hi, T.ctw_h, (u64)N);
                Ed[trig_ri<G>(ln, N)] = -t.y * (fc + fc);      // j = N: y2 = 0
            } else {
                const double2 t = i ? cmul(t0, unit_rot<16>(i)) : t0;
                const double y1 = -t.y * (fa[i] + fb[i]), y2 = 0.5 * (fa[i] - fb[i]);
                Ed[trig_ri<G>(ln, j)] = y1 + y2;
                Ed[trig_ri<G>(ln, n - j)] = y1 - y2;
            }
        } else {
            if (j == 0) E[G::phys(ln, 0)] = make_double2(fa[i], 2.0 * fb[i]);
            else {
                const double2 t = i ? cmul(t0, unit_rot<16>(i)) : t0;                      // (cos, -sin)(j pi / n)
                E[G::phys(ln, j)] = make_double2(fa[i] * t.x - fb[i] * t.y, fb[i] * t.x + fa[i] * t.y);
            }
        }
    }
    double sum = 0.0;                                          // cosft1: the line's `sum`, known to every thread of the line
    if (mode == COS1) seg_scan<G>(acc, tid, ws, sum);
    NRB_SYNC();

    const bool inverse = mode == COS2I_PRE;
    if (inverse) {
        // ---- inverse realft untangling in place (Real_FT.rs:133-176): a pair (k, N - k) has one owner ----
#pragma unroll
        for (int i = 0; i < G::PPT / 2; ++i) {
            const int k = lt + i * TPL;                        // 0 <= k < N/2
            if (k == 0) {
                const double2 a = E[G::phys(ln, 0)];
                E[G::phys(ln, 0)] = make_double2(0.5 * (a.x + a.y), 0.5 * (a.x - a.y));
            } else {
                const double2 a = E[G::phys(ln, k)], b = E[G::phys(ln, N - k)];
                double2 oa, ob;
                untangle_pair<-1>(a, b, NRB_LDG(T.rtw + k), oa, ob);
                E[G::phys(ln, k)] = oa;
                E[G::phys(ln, N - k)] = ob;
            }
        }
        NRB_SYNC();
    }

    // ---- c2c transform of the N complex points of every line (forward: reference isign = +1, re / im swapped on the way in) ----
    {
        double2 v[G::PPT];
#pragma unroll
        for (int i = 0; i < S0::BPT; ++i) {
            int l2, jj;
            S0::coords(tid, i, l2, jj);
#pragma unroll
            for (int r = 0; r < S0::R; ++r) {
                const double2 x = E[G::phys(l2, jj + r * S0::NB)];
                v[i * S0::R + r] = inverse ? x : cswap(x);
            }
        }
        NRB_SYNC();                                            // everyone has read before the first exchange overwrites
        ChainM<G, 0>::run(T, E, tid, v);
        mid_scatter<G, G::NST - 1>(E, tid, v);                 // natural order; forward: the true bin is cswap(E[..])
    }
    NRB_SYNC();

    if (inverse) {
        // ---- cosft2 inverse post-processing (Cos_FT2.rs:142-162) in place: a pair (i, n - 1 - i) has one owner ----
        const double2 tp0 = two_level_tw(T.ctw_lo, T.ctw_hi, T.ctw_h, (u64)(2 * lt + 1));
#pragma unroll
        for (int i = 0; i < G::PPT; ++i) {
            const int j = lt + i * TPL;
            const double2 t = i ? cmul(tp0, unit_rot<16>(i)) : tp0;                          // sin((2j+1) pi/(2n)) = -t.y
            const double gi = Ed[trig_ri<G>(ln, j)], gm = Ed[trig_ri<G>(ln, n - 1 - j)];
            const double y1 = gi + gm, y2 = (0.5 / -t.y) * (gi - gm);
            Ed[trig_ri<G>(ln, j)] = 0.5 * (y1 + y2);
            Ed[trig_ri<G>(ln, n - 1 - j)] = 0.5 * (y1 - y2);
        }
    } else {
        // ---- realft untangling in place (Real_FT.rs:43-80) ----
#pragma unroll
        for (int i = 0; i < G::PPT / 2; ++i) {
            const int k = lt + i * TPL;
            if (k == 0) {
                const double2 z0 = cswap(E[G::phys(ln, 0)]);
                E[G::phys(ln, 0)] = make_double2(z0.x + z0.y, z0.x - z0.y);
                E[G::phys(ln, N / 2)] = cswap(E[G::phys(ln, N / 2)]);
            } else {
                const double2 a = cswap(E[G::phys(ln, k)]), b = cswap(E[G::phys(ln, N - k)]);
                double2 oa, ob;
                untangle_pair<1>(a, b, NRB_LDG(T.rtw + k), oa, ob);
                E[G::phys(ln, k)] = oa;
                E[G::phys(ln, N - k)] = ob;
            }
        }
        NRB_SYNC();
        // ---- running sums: thread lt owns positions lt * 8 .. lt * 8 + 7 of its line (scan_term's rules) ----
        double term[G::PPT], even[G::PPT];
        double loc = 0.0;
        const double g0y = E[G::phys(ln, 0)].y;
        // cosft2: exp(-i k pi / n) for k = N - 1 - pos walks down by one per position: one look-up and the step exp(+i pi / n)
        double2 tk = make_double2(1.0, 0.0), tstep = tk;
        if (mode == COS2F) {
            tk = two_level_tw(T.ctw_lo, T.ctw_hi, T.ctw_h, (u64)(2 * (N - 1 - lt * G::PPT)));
            tstep = cconj(two_level_tw(T.ctw_lo, T.ctw_hi, T.ctw_h, 2));
        }
#pragma unroll
        for (int i = 0; i < G::PPT; ++i) {
            const int pos = lt * G::PPT + i;
            const int k = mode == COS2F ? N - 1 - pos : pos;
            double2 z = E[G::phys(ln, k)];
            if (mode == COS1) { term[i] = k ? z.y : 0.0; even[i] = z.x; }                    // Cos_FT.rs:64-67
            else if (mode == SINFT) { term[i] = k ? z.x : 0.5 * z.x; even[i] = k ? z.y : 0.0; }
            else {                                                                            // Cos_FT2.rs:54-85
                if (k) z = make_double2(z.x * tk.x + z.y * tk.y, z.y * tk.x - z.x * tk.y);
                tk = cmul(tk, tstep);
                term[i] = z.y; even[i] = z.x;
            }
            loc += term[i];
        }
        double total;
        const double incl = seg_scan<G>(loc, tid, ws, total);
        NRB_SYNC();                                            // every thread holds its eight bins (and g0y) before they are overwritten
        double run = (mode == COS1 ? sum : mode == COS2F ? 0.5 * g0y : 0.0) + (incl - loc);
#pragma unroll
        for (int i = 0; i < G::PPT; ++i) {
            const int pos = lt * G::PPT + i;
            const int k = mode == COS2F ? N - 1 - pos : pos;
            double odd;
            if (mode == COS2F) { odd = run; run += term[i]; }   // exclusive, from the top
            else { run += term[i]; odd = run; }                 // inclusive
            E[G::phys(ln, k)] = make_double2(even[i], odd);
        }
        if (mode == COS1 && lt == 0 && q_own < T.count) T.io[(i64)q_own * T.ld + 1 + n] = g0y;   // Cos_FT.rs:61  y[n+1] = y[2]
    }
    NRB_SYNC();

    // ---- store: f[2k] = E[k].x, f[2k+1] = E[k].y in aligned 16-byte chunks.  On an odd line chunk c is (E[c-1].y, E[c].x):
    // the left half comes from the neighbouring lane (chunk c - 1 of the same line) by a shuffle, so every chunk costs one
    // 16-byte shared-memory read; scalar head / tail.
#pragma unroll
    for (int i = 0; i < ITER; ++i) {
        const int it = tid + i * G::NT;
        const int l = it / CH, c = it - l * CH;
        const u64 q = (u64)tile * G::L + (u64)l;
        const bool ok = it < G::L * CH && q < T.count;
        double2 own = make_double2(0.0, 0.0);
        if (it < G::L * CH && c < N) own = E[G::phys(l, c)];
        double left = NRB_SHFL(own.y, (tid - 1) & 31);
        if (!ok) continue;
        double *gb = T.io + (i64)q * T.ld + 1;
        const int par = (int)((reinterpret_cast<size_t>(gb) >> 3) & 1);
        if (par == 0) {
            if (c < N) NRB_STS(reinterpret_cast<double2 *>(gb) + c, own);
        } else if (c == 0) {
            gb[0] = own.x;
        } else {
            if ((tid & 31) == 0) left = E[G::phys(l, c - 1)].y;   // the neighbour is in another warp
            if (c == N) gb[n - 1] = left;
            else NRB_STS(reinterpret_cast<double2 *>(gb + 2 * c - 1), make_double2(left, own.x));
        }
    }
}

// ------------------------------------------------------------------ twofft

template <int LOG2N>
NRB_DEV void twofft_cta(const TwoFFTParams &T, double2 *E, unsigned tile, int tid)
{
    typedef GeoT<LOG2N> G;
    typedef Stage2<G, 0> S0;
    constexpr int N = G::N;                                    // complex points per line = the real length n
    // ---- pack (FFT_2.rs:33-37): point j = (d1[j], d2[j]); 16-byte loads of two points' worth of each signal ----
    // (a caller of the plan API may hand over signals that are only 8-byte aligned: scalar loads then); all of a thread's
    // loads are issued before the first of them is used
    const bool vec = ((reinterpret_cast<size_t>(T.d1) | reinterpret_cast<size_t>(T.d2)) & 15) == 0;
    constexpr int PITER = (G::L * (N / 2)) / G::NT;            // = 4: TILE / 2 chunks of two points over TILE / 8 threads
    static_assert(PITER * G::NT == G::L * (N / 2), "twofft: the chunks of a tile divide evenly over the threads");
    {
        double2 a[PITER], b[PITER];
#pragma unroll
        for (int i = 0; i < PITER; ++i) {
            const int it = tid + i * G::NT;
            const int l = it / (N / 2), c = it - l * (N / 2);
            const u64 q = (u64)tile * G::L + (u64)l;
            a[i] = make_double2(0.0, 0.0);
            b[i] = a[i];
            if (q < T.count) {
                const double *p1 = T.d1 + (i64)q * N + 2 * c, *p2 = T.d2 + (i64)q * N + 2 * c;
                if (vec) { a[i] = NRB_LDS(reinterpret_cast<const double2 *>(p1)); b[i] = NRB_LDS(reinterpret_cast<const double2 *>(p2)); }
                else { a[i] = make_double2(NRB_LDS(p1), NRB_LDS(p1 + 1)); b[i] = make_double2(NRB_LDS(p2), NRB_LDS(p2 + 1)); }
            }
        }
#pragma unroll
        for (int i = 0; i < PITER; ++i) {
            const int it = tid + i * G::NT;
            const int l = it / (N / 2), c = it - l * (N / 2);
            E[G::phys(l, 2 * c)] = make_double2(a[i].x, b[i].x);
            E[G::phys(l, 2 * c + 1)] = make_double2(a[i].y, b[i].y);
        }
    }
    NRB_SYNC();
    // ---- four1(fft1, n, 1) (FFT_2.rs:13) ----
    {
        double2 v[G::PPT];
#pragma unroll
        for (int i = 0; i < S0::BPT; ++i) {
            int l2, jj;
            S0::coords(tid, i, l2, jj);
#pragma unroll
            for (int r = 0; r < S0::R; ++r) v[i * S0::R + r] = cswap(E[G::phys(l2, jj + r * S0::NB)]);
        }
        NRB_SYNC();
        ChainM<G, 0>::run(T, E, tid, v);
        mid_scatter<G, G::NST - 1>(E, tid, v);                 // natural order; the true bin is cswap(E[..])
    }
    NRB_SYNC();
    // ---- separation (FFT_2.rs:53-90 with the 0-based mirror n - k, ledger D9): item k owns bins k and n - k ----
    for (int it = tid; it < G::L * (N / 2 + 1); it += G::NT) {
        const int l = it / (N / 2 + 1), k = it - l * (N / 2 + 1);
        const u64 q = (u64)tile * G::L + (u64)l;
        if (q >= T.count) continue;
        double2 *f1 = T.f1 + (i64)q * (N + 1), *f2 = T.f2 + (i64)q * (N + 1);
        if (k == 0) {
            const double2 z = cswap(E[G::phys(l, 0)]);
            f1[0] = make_double2(z.x, 0.0);
            f2[0] = make_double2(z.y, 0.0);
            f1[N] = make_double2(0.0, 0.0);
            f2[N] = make_double2(0.0, 0.0);
        } else {
            const int m = N - k;
            const double2 a = cswap(E[G::phys(l, k)]), b = cswap(E[G::phys(l, m)]);
            const double rep = 0.5 * (a.x + b.x), rem = 0.5 * (a.x - b.x);
            const double aip = 0.5 * (a.y + b.y), aim = 0.5 * (a.y - b.y);
            NRB_STS(f1 + k, make_double2(rep, aim));
            NRB_STS(f2 + k, make_double2(aip, -rem));
            if (m != k) { NRB_STS(f1 + m, make_double2(rep, -aim)); NRB_STS(f2 + m, make_double2(aip, rem)); }
        }
    }
}

} // namespace nrb
