// fft_pass.cuh -- the batched shared-memory Stockham FFT pass (f64, complex interleaved).
//
// One CTA transforms a tile of L = TILE/N lines of N = 2^LOG2N points (TILE = 4096 points,
// 8192 for N = 8192) with TILE/16 threads; every thread owns 16 points per stage, i.e.
// 16/R radix-R butterflies.  Stockham autosort: stage s reads x[j + r*N/R], multiplies by
// exp(-2 pi i (j mod Ns) r / (Ns R)), does an R-point DFT and writes
// y[(j - j mod Ns)*R + (j mod Ns) + r*Ns]; the first stage reads global memory directly
// and the last one writes it directly, so an NST-stage transform makes NST-1 round trips
// through shared memory.
//
// Two thread/shared-memory layouts:
//   ROW  lines are contiguous in global memory.  thread -> (j fastest, then line);
//        smem [line][n + n/8] (one pad element per 8 keeps every Stockham write pattern
//        of radix 2/4/8/16 stages conflict-free for 16-byte accesses).
//   COL  consecutive lines are adjacent in global memory (element stride >= #lines).
//        thread -> (line fastest, then j); smem [n][line]: the 8 threads of a 16-byte
//        shared-memory phase always touch 8 consecutive elements -> conflict-free with no
//        padding, and every global access is a run of L*16 bytes.
//
// Replaces (reference, /root/reference/src): four1 bit reversal + Danielson-Lanczos stages
// FFT_1.rs:8-43; the per-dimension loops of NR fourn (call shape Real_FT3.rs:35); the realft
// untangling loops Real_FT.rs:49-80,145-176 and DC/Nyquist handling :43-45,:133-135; the
// z-direction part of the rlft3 loop nest Real_FT3.rs:60-127.
//
// Direction: the core always computes exp(-2 pi i jk/N).  dir = +1 (the reference's
// isign = +1, exp(+...)) is obtained by swapping re/im on the way in and out (free: it is a
// compile-time register renaming).
#pragma once
#include "nrb_common.h"

// REAL inverse pre-pass thread mapping: 1 = the line's own thread group prepares it (warp-level sync
// after it), 0 = flat index mapping + one CTA barrier.  Measured on B200: 0 is 20 % faster (5.2 vs 4.35 TB/s).
#ifndef NRB_TW_POWERS
#define NRB_TW_POWERS 1
#endif
#ifndef NRB_TW_POWERS_COL
#define NRB_TW_POWERS_COL 1
#endif
// radix-8 twiddle powers: 1 = all built from one table load (6 complex products), 3 = w, w^2 and w^4 loaded, the other
// four built from them (4 complex products, 8 fewer FP64 instructions per butterfly, 2 more L1 hits).  Not measured yet:
// build a variant with -DNRB_TW_LOADS=3 and compare (the passes are issue / latency bound, FP64 is ~40 % of their mix).
#ifndef NRB_TW_LOADS
#define NRB_TW_LOADS 1
#endif
#ifndef NRB_PRE_OWN
#define NRB_PRE_OWN 0
#endif
#ifndef NRB_REAL_INV_DIRECT
#define NRB_REAL_INV_DIRECT 0
#endif
// inverse REAL pre-pass: untangle twiddles of a thread's items from one table load + fixed rotations (1) or one load each (0)
#ifndef NRB_REAL_INV_TWROT
#define NRB_REAL_INV_TWROT 1
#endif

namespace nrb {

// ------------------------------------------------------------------ complex helpers
NRB_DEV double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
NRB_DEV double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
NRB_DEV double2 cmul(double2 a, double2 b)
{
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
NRB_DEV double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
NRB_DEV double2 cswap(double2 a) { return make_double2(a.y, a.x); }
NRB_DEV double2 mul_mi(double2 a) { return make_double2(a.y, -a.x); }  // a * (-i)
template <int DIR> NRB_DEV double2 io_swap(double2 a) { return DIR > 0 ? cswap(a) : a; }

// ------------------------------------------------------------------ butterflies (forward, e^{-})
template <int R> struct Bfly;

template <> struct Bfly<2> {
    NRB_DEVM static void run(double2 *v)
    {
        double2 t = v[0];
        v[0] = cadd(t, v[1]);
        v[1] = csub(t, v[1]);
    }
};

NRB_DEV void bfly4(double2 &v0, double2 &v1, double2 &v2, double2 &v3)
{
    double2 a = cadd(v0, v2), b = csub(v0, v2), c = cadd(v1, v3), d = mul_mi(csub(v1, v3));
    v0 = cadd(a, c);
    v1 = cadd(b, d);
    v2 = csub(a, c);
    v3 = csub(b, d);
}

template <> struct Bfly<4> {
    NRB_DEVM static void run(double2 *v) { bfly4(v[0], v[1], v[2], v[3]); }
};

template <> struct Bfly<8> {
    NRB_DEVM static void run(double2 *v)
    {
        const double h = 0.70710678118654752440;
        bfly4(v[0], v[2], v[4], v[6]);   // even: E[k] in v[0],v[2],v[4],v[6]
        bfly4(v[1], v[3], v[5], v[7]);   // odd : O[k] in v[1],v[3],v[5],v[7]
        double2 o1 = make_double2((v[3].x + v[3].y) * h, (v[3].y - v[3].x) * h);   // * W8^1
        double2 o2 = mul_mi(v[5]);                                                 // * W8^2
        double2 o3 = make_double2((v[7].y - v[7].x) * h, -(v[7].x + v[7].y) * h);  // * W8^3
        double2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1];
        v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
        v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
        v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
        v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
    }
};

template <> struct Bfly<16> {
    NRB_DEVM static void run(double2 *v)
    {
        // n = 4a + b, k = k1 + 4 k2
        const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;  // cos/sin(pi/8)
        const double h = 0.70710678118654752440;
        // step 1: for each b, DFT4 over a: inputs v[b], v[4+b], v[8+b], v[12+b] -> u_b[k1] in place
        bfly4(v[0], v[4], v[8], v[12]);
        bfly4(v[1], v[5], v[9], v[13]);
        bfly4(v[2], v[6], v[10], v[14]);
        bfly4(v[3], v[7], v[11], v[15]);
        // step 2: u_b[k1] *= W16^(b*k1); u_b[k1] sits in v[4*k1 + b]
        v[5] = cmul(v[5], make_double2(c1, -s1));                                  // W^1
        v[6] = make_double2((v[6].x + v[6].y) * h, (v[6].y - v[6].x) * h);         // W^2
        v[7] = cmul(v[7], make_double2(s1, -c1));                                  // W^3
        v[9] = make_double2((v[9].x + v[9].y) * h, (v[9].y - v[9].x) * h);         // W^2
        v[10] = mul_mi(v[10]);                                                     // W^4
        v[11] = make_double2((v[11].y - v[11].x) * h, -(v[11].x + v[11].y) * h);   // W^6
        v[13] = cmul(v[13], make_double2(s1, -c1));                                // W^3
        v[14] = make_double2((v[14].y - v[14].x) * h, -(v[14].x + v[14].y) * h);   // W^6
        v[15] = cmul(v[15], make_double2(-c1, s1));                                // W^9
        // step 3: for each k1, DFT4 over b: v[4k1+0..3] -> X[k1 + 4 k2] at v[4k1 + k2]
        bfly4(v[0], v[1], v[2], v[3]);
        bfly4(v[4], v[5], v[6], v[7]);
        bfly4(v[8], v[9], v[10], v[11]);
        bfly4(v[12], v[13], v[14], v[15]);
        // now v[4k1 + k2] = X[k1 + 4k2]; transpose to natural order v[k1 + 4k2]
        double2 t;
        t = v[1];  v[1] = v[4];   v[4] = t;
        t = v[2];  v[2] = v[8];   v[8] = t;
        t = v[3];  v[3] = v[12];  v[12] = t;
        t = v[6];  v[6] = v[9];   v[9] = t;
        t = v[7];  v[7] = v[13];  v[13] = t;
        t = v[11]; v[11] = v[14]; v[14] = t;
    }
};

// ------------------------------------------------------------------ compile-time geometry
template <int LOG2N, int LAYOUT, int VARIANT> struct Geo {
    static constexpr int N = 1 << LOG2N;
    static constexpr int TL = tile_log2(LOG2N, LAYOUT);
    static constexpr int TILE = 1 << TL;
    static constexpr int L = TILE / N;                 // lines per tile
    static constexpr int NT = cta_threads(LOG2N, LAYOUT);      // threads per CTA
    static constexpr int PPT = points_per_thread(LAYOUT, LOG2N);
    static constexpr int NST = radix_plan(LOG2N).nst;
    static constexpr int LP = row_line_pitch(LOG2N);   // ROW: line pitch
    static constexpr int CP = col_pitch(LOG2N, VARIANT); // COL: pitch of one n
    NRB_DEVM static int phys(int l, int n)
    {
        if (LAYOUT == LAYOUT_ROW) return l * LP + n + (n >> 3);
        if (VARIANT == VAR_XPOSE && L > 1 && L < 8) {
            // 8-thread phases of the stages touch an aligned run of 8 elements (a >> 3 constant), the
            // line-contiguous read-back touches a = n*L + l for 8 consecutive n: XOR the missing bits
            // of n into the low 3 bits so both patterns hit 8 distinct 16-byte bank groups
            const int a = n * L + l;
            return a ^ ((a >> 3) & (L == 4 ? 3 : 1));
        }
        return n * CP + l;
    }
};

NRB_DEV i64 line_base(u64 q, i64 s0, i64 s1, i64 s2, int logA, int logB)
{
    const u64 q2 = q & ((1ull << logB) - 1ull);
    const u64 q1 = (q >> logB) & ((1ull << logA) - 1ull);
    const u64 q0 = q >> (logA + logB);
    return (i64)q0 * s0 + (i64)q1 * s1 + (i64)q2 * s2;
}

NRB_DEV i64 elem_off(int n, i64 es, int eshift, i64 es_hi)
{
    return (i64)(n & ((1 << eshift) - 1)) * es + (i64)(n >> eshift) * es_hi;
}

// four-step twiddle exp(-2 pi i m / M), m = q1 * k, from the two-level table
NRB_DEV double2 fourstep_tw(const PassParams &P, u64 q, unsigned k)
{
    const unsigned q1 = (unsigned)((q >> P.logB) & ((1ull << P.logA) - 1ull));
    const unsigned m = q1 * k;
    const double2 lo = NRB_LDG(P.tw_lo + (m & ((1u << P.tw_h) - 1u)));
    const double2 hi = NRB_LDG(P.tw_hi + (m >> P.tw_h));
    return cmul(lo, hi);
}

// exp(-2 pi i m / M) for an explicit exponent m < M
NRB_DEV double2 fourstep_tw_m(const PassParams &P, unsigned m)
{
    const double2 lo = NRB_LDG(P.tw_lo + (m & ((1u << P.tw_h) - 1u)));
    const double2 hi = NRB_LDG(P.tw_hi + (m >> P.tw_h));
    return cmul(lo, hi);
}
NRB_DEV unsigned line_q1(const PassParams &P, u64 q)
{
    return (unsigned)((q >> P.logB) & ((1ull << P.logA) - 1ull));
}

// cos(m pi / 16), 0 <= m <= 16
NRB_HD constexpr double cos_pi16(int m)
{
    constexpr double c[17] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708, 0.70710678118654752440,
                              0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785, 0.0,
                              -0.19509032201612826785, -0.38268343236508977173, -0.55557023301960222474, -0.70710678118654752440,
                              -0.83146961230254523708, -0.92387953251128675613, -0.98078528040323044913, -1.0};
    return c[m];
}
// exp(-i pi r / R) for compile-time R (a power of two <= 16) and unrolled r < R
template <int R> NRB_DEV double2 unit_rot(int r)
{
    const int m = r * (16 / R);          // angle = m * pi/16, 0 <= m < 16; sin(m pi/16) = cos(|8 - m| pi/16)
    return make_double2(cos_pi16(m), -cos_pi16(m < 8 ? 8 - m : m - 8));
}

// One bin of NR's realft untangling: out_k from (Z_k, Z_{N-k}) and t = exp(-i pi k / N), valid for every
// 1 <= k <= N-1 (the pair form `untangle_pair` produces out_k and out_{N-k} together; this form lets each
// thread finish the bins it already holds in registers).
template <int DIR>
NRB_DEV double2 untangle_bin(double2 a, double2 b, double2 t)
{
    const double2 w = DIR > 0 ? make_double2(t.x, -t.y) : t;
    const double c2 = DIR > 0 ? -0.5 : 0.5;
    const double h1r = 0.5 * (a.x + b.x), h1i = 0.5 * (a.y - b.y);
    const double h2r = -c2 * (a.y + b.y), h2i = c2 * (a.x - b.x);
    return make_double2(h1r + w.x * h2r - w.y * h2i, h1i + w.x * h2i + w.y * h2r);
}
NRB_DEV double2 dc_inverse_speq(double2 g0, double2 gn);

// forward REAL passes whose last stage has one butterfly per thread and at most 32 butterflies per line
// untangle in registers: the partner bins N-k live in lane (NB - j) of the same warp (warp shuffles)
template <int LOG2N, int LAYOUT> NRB_HD constexpr bool real_fwd_in_registers()
{
    return LAYOUT == LAYOUT_ROW && radix_plan(LOG2N).r[radix_plan(LOG2N).nst - 1] == points_per_thread(LAYOUT, LOG2N) &&
           ((1 << LOG2N) / radix_plan(LOG2N).r[radix_plan(LOG2N).nst - 1]) <= 32;
}

// ROW tiles give every line to a fixed group of TPL = N/PPT threads for the whole transform (thread ->
// line = tid / TPL, butterflies (tid % TPL) + i*TPL).  When that group fits a warp (N <= 32*PPT) nothing
// ever crosses warps, so the barriers between stages are __syncwarp() instead of __syncthreads().
template <int LOG2N, int LAYOUT> NRB_HD constexpr bool line_owned()
{
    return LAYOUT == LAYOUT_ROW && (1 << LOG2N) >= points_per_thread(LAYOUT, LOG2N);
}
template <int LOG2N, int LAYOUT> NRB_HD constexpr bool warp_owned()
{
    return line_owned<LOG2N, LAYOUT>() && ((1 << LOG2N) / points_per_thread(LAYOUT, LOG2N)) <= 32;
}
template <int LOG2N, int LAYOUT> NRB_DEV void stage_sync()
{
    if (warp_owned<LOG2N, LAYOUT>()) NRB_SYNCWARP();
    else NRB_SYNC();
}

// ------------------------------------------------------------------ one Stockham stage
// SRC_G: inputs come from global memory (else shared); DST_G: outputs go to global memory.
// SIMPLE (compile-time copy of PassParams::simple, see simple_ok / pass_is_simple): the element index is not split
// (offset = n * es, with es = 1 for ROW) -- then a thread's global addresses are one base pointer plus multiples of a
// fixed step (immediate offsets for ROW), instead of ~13 integer instructions per element for the general form.
// ncu on the N = 8192 pass: 45 % of all executed instructions were address arithmetic (profiles/r01_tuning.md #36).
template <int LOG2N, int LAYOUT> NRB_HD constexpr bool simple_ok()
{
    // every butterfly of a thread must belong to the same line
    return LAYOUT == LAYOUT_ROW ? line_owned<LOG2N, LAYOUT>()
                                : (cta_threads(LOG2N, LAYOUT) % lines_per_tile(LOG2N, LAYOUT)) == 0;
}

// TMA_IO (fft_tma.cuh: the tile is brought into and taken out of shared memory by bulk tensor copies, so every stage
// runs shared -> shared): bit 0 = this stage reads the tile as it came from global memory, bit 1 = it writes the tile
// that goes back to global memory; those are the two places where the isign = +1 re/im swap has to happen instead.
template <int LOG2N, int LAYOUT, int DIR, int VARIANT, int S, bool SRC_G, bool DST_G, bool SIMPLE = false, int TMA_IO = 0>
NRB_DEV void fft_stage(const PassParams &P, double2 *sm, unsigned tile, int tid)
{
    typedef Geo<LOG2N, LAYOUT, VARIANT> G;
    constexpr int R = radix_plan(LOG2N).r[S];
    constexpr int NS = stage_ns(LOG2N, S);
    constexpr int NB = G::N / R;           // butterflies per line
    constexpr int BPT = G::PPT / R;        // butterflies per thread
    static_assert(BPT >= 1, "radix larger than points per thread");
    constexpr bool FAST = SIMPLE && simple_ok<LOG2N, LAYOUT>() && (VARIANT == VAR_PLAIN || VARIANT == VAR_XPOSE);
    // ROW shared-memory index n + (n >> 3): for the strides of a butterfly's legs the pad term is a compile-time
    // constant -- phys(l, a + r*D) = phys(l, a) + r*D + (r*D >> 3) whenever D is a multiple of 8 or (writes) a + r*D
    // stays inside the 8-block of a (checked exhaustively for every radix plan) -- so the legs are immediate offsets
    constexpr bool ROW_RD_CONST = LAYOUT == LAYOUT_ROW && (NB % 8) == 0;
    constexpr bool ROW_WR_CONST = LAYOUT == LAYOUT_ROW && (S == 0 || R == 8);

    double2 v[BPT][R];
    int ln[BPT], jj[BPT];

#pragma unroll
    for (int i = 0; i < BPT; ++i) {
        const int b = tid + i * G::NT;
        if (LAYOUT == LAYOUT_COL) { ln[i] = b & (G::L - 1); jj[i] = b / G::L; }
        else if (line_owned<LOG2N, LAYOUT>()) {
            constexpr int TPL = (G::N / G::PPT) > 0 ? (G::N / G::PPT) : 1;
            ln[i] = tid / TPL; jj[i] = (tid & (TPL - 1)) + i * TPL;
        }
        else                      { jj[i] = b & (NB - 1);   ln[i] = b / NB; }
    }

    // ---- gather
#pragma unroll
    for (int i = 0; i < BPT; ++i) {
        if (SRC_G) {
            const u64 q = P.q_begin + (u64)tile * G::L + (u64)ln[i];
            const bool ok = q < P.q_end;
            const double2 *src = P.in + line_base(q, P.in_s0, P.in_s1, P.in_s2, P.logA, P.logB);
            if (VARIANT == VAR_REAL && DIR < 0 && NRB_REAL_INV_DIRECT) {
                // (measured 10 % slower than the shared-memory pre-pass on B200, kept for reference, off)
                // inverse real transform: the untangling (Real_FT.rs:145-176) is applied on the way in.
                // Bin k needs F_k and F_{N-k}: both are read straight from global memory (the partner read
                // hits L2 -- the same CTA reads it as its own element), so no shared-memory round trip.
                // (own loads first, then the partner loads in groups of at most 4 to bound register use)
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    v[i][r] = make_double2(0.0, 0.0);
                    if (ok) v[i][r] = NRB_LDS(src + (i64)(jj[i] + r * NB) * P.in_es);
                }
                constexpr int CH = R < 4 ? R : 4;
#pragma unroll
                for (int r0 = 0; r0 < R; r0 += CH) {
                    double2 pb[CH];
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        const int k = jj[i] + (r0 + c) * NB;
                        pb[c] = make_double2(0.0, 0.0);
                        if (ok && k != 0) pb[c] = NRB_LDS(src + (i64)(G::N - k) * P.in_es);
                    }
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        const int r = r0 + c;
                        const int k = jj[i] + r * NB;
                        if (k == 0) {
                            const double2 g0 = v[i][r];
                            if (P.real_mode == REAL_SPEQ) v[i][r] = ok ? dc_inverse_speq(g0, P.speq[q]) : g0;
                            else v[i][r] = make_double2(0.5 * (g0.x + g0.y), 0.5 * (g0.x - g0.y));
                        } else {
                            v[i][r] = untangle_bin<-1>(v[i][r], pb[c], NRB_LDG(P.rtw + k));
                        }
                    }
                }
            } else if (!(FAST && SRC_G)) {
                if (P.in_peer_on) {   // exchange blocks through the pointer table: local receive buffer or a peer's send buffer (NVLink loads)
                    const i64 lb = line_base(q, P.in_s0, P.in_s1, P.in_s2, P.logA, P.logB);
                    const int zsel = (((unsigned)q & P.peer_zmask) >= P.peer_zthr) ? 8 : 0;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int n = jj[i] + r * NB;
                        double2 x = make_double2(0.0, 0.0);
                        if (ok) x = NRB_LDS(P.in_peer[zsel + (n >> P.in_eshift)] + lb + (i64)(n & ((1 << P.in_eshift) - 1)) * P.in_es);
                        v[i][r] = io_swap<DIR>(x);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        double2 x = make_double2(0.0, 0.0);
                        if (ok) x = NRB_LDS(src + elem_off(jj[i] + r * NB, P.in_es, P.in_eshift, P.in_es_hi));
                        v[i][r] = io_swap<DIR>(x);
                    }
                }
            }
        } else if (ROW_RD_CONST) {
            const double2 *sp = sm + G::phys(ln[i], jj[i]);
#pragma unroll
            for (int r = 0; r < R; ++r) v[i][r] = sp[r * NB + ((r * NB) >> 3)];
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double2 x = sm[G::phys(ln[i], jj[i] + r * NB)];
                v[i][r] = (TMA_IO & 1) ? io_swap<DIR>(x) : x;
            }
        }
    }
    if (SRC_G && FAST) {
        // one line per thread: base pointer once, then fixed steps (jj[i] = jj[0] + i * JSTEP)
        constexpr int JSTEP = LAYOUT == LAYOUT_ROW ? (G::N / G::PPT) : (G::NT / G::L);
        const u64 q = P.q_begin + (u64)tile * G::L + (u64)ln[0];
        const bool ok = q < P.q_end;
        const double2 *src = P.in + line_base(q, P.in_s0, P.in_s1, P.in_s2, P.logA, P.logB);
        if (LAYOUT == LAYOUT_ROW) {
            src += jj[0];
#pragma unroll
            for (int i = 0; i < BPT; ++i) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    double2 x = make_double2(0.0, 0.0);
                    if (ok) x = NRB_LDS(src + (i * JSTEP + r * NB));
                    v[i][r] = io_swap<DIR>(x);
                }
            }
        } else {
            const i64 es = P.in_es;
            src += (i64)jj[0] * es;
#pragma unroll
            for (int i = 0; i < BPT; ++i) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    double2 x = make_double2(0.0, 0.0);
                    if (ok) x = NRB_LDS(src + (i64)(i * JSTEP + r * NB) * es);
                    v[i][r] = io_swap<DIR>(x);
                }
            }
        }
    }

    // ---- twiddle + butterfly
#ifndef NRB_SKIP_MATH   /* timing experiments only: memory skeleton of the pass */
#pragma unroll
    for (int i = 0; i < BPT; ++i) {
        if (NS > 1) {
            const int jm = jj[i] & (NS - 1);
            const double2 *tp = P.tw + stage_tw_off(LOG2N, S) + jm * (R - 1);
            if ((LAYOUT == LAYOUT_ROW ? NRB_TW_POWERS : NRB_TW_POWERS_COL) && R >= 4) {
                // ROW kernels are L1/TEX-bound (every lane needs its own twiddles): load w^1 only and build
                // the other powers with complex multiplies on the half-idle FP64 pipe (<= 3 products deep)
                double2 w[R];
                w[1] = NRB_LDG(tp);
                if (NRB_TW_LOADS == 3 && R == 8) {
                    w[2] = NRB_LDG(tp + 1);
                    w[4] = NRB_LDG(tp + 3);
                    w[3] = cmul(w[2], w[1]);
                    w[5] = cmul(w[4], w[1]);
                    w[6] = cmul(w[4], w[2]);
                    w[7] = cmul(w[4], w[3]);
                } else {
                    w[2] = cmul(w[1], w[1]);
                    w[3] = cmul(w[2], w[1]);
                    if (R >= 8) {
                        w[4] = cmul(w[2], w[2]);
                        w[5] = cmul(w[4], w[1]);
                        w[6] = cmul(w[3], w[3]);
                        w[7] = cmul(w[4], w[3]);
                    }
                }
                if (R >= 16) {
#pragma unroll
                    for (int r = 8; r < R; ++r) w[r] = cmul(w[r - 4], w[4]);
                }
#pragma unroll
                for (int r = 1; r < R; ++r) v[i][r] = cmul(v[i][r], w[r]);
            } else {
#pragma unroll
                for (int r = 1; r < R; ++r) v[i][r] = cmul(v[i][r], NRB_LDG(tp + (r - 1)));
            }
        }
        Bfly<R>::run(v[i]);
    }
#endif

    if (!SRC_G && !DST_G) stage_sync<LOG2N, LAYOUT>();   // everyone has read before anyone overwrites

    // ---- scatter
#pragma unroll
    for (int i = 0; i < BPT; ++i) {
        const int jm = jj[i] & (NS - 1);
        const int kb = (jj[i] - jm) * R + jm;
        if (DST_G && VARIANT == VAR_REAL && DIR > 0) {
            // forward real transform, untangling in registers (see real_fwd_in_registers): this thread holds
            // Z_k for k = j + r*NB; Z_{N-k} is register R-1-r of lane NB-j (for j = 0: own register R-r).
            const u64 q = P.q_begin + (u64)tile * G::L + (u64)ln[i];
            const bool ok = q < P.q_end;
            const int j = jj[i];
            const int lane = tid & 31;
            const int partner = (lane & ~(NB - 1)) | ((NB - j) & (NB - 1));
            double2 *dst = P.out + line_base(q, P.out_s0, P.out_s1, P.out_s2, P.logA, P.logB);
            const double2 rt0 = NRB_LDG(P.rtw + j);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                // value this lane publishes under index R-1-r; the warp runs in lockstep, so each bin is
                // exchanged, untangled and stored before the next one (keeps register use at one butterfly)
                const double2 mine = (j == 0) ? v[i][(R - r) & (R - 1)] : v[i][R - 1 - r];
                const double2 pz = make_double2(NRB_SHFL(mine.x, partner), NRB_SHFL(mine.y, partner));
                const int k = j + r * NB;
                const double2 zk = cswap(v[i][r]);
                if (k == 0) {
                    const double f0 = zk.x + zk.y, fn = zk.x - zk.y;
                    if (ok) {
                        if (P.real_mode == REAL_SPEQ) { NRB_STS(dst, make_double2(f0, 0.0)); P.speq[q] = make_double2(fn, 0.0); }
                        else NRB_STS(dst, make_double2(f0, fn));
                    }
                } else {
                    // exp(-i pi k/N), k = j + r*NB: one table load per thread, the other R-1 by the fixed
                    // rotations exp(-i pi r/R)
                    const double2 t = NRB_TW_POWERS ? cmul(rt0, unit_rot<R>(r)) : NRB_LDG(P.rtw + k);
                    const double2 f = untangle_bin<1>(zk, cswap(pz), t);
                    if (ok) NRB_STS(dst + (i64)k * P.out_es, f);
                }
            }
        } else if (DST_G && FAST && !P.out_peer_on) {
            const u64 q = P.q_begin + (u64)tile * G::L + (u64)ln[0];
            if (q < P.q_end) {
                double2 *dst = P.out + line_base(q, P.out_s0, P.out_s1, P.out_s2, P.logA, P.logB);
                double2 tw = make_double2(1.0, 0.0), tw_step = make_double2(1.0, 0.0);
                if (P.tw_on) {
                    const unsigned q1 = line_q1(P, q);
                    tw = fourstep_tw_m(P, q1 * (unsigned)kb);
                    tw_step = fourstep_tw_m(P, q1 * (unsigned)NS);
                }
                const i64 es = LAYOUT == LAYOUT_ROW ? (i64)1 : P.out_es;
                dst += (i64)kb * es;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    double2 y = v[i][r];
                    if (P.tw_on) { y = cmul(y, tw); tw = cmul(tw, tw_step); }
                    if (LAYOUT == LAYOUT_ROW) NRB_STS(dst + r * NS, io_swap<DIR>(y));
                    else NRB_STS(dst + (i64)(r * NS) * es, io_swap<DIR>(y));
                }
            }
        } else if (DST_G) {
            const u64 q = P.q_begin + (u64)tile * G::L + (u64)ln[i];
            if (q < P.q_end) {
                const i64 lb = line_base(q, P.out_s0, P.out_s1, P.out_s2, P.logA, P.logB);
                double2 *dst = P.out + lb;
                // four-step twiddle W^(q1*k), k = kb + r*NS: geometric in r -> two table look-ups
                // (base and ratio) and R-1 complex multiplies instead of R look-ups
                double2 tw = make_double2(1.0, 0.0), tw_step = make_double2(1.0, 0.0);
                if (P.tw_on) {
                    const unsigned q1 = line_q1(P, q);
                    tw = fourstep_tw_m(P, q1 * (unsigned)kb);
                    tw_step = fourstep_tw_m(P, q1 * (unsigned)NS);
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int k = kb + r * NS;
                    double2 y = v[i][r];
                    if (P.tw_on) { y = cmul(y, tw); tw = cmul(tw, tw_step); }
                    if (P.out_peer_on) {   // store straight into the owning peer's receive buffer (NVLink), or the local send buffer
                        const int zsel = (((unsigned)q & P.peer_zmask) >= P.peer_zthr) ? 8 : 0;
                        double2 *pd = P.out_peer[zsel + (k >> P.out_eshift)] + P.out_peer_off + lb +
                                      (i64)(k & ((1 << P.out_eshift) - 1)) * P.out_es;
                        NRB_STS(pd, io_swap<DIR>(y));
                    } else {
                        NRB_STS(dst + elem_off(k, P.out_es, P.out_eshift, P.out_es_hi), io_swap<DIR>(y));
                    }
                }
            }
        } else if (ROW_WR_CONST) {
            double2 *sp = sm + G::phys(ln[i], kb);
#pragma unroll
            for (int r = 0; r < R; ++r) sp[r * NS + ((r * NS) >> 3)] = v[i][r];
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) sm[G::phys(ln[i], kb + r * NS)] = (TMA_IO & 2) ? io_swap<DIR>(v[i][r]) : v[i][r];
        }
    }
}

// run stages FIRST..NST-1; stage FIRST reads global iff SRC_G0, the last stage writes global
// iff DST_GL.  Barriers: after every stage that wrote shared memory.
// TMA: mask of the re/im swaps a TMA-fed pass needs inside its stages (fft_tma.cuh): 1 = on the first stage's shared read
// (the tile came from global memory as is), 2 = on the last stage's shared write (the tile goes back to global as is)
template <int LOG2N, int LAYOUT, int DIR, int VARIANT, int S, bool SRC_G0, bool DST_GL, bool SIMPLE = false, int TMA = 0>
struct StageRunner {
    NRB_DEVM static void run(const PassParams &P, double2 *sm, unsigned tile, int tid)
    {
        constexpr int NST = radix_plan(LOG2N).nst;
        constexpr bool last = (S == NST - 1);
        constexpr bool src_g = (S == 0) && SRC_G0;
        constexpr bool dst_g = last && DST_GL;
        constexpr int tma_io = ((S == 0 && (TMA & 1)) ? 1 : 0) | ((last && (TMA & 2)) ? 2 : 0);
        fft_stage<LOG2N, LAYOUT, DIR, VARIANT, S, src_g, dst_g, SIMPLE, tma_io>(P, sm, tile, tid);
        if (!dst_g) stage_sync<LOG2N, LAYOUT>();
        StageRunner<LOG2N, LAYOUT, DIR, VARIANT, last ? -1 : S + 1, SRC_G0, DST_GL, SIMPLE, TMA>::run(P, sm, tile, tid);
    }
};
template <int LOG2N, int LAYOUT, int DIR, int VARIANT, bool SRC_G0, bool DST_GL, bool SIMPLE, int TMA>
struct StageRunner<LOG2N, LAYOUT, DIR, VARIANT, -1, SRC_G0, DST_GL, SIMPLE, TMA> {
    NRB_DEVM static void run(const PassParams &, double2 *, unsigned, int) {}
};

// ------------------------------------------------------------------ real untangle (NR realft)
// Forward (c2 = -0.5, w = exp(+i pi k/N)):  h1 = (Zk + conj Zm)/2, h2 = -(i/2)(Zk - conj Zm),
//   Fk = h1 + w h2, Fm = conj(h1 - w h2)               [Real_FT.rs:66-74 with 0-based pairs]
// Inverse (c2 = +0.5, w = exp(-i pi k/N)):  h2 = +(i/2)(Fk - conj Fm), same recombination
//                                                       [Real_FT.rs:162-170]
template <int DIR>
NRB_DEV void untangle_pair(double2 a, double2 b, double2 t /* exp(-i pi k/N) */, double2 &oa, double2 &ob)
{
    const double2 w = DIR > 0 ? cconj(t) : t;
    const double c2 = DIR > 0 ? -0.5 : 0.5;
    const double h1r = 0.5 * (a.x + b.x), h1i = 0.5 * (a.y - b.y);
    const double h2r = -c2 * (a.y + b.y), h2i = c2 * (a.x - b.x);
    const double tr = w.x * h2r - w.y * h2i, ti = w.x * h2i + w.y * h2r;
    oa = make_double2(h1r + tr, h1i + ti);
    ob = make_double2(h1r - tr, -h1i + ti);
}

// DC / Nyquist element.  Forward: Z0 = (a, b) -> F0 = a + b, FN = a - b  (Real_FT.rs:43-45).
// Inverse packed: (F0, FN) -> Z0 = ((F0+FN)/2, (F0-FN)/2)               (Real_FT.rs:133-135).
// Inverse speq (rlft3): g0, gN complex planes -> Z0 = [(1+i) g0 + (1-i) conj(gN)] / 2, which
// is what NR's i3 == 1 branch (Real_FT3.rs:73-87) followed by the x/y transforms amounts to.
NRB_DEV double2 dc_inverse_speq(double2 g0, double2 gn)
{
    return make_double2(0.5 * ((g0.x - g0.y) + (gn.x - gn.y)), 0.5 * ((g0.x + g0.y) - (gn.x + gn.y)));
}

// ------------------------------------------------------------------ L2 prefetch of a later tile
// One prefetch per 128-byte piece of the tile's input footprint (exact addresses of lines that exist).
template <int LOG2N, int LAYOUT, int VARIANT>
NRB_DEV void prefetch_tile(const PassParams &P, unsigned tile, int tid)
{
    typedef Geo<LOG2N, LAYOUT, VARIANT> G;
    const u64 q0 = P.q_begin + (u64)tile * G::L;
    if (LAYOUT == LAYOUT_COL) {
        constexpr int CPR = (G::L + 7) / 8;                 // 128-byte pieces per row (L lines of 16 bytes)
        for (int u = tid; u < G::N * CPR; u += G::NT) {
            const int r = u / CPR, c = u % CPR;
            const u64 q = q0 + (u64)c * 8;
            if (q < P.q_end)
                NRB_PREFETCH_L2(P.in + line_base(q, P.in_s0, P.in_s1, P.in_s2, P.logA, P.logB) + elem_off(r, P.in_es, P.in_eshift, P.in_es_hi));
        }
    } else {
        for (int u = tid; u < G::TILE / 8; u += G::NT) {
            const int e = u * 8, l = e >> LOG2N, n = e & (G::N - 1);
            const u64 q = q0 + (u64)l;
            if (q < P.q_end && G::N >= 8)
                NRB_PREFETCH_L2(P.in + line_base(q, P.in_s0, P.in_s1, P.in_s2, P.logA, P.logB) + elem_off(n, P.in_es, P.in_eshift, P.in_es_hi));
        }
    }
}

// ------------------------------------------------------------------ the pass body
// TILE_IN_SMEM (PLAIN and XPOSE, fft_tma.cuh): the tile has already been brought into shared memory by a bulk tensor copy, in
// the layout of Geo::phys, so the first stage reads shared memory too; the transposing epilogue is unchanged.
template <int LOG2N, int LAYOUT, int DIR, int VARIANT, bool SIMPLE = false, bool TILE_IN_SMEM = false>
NRB_DEV void fft_pass_body(const PassParams &P, double2 *sm, unsigned tile, int tid)
{
    typedef Geo<LOG2N, LAYOUT, VARIANT> G;
    if (P.tile_nsel > 0)   // this launch covers a subset of the pass's tiles (PassParams::tile_run)
        tile = (((((tile >> P.tile_run) << P.tile_nsel) | (unsigned)P.tile_sel) << P.tile_run) | (tile & ((1u << P.tile_run) - 1u)));

    if (VARIANT == VAR_PLAIN) {
        StageRunner<LOG2N, LAYOUT, DIR, VARIANT, 0, !TILE_IN_SMEM, true, SIMPLE, TILE_IN_SMEM ? 1 : 0>::run(P, sm, tile, tid);
        return;
    }

    if (VARIANT == VAR_XPOSE) {
        // COL layout; last stage to shared memory, then row-like (line-contiguous) store with the
        // four-step twiddle W^(q1(l)*k).  A thread's elements share k (C distinct values when N > NT)
        // and walk the lines with a fixed step D, so the twiddle is a geometric sequence per k.
        StageRunner<LOG2N, LAYOUT, DIR, VARIANT, 0, !TILE_IN_SMEM, false, SIMPLE, TILE_IN_SMEM ? 1 : 0>::run(P, sm, tile, tid);
        constexpr int C = (G::N > G::NT) ? G::N / G::NT : 1;
        constexpr int D = (G::NT >= G::N) ? G::NT / G::N : 1;
        constexpr int E = G::PPT / C;
        const u64 q_tile = P.q_begin + (u64)tile * G::L;
        const bool linear = P.logB == 0 && ((1ull << P.logA) >= (u64)G::L);   // no wrap of q1 inside the tile
        const bool geometric = P.tw_on && linear;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int idx0 = tid + c * G::NT;
            const int k = idx0 & (G::N - 1), l0 = idx0 >> LOG2N;
            double2 tw = make_double2(1.0, 0.0), tw_step = make_double2(1.0, 0.0);
            if (geometric) {
                tw = fourstep_tw_m(P, line_q1(P, q_tile + (u64)l0) * (unsigned)k);
                tw_step = fourstep_tw_m(P, (unsigned)(D * k));
            }
            if (SIMPLE && linear) {
                // the lines of a tile advance q1 only: one line base, then a fixed step per line (SIMPLE, fft_stage)
                double2 *dst = P.out + line_base(q_tile + (u64)l0, P.out_s0, P.out_s1, P.out_s2, P.logA, P.logB) + (i64)k * P.out_es;
                const i64 step = (i64)D * P.out_s1;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int l = l0 + e * D;
                    if (q_tile + (u64)l < P.q_end) {
                        double2 y = sm[G::phys(l, k)];
                        if (geometric) y = cmul(y, tw);
                        NRB_STS(dst + (i64)e * step, io_swap<DIR>(y));
                    }
                    tw = cmul(tw, tw_step);
                }
                continue;
            }
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int l = l0 + e * D;
                const u64 q = q_tile + (u64)l;
                if (q < P.q_end) {
                    double2 y = sm[G::phys(l, k)];
                    if (geometric) y = cmul(y, tw);
                    else if (P.tw_on) y = cmul(y, fourstep_tw(P, q, (unsigned)k));
                    double2 *dst = P.out + line_base(q, P.out_s0, P.out_s1, P.out_s2, P.logA, P.logB);
                    NRB_STS(dst + (i64)k * P.out_es, io_swap<DIR>(y));
                }
                tw = cmul(tw, tw_step);
            }
        }
        return;
    }

    if (VARIANT == VAR_REAL) {
        constexpr int HALF = G::N / 2;   // pair items per line (k = 0 handles DC, Nyquist, middle)
        constexpr bool OWN = NRB_PRE_OWN && line_owned<LOG2N, LAYOUT>() && (G::N / 2) >= (G::N / G::PPT) && (G::N / G::PPT) >= 1;
        constexpr int TPL = OWN ? G::N / G::PPT : 1;
        constexpr int ITEMS = (G::L * HALF + G::NT - 1) / G::NT;
        if (DIR > 0 && real_fwd_in_registers<LOG2N, LAYOUT>()) {
            // c2c with the untangling done in registers by the last stage
            StageRunner<LOG2N, LAYOUT, DIR, VARIANT, 0, true, true>::run(P, sm, tile, tid);
        } else if (DIR > 0) {
            // c2c (swapped domain) -> shared memory -> untangle -> global
            StageRunner<LOG2N, LAYOUT, DIR, VARIANT, 0, true, false>::run(P, sm, tile, tid);
#pragma unroll 4
            for (int i = 0; i < ITEMS; ++i) {
                const int idx = tid + i * G::NT;
                if (idx >= G::L * HALF) break;
                const int k = idx & (HALF - 1), l = idx / HALF;
                const u64 q = P.q_begin + (u64)tile * G::L + (u64)l;
                if (q >= P.q_end) continue;
                double2 *dst = P.out + line_base(q, P.out_s0, P.out_s1, P.out_s2, P.logA, P.logB);
                if (k == 0) {
                    const double2 z0 = cswap(sm[G::phys(l, 0)]);
                    const double f0 = z0.x + z0.y, fn = z0.x - z0.y;
                    if (P.real_mode == REAL_SPEQ) {
                        dst[0] = make_double2(f0, 0.0);
                        P.speq[q] = make_double2(fn, 0.0);
                    } else {
                        dst[0] = make_double2(f0, fn);
                    }
                    if (HALF >= 1 && G::N >= 2) dst[(i64)HALF * P.out_es] = cswap(sm[G::phys(l, HALF)]);
                } else {
                    const double2 a = cswap(sm[G::phys(l, k)]), b = cswap(sm[G::phys(l, G::N - k)]);
                    double2 oa, ob;
                    untangle_pair<DIR>(a, b, NRB_LDG(P.rtw + k), oa, ob);
                    NRB_STS(dst + (i64)k * P.out_es, oa);
                    NRB_STS(dst + (i64)(G::N - k) * P.out_es, ob);
                }
            }
        } else {
            // global -> untangle -> shared memory -> c2c -> global.  All pair loads are issued
            // before any is used (k = 0 pairs element 0 with the untouched middle element N/2).
            double2 a[ITEMS], b[ITEMS];
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                // line-owned tiles: the thread group of a line also prepares that line (warp-level sync)
                const int idx = OWN ? (tid / TPL) * HALF + (tid & (TPL - 1)) + i * TPL : tid + i * G::NT;
                const int k = idx & (HALF - 1), l = idx / HALF;
                const u64 q = P.q_begin + (u64)tile * G::L + (u64)l;
                const bool ok = (idx < G::L * HALF) && (q < P.q_end);
                const double2 *src = P.in + line_base(q, P.in_s0, P.in_s1, P.in_s2, P.logA, P.logB);
                a[i] = make_double2(0.0, 0.0);
                b[i] = make_double2(0.0, 0.0);
                if (ok) {
                    a[i] = NRB_LDS(src + (i64)k * P.in_es);
                    if (G::N >= 2) b[i] = NRB_LDS(src + (i64)(k == 0 ? HALF : G::N - k) * P.in_es);
                }
            }
            // exp(-i pi k / N) of the thread's items: with the flat mapping k advances by NT from item to item (and starts over
            // on the next line), so one table load and the fixed rotations exp(-i pi m NT / N) replace one load per item
            // (the pass is L1/TEX-bound: ncu 92 %)
            constexpr bool TWROT = NRB_REAL_INV_TWROT && !OWN && HALF >= 1 && (G::NT <= HALF ? (HALF % G::NT == 0 && G::N / G::NT <= 16) : (G::NT % HALF == 0));
            constexpr int MROT = (G::NT <= HALF && HALF >= 1) ? HALF / (G::NT > 0 ? G::NT : 1) : 1;
            constexpr int RROT = (G::N / G::NT) >= 1 && (G::N / G::NT) <= 16 ? (G::N / G::NT) : 1;
            double2 rt_base = make_double2(1.0, 0.0);
            if (TWROT) rt_base = NRB_LDG(P.rtw + (tid & (HALF - 1)));
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const int idx = OWN ? (tid / TPL) * HALF + (tid & (TPL - 1)) + i * TPL : tid + i * G::NT;
                if (idx >= G::L * HALF) break;
                const int k = idx & (HALF - 1), l = idx / HALF;
                if (k == 0) {
                    const u64 q = P.q_begin + (u64)tile * G::L + (u64)l;
                    double2 z0 = make_double2(0.5 * (a[i].x + a[i].y), 0.5 * (a[i].x - a[i].y));
                    if (P.real_mode == REAL_SPEQ)
                        z0 = (q < P.q_end) ? dc_inverse_speq(a[i], P.speq[q]) : make_double2(0.0, 0.0);
                    sm[G::phys(l, 0)] = z0;
                    if (G::N >= 2) sm[G::phys(l, HALF)] = b[i];
                } else {
                    double2 oa, ob;
                    const double2 rt = !TWROT ? NRB_LDG(P.rtw + k) : ((i % MROT) == 0 ? rt_base : cmul(rt_base, unit_rot<RROT>(i % MROT)));
                    untangle_pair<DIR>(a[i], b[i], rt, oa, ob);
                    sm[G::phys(l, k)] = oa;
                    sm[G::phys(l, G::N - k)] = ob;
                }
            }
            if (OWN) stage_sync<LOG2N, LAYOUT>(); else NRB_SYNC();
            StageRunner<LOG2N, LAYOUT, DIR, VARIANT, 0, false, true>::run(P, sm, tile, tid);
        }
        return;
    }
}

} // namespace nrb
