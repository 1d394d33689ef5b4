// conv_mid.cuh -- the fused middle of convlv / correl / autocorrel_fast for lines longer than one tile.
//
// The two-pass transforms of those routines keep the spectrum in transposed order (plan.cpp, conv_split): bin
// kf + F*kr of a signal sits at position kf*REST + kr, and the realft untangling pairs row kf with row F - kf.
// Without this kernel the middle of the pipeline is three launches and three HBM round trips per signal:
//   contiguous REST-point forward pass per row | untangle * spectral op * re-tangle (aux_spectral_zt) | contiguous
//   REST-point inverse pass per row (with the four-step twiddle on its output).
// Here one CTA owns the row pair (kf, F - kf) of one signal -- rows 0 and F/2, which pair within themselves, share a
// CTA -- and does all three steps on the pair while it is on chip: 5 -> 3 passes per signal for convlv, 7 -> 5 for
// correl, 5 -> 3 for autocorrel_fast ("fuse the spectral multiply into the last forward pass", BASELINE.json).
// The second operand (convlv: the response spectrum shared by the whole batch and L2-resident; correl: the other
// signal's finished spectrum) is read from global memory in the same transposed order.
//
// Reference: Convolve.rs:96 (realft forward), :112-129 (spectral multiply / divide with the 1/no2 scale), :131 (realft
// inverse); Correlation.rs:73-74, :91-92; NR realft untangling Real_FT.rs:49-80 / :145-176 (oracle ledger D1, D6, D7).
//
// 2*REST points per CTA, 16 per thread (REST = 4096: 512 threads, 144 KiB of shared memory, one CTA per SM); the
// Stockham stages and their register layout are those of fft_pass2.cuh (Stage2, v2_compute), the exchange moves
// 16-byte elements through the padded buffer of fft_pass.cuh.
#pragma once
#include "aux_kernels.cuh"
#include "fft_pass2.cuh"

namespace nrb {


// points per thread of the fused middle kernel: 16 = 512 threads of 128 registers for 4096-point rows, 8 = 1024 threads
// of 64 registers (twice the warps to hide the latency of the kernel's serial phases)
#ifndef NRB_MID_PPT
#define NRB_MID_PPT 8      /* measured: 0.670 -> 0.612 ms per 16 signals of 2^22 (profiles/r02_tuning.md #51) */
#endif
template <int LOG2R> struct GeoM {
    static constexpr int LOG2N = LOG2R, LAYOUT = LAYOUT_ROW, VARIANT = VAR_PLAIN;
    static constexpr int N = 1 << LOG2R;
    static constexpr int L = 2;
    static constexpr int TILE = 2 * N;
    static constexpr int PPT = (2 << LOG2R) / NRB_MID_PPT > 1024 ? 16 : ((1 << LOG2R) >= NRB_MID_PPT ? NRB_MID_PPT : 16);
    static constexpr int NT = TILE / PPT;
    static constexpr int NST = radix_plan(LOG2R).nst;
    static constexpr int LP = N + (N >> 3);
    static constexpr size_t SMEM_BYTES = (size_t)L * LP * sizeof(double2);
    static_assert(N >= PPT, "conv_mid: rows shorter than the points per thread are not built");
    static_assert(NST >= 2, "conv_mid needs at least two stages");
    NRB_DEVM static int phys(int l, int n) { return l * LP + n + (n >> 3); }
};

// scatter of stage S (Stockham autosort order) into the exchange buffer
template <class G, int S>
NRB_DEV void mid_scatter(double2 *E, int tid, const double2 *v)
{
    typedef Stage2<G, S> A;
#pragma unroll
    for (int i = 0; i < A::BPT; ++i) {
        int ln, jj;
        A::coords(tid, i, ln, jj);
        const int jm = jj & (A::NS - 1);
        double2 *sp = E + G::phys(ln, (jj - jm) * A::R + jm);
#pragma unroll
        for (int r = 0; r < A::R; ++r) sp[r * A::NS + ((r * A::NS) >> 3)] = v[i * A::R + r];   // constant pad term: fft_stage
    }
}
// gather of stage S from the exchange buffer
template <class G, int S>
NRB_DEV void mid_gather(const double2 *E, int tid, double2 *v)
{
    typedef Stage2<G, S> B;
#pragma unroll
    for (int i = 0; i < B::BPT; ++i) {
        int ln, jj;
        B::coords(tid, i, ln, jj);
#pragma unroll
        for (int r = 0; r < B::R; ++r) v[i * B::R + r] = E[G::phys(ln, jj + r * B::NB)];
    }
}

// NRB_TW_PREFETCH = 1: the table look-ups of stage S + 1 (one twiddle per butterfly) are issued before the exchange that
// precedes it, so their L1 / L2 latency hides behind the scatter, the two barriers and the gather (ncu on conv_mid: the first
// FP64 instruction of every stage waits on that load, 9 % of the warp samples; profiles/r02_ncu_full_conv_mid.md)
#ifndef NRB_TW_PREFETCH
#define NRB_TW_PREFETCH 0
#endif
template <class G, int S, class PP>
NRB_DEV void v2_tw_load(const PP &P, int tid, double2 *w1)
{
    typedef Stage2<G, S> T;
#pragma unroll
    for (int i = 0; i < T::BPT; ++i) {
        int ln, jj;
        T::coords(tid, i, ln, jj);
        const int jm = jj & (T::NS - 1);
        w1[i] = NRB_LDG(P.tw + stage_tw_off(G::LOG2N, S) + jm * (T::R - 1));
    }
}
// v2_compute with the first twiddle of every butterfly already in registers (radix >= 4 stages with NS > 1)
template <class G, int S, class PP>
NRB_DEV void v2_compute_pre(const PP &P, int tid, double2 *v, const double2 *w1)
{
    typedef Stage2<G, S> T;
    constexpr int R = T::R;
#pragma unroll
    for (int i = 0; i < T::BPT; ++i) {
        double2 w[R];
        w[1] = w1[i];
        w[2] = cmul(w[1], w[1]);
        w[3] = cmul(w[2], w[1]);
        if (R >= 8) {
            w[4] = cmul(w[2], w[2]);
            w[5] = cmul(w[4], w[1]);
            w[6] = cmul(w[3], w[3]);
            w[7] = cmul(w[4], w[3]);
        }
#pragma unroll
        for (int r = 1; r < R; ++r) v[i * R + r] = cmul(v[i * R + r], w[r]);
        Bfly<R>::run(v + i * R);
    }
}
// stages S .. NST-1 on registers with the exchanges between them; leaves the last stage's outputs in v
template <class G, int S> struct ChainM {
    static constexpr int SN = S + 1 < G::NST ? S + 1 : 0;
    // prefetch applies when the next stage exists, is radix 4 or 8 and has twiddles (NS > 1 always holds for S + 1 >= 1)
    static constexpr bool PRE = NRB_TW_PREFETCH && S + 1 < G::NST && Stage2<G, SN>::R >= 4 && Stage2<G, SN>::R <= 8;
    template <class PP> NRB_DEVM static void run(const PP &M, double2 *E, int tid, double2 *v)
    {
        v2_compute<G, S>(M, tid, v);
        if (S + 1 < G::NST) {
            double2 w1[Stage2<G, SN>::BPT];
            if (PRE) v2_tw_load<G, SN>(M, tid, w1);
            mid_scatter<G, S>(E, tid, v);
            NRB_SYNC();
            mid_gather<G, SN>(E, tid, v);
            NRB_SYNC();
            if (PRE) ChainM<G, (S + 1 < G::NST ? S + 1 : -1)>::run_pre(M, E, tid, v, w1);
            else ChainM<G, (S + 1 < G::NST ? S + 1 : -1)>::run(M, E, tid, v);
        }
    }
    // the same with this stage's first twiddles already loaded
    template <class PP> NRB_DEVM static void run_pre(const PP &M, double2 *E, int tid, double2 *v, const double2 *w1_this)
    {
        v2_compute_pre<G, S>(M, tid, v, w1_this);
        if (S + 1 < G::NST) {
            double2 w1[Stage2<G, SN>::BPT];
            if (PRE) v2_tw_load<G, SN>(M, tid, w1);
            mid_scatter<G, S>(E, tid, v);
            NRB_SYNC();
            mid_gather<G, SN>(E, tid, v);
            NRB_SYNC();
            if (PRE) ChainM<G, (S + 1 < G::NST ? S + 1 : -1)>::run_pre(M, E, tid, v, w1);
            else ChainM<G, (S + 1 < G::NST ? S + 1 : -1)>::run(M, E, tid, v);
        }
    }
};
template <class G> struct ChainM<G, -1> {
    template <class PP> NRB_DEVM static void run(const PP &, double2 *, int, double2 *) {}
    template <class PP> NRB_DEVM static void run_pre(const PP &, double2 *, int, double2 *, const double2 *) {}
};

template <int LOG2R>
NRB_DEV void conv_mid_cta(const ConvMidParams &M, double2 *E, unsigned tile, int tid)
{
    typedef GeoM<LOG2R> G;
    typedef Stage2<G, 0> S0;
    typedef Stage2<G, G::NST - 1> SL;
    const u64 F = 1ull << M.f, REST = (u64)G::N, N = F * REST;
    const u64 tps = F / 2;                                   // tiles per signal (F >= 2)
    const u64 sig = (u64)tile / tps, t = (u64)tile % tps;
    const u64 row0 = t, row1 = t == 0 ? F / 2 : F - t;       // t = 0: the two self-paired rows
    double2 *z = M.data + (i64)sig * M.data_stride;
    double2 v[G::PPT];
    if (M.prefetch_dist > 0) {
        // L2 prefetch of a later tile's two rows (and of its second operand's, when that is per signal): one 128-byte
        // line per thread and row piece
        const u64 pt = (u64)tile + (u64)M.prefetch_dist;
        if (pt < M.count * tps) {
            const u64 psig = pt / tps, ptt = pt % tps;
            const u64 prow[2] = {ptt, ptt == 0 ? F / 2 : F - ptt};
            for (int e = tid * 8; e < 2 * G::N; e += G::NT * 8) {
                const u64 off = prow[e >> LOG2R] * REST + (u64)(e & (G::N - 1));
                NRB_PREFETCH_L2(M.data + (i64)psig * M.data_stride + off);
                if (M.b_stride && M.op != SPEC_AUTOCORREL) NRB_PREFETCH_L2(M.b + (i64)psig * M.b_stride + off);
            }
        }
    }

    // ---- forward REST-point transforms of both rows (reference isign = +1: re / im swapped on the way in)
#pragma unroll
    for (int i = 0; i < S0::BPT; ++i) {
        int ln, jj;
        S0::coords(tid, i, ln, jj);
        const double2 *src = z + (ln == 0 ? row0 : row1) * REST + (u64)jj;
#pragma unroll
        for (int r = 0; r < S0::R; ++r) v[i * S0::R + r] = cswap(NRB_LDS(src + r * S0::NB));
    }
    ChainM<G, 0>::run(M, E, tid, v);
    mid_scatter<G, G::NST - 1>(E, tid, v);                   // natural order; the true bin is cswap(E[..])
    NRB_SYNC();

    // ---- untangle, spectral op, re-tangle on the pair (aux_spectral_zt's rules, one owner per pair)
    {
        const bool self = M.op == SPEC_AUTOCORREL;
        const int op = self ? SPEC_CORREL : M.op;
        const double inv = 1.0 / (double)N;
        const double2 *zb = self ? nullptr : M.b + (i64)sig * M.b_stride;
#pragma unroll 4
        for (int i = 0; i < G::PPT; ++i) {
            const int idx = tid + i * G::NT;
            const int line = idx >> LOG2R;
            const u64 kr = (u64)(idx & (G::N - 1));
            const u64 kf = line == 0 ? row0 : row1;
            if (t != 0 && line == 1) continue;               // the pair is owned by its row-kf element
            if (kf == 0 && kr == 0) {                        // k = 0: DC and Nyquist share the element
                const double2 a0 = cswap(E[G::phys(0, 0)]);
                double2 r0 = self ? a0 : NRB_LDS(zb);
                r0 = make_double2(r0.x + r0.y, r0.x - r0.y);
                const double g0 = spectral_op_real(op, a0.x + a0.y, r0.x, inv);
                const double gn = spectral_op_real(op, a0.x - a0.y, r0.y, inv);
                E[G::phys(0, 0)] = make_double2(0.5 * (g0 + gn), 0.5 * (g0 - gn));
                continue;
            }
            if (kf == 0 && kr == REST / 2) {                 // k = N/2: untangling is the identity
                const double2 am = cswap(E[G::phys(0, (int)kr)]);
                const double2 rm = self ? am : NRB_LDS(zb + kr);
                E[G::phys(0, (int)kr)] = spectral_op(op, am, rm, inv);
                continue;
            }
            int pline;
            u64 pkr;
            if (kf == 0) { if (kr > REST / 2) continue; pline = 0; pkr = REST - kr; }
            else if (2 * kf == F) { if (kr >= REST / 2) continue; pline = line; pkr = REST - 1 - kr; }
            else { pline = 1; pkr = REST - 1 - kr; }
            const u64 pkf = pline == 0 ? row0 : row1;
            const u64 k = kf + F * kr;
            const double2 tw = two_level_tw(M.rtw_lo, M.rtw_hi, M.rtw_h, k);
            const int ia = G::phys(line, (int)kr), im = G::phys(pline, (int)pkr);
            double2 fa, fm, ra, rm;
            untangle_pair<1>(cswap(E[ia]), cswap(E[im]), tw, fa, fm);
            if (self) { ra = fa; rm = fm; }
            else untangle_pair<1>(NRB_LDS(zb + kf * REST + kr), NRB_LDS(zb + pkf * REST + pkr), tw, ra, rm);
            const double2 ga = spectral_op(op, fa, ra, inv), gm = spectral_op(op, fm, rm, inv);
            double2 oa, ob;
            untangle_pair<-1>(ga, gm, tw, oa, ob);
            E[ia] = oa;
            E[im] = ob;
        }
    }
    NRB_SYNC();

    // ---- inverse REST-point transforms (isign = -1: no swap), four-step twiddle on the output, store in place
    mid_gather<G, 0>(E, tid, v);
    NRB_SYNC();                                              // everyone has read before the first exchange overwrites
    ChainM<G, 0>::run(M, E, tid, v);
#pragma unroll
    for (int i = 0; i < SL::BPT; ++i) {
        int ln, jj;
        SL::coords(tid, i, ln, jj);
        const int jm = jj & (SL::NS - 1);
        const int kb = (jj - jm) * SL::R + jm;
        const u64 kf = ln == 0 ? row0 : row1;
        double2 *dst = z + kf * REST + (u64)kb;
        const unsigned q1 = (unsigned)kf;
        const unsigned hm = (1u << M.fs_h) - 1u;
        const unsigned m0 = q1 * (unsigned)kb, m1 = q1 * (unsigned)SL::NS;
        double2 w = cmul(NRB_LDG(M.fs_lo + (m0 & hm)), NRB_LDG(M.fs_hi + (m0 >> M.fs_h)));
        const double2 ws = cmul(NRB_LDG(M.fs_lo + (m1 & hm)), NRB_LDG(M.fs_hi + (m1 >> M.fs_h)));
#pragma unroll
        for (int r = 0; r < SL::R; ++r) {
            NRB_STS(dst + r * SL::NS, cmul(v[i * SL::R + r], w));
            w = cmul(w, ws);
        }
    }
}

} // namespace nrb
