// fft_pass2.cuh -- the "big tile" FFT pass: persistent CTAs, asynchronous prefetch of the next tile into a
// thread-private shared-memory landing zone, and a shared-memory exchange that moves re and im parts separately.
// AN EXPERIMENT, OFF BY DEFAULT (big_row_mask / big_col_mask): measured no faster than fft_pass.cuh, see below.
//
// Idea (profiles/r01_prefetch_row_bench.txt, profiles/r01_tuning.md #34): a tile of 8192 points fills the register
// file of an SM (512 threads x 16 points), so only ONE CTA is resident and nothing overlaps its global loads with
// its shared-memory stages.  Here every thread issues cp.async copies of the 16 elements IT will need for the NEXT
// tile into 16 private slots while the current tile is in its exchange rounds; it later waits for its own copies
// only (no barrier, no mbarrier: nobody else touches those slots).  The landing zone costs TILE x 16 bytes, which is
// paid for by exchanging the re and im parts one after the other through a buffer of half the size (8-byte
// accesses; a per-exchange XOR swizzle keeps the 16-lane phases conflict-free): 128 KiB + 64 KiB for an 8192-point
// tile.  The memory skeleton of this structure runs at 5.33 TB/s against 3.79 TB/s for the plain pass's skeleton.
//
// Result (profiles/r01_tuning.md #35, profiles/r01_big_tile_ab_{1,2,3}.txt): with the math back the kernel does
// 3.53 TB/s on N = 8192 against 3.56 TB/s for fft_pass.cuh, and is slower on every other length: those passes are
// bound by instruction issue and latency at 16 warps per SM, not by exposed load latency, and the split exchange
// doubles the shared-memory instructions.  Kept, with its tests, as the record of that experiment and because its
// stage helpers (Stage2, v2_compute) carry the fused convolution middle (conv_mid.cuh).
//
// Same mathematics, addressing (PassParams) and twiddle tables as fft_pass.cuh: Stockham autosort, radix plan of
// nrb_common.h, first stage fed from the landing zone, last stage stores to global memory.
// Replaces the same reference loops as fft_pass.cuh (FFT_1.rs:8-43 bit reversal + Danielson-Lanczos stages).
#pragma once
#include "fft_pass.cuh"

#if defined(NRB_EMU)
#define NRB_CP_ASYNC16(smem_ptr, gptr) (*(smem_ptr) = *(gptr))
#define NRB_CP_ASYNC_COMMIT() ((void)0)
#define NRB_CP_ASYNC_WAIT_ALL() ((void)0)
#else
#define NRB_CP_ASYNC16(smem_ptr, gptr)                                                                         \
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_ptr)), "l"(gptr) : "memory")
#define NRB_CP_ASYNC_COMMIT() asm volatile("cp.async.commit_group;\n" ::: "memory")
#define NRB_CP_ASYNC_WAIT_ALL() asm volatile("cp.async.wait_group 0;\n" ::: "memory")
#endif

namespace nrb {

// big-tile geometry: which (layout, log2n) have a v2 kernel, and its tile / thread counts
NRB_HD constexpr int v2_tile_log2(int log2n, int layout) { return (layout == LAYOUT_ROW ? (log2n >= 12 ? log2n : 12) : 13); }
// experiment switches (profiles/r01_tuning.md #33): points per thread, and NRB_V2_DIRECT = 1: no landing zone and no
// persistent loop (one CTA per tile, first stage loads global memory as fft_pass.cuh does) -- only the split exchange
#ifndef NRB_V2_PPT
#define NRB_V2_PPT 16
#endif
#ifndef NRB_V2_DIRECT
#define NRB_V2_DIRECT 0
#endif
NRB_HD constexpr int v2_ppt() { return NRB_V2_PPT; }

template <int LOG2N_, int LAYOUT_, int VARIANT_> struct Geo2 {
    static constexpr int LOG2N = LOG2N_, LAYOUT = LAYOUT_, VARIANT = VARIANT_;
    static constexpr int N = 1 << LOG2N;
    static constexpr int TL = v2_tile_log2(LOG2N, LAYOUT);
    static constexpr int TILE = 1 << TL;
    static constexpr int L = TILE / N;
    static constexpr int PPT = v2_ppt();
    static constexpr int NT = TILE / PPT;
    static constexpr int NST = radix_plan(LOG2N).nst;
    static constexpr int LOG2R0 = radix_plan(LOG2N).r[0] == 2 ? 1 : radix_plan(LOG2N).r[0] == 4 ? 2 : 3;
    static constexpr int LP = N;                                          // ROW: line pitch of the exchange buffer (no padding)
    static constexpr int CP = L + ((VARIANT == VAR_XPOSE) ? 1 : 0);       // COL: pitch of one n
    static constexpr int E_ELEMS = LAYOUT == LAYOUT_ROW ? L * LP : N * CP;   // exchange buffer, in doubles
    static constexpr size_t SMEM_BYTES = (NRB_V2_DIRECT ? 0 : (size_t)TILE * 16) + (size_t)E_ELEMS * 8;
    static_assert(N >= PPT, "big-tile pass needs N >= points per thread");
    static_assert(L >= 1, "line longer than the tile");
    // Index (in doubles) of element n of line l in the exchange that follows stage S.  8-byte accesses are served in
    // 16-lane phases, one 128-byte wavefront per phase if the 16 lanes hit 16 different 8-byte bank pairs.
    // ROW: every exchange is written once and read once, so each one has its own XOR swizzle of the low 4 index bits
    // by higher bits (measured: the one-pad-per-8 rule of the 16-byte exchange costs 2x the wavefronts here, the pad
    // itself breaks a 16-lane run).  Reads are 16 aligned consecutive elements (any such swizzle keeps them a
    // permutation of one 16-block); writes of stage S with Ns < 16 come in 16/Ns groups whose bases differ by
    // Ns*R: the swizzle adds a different multiple of Ns to each group.
    // COL: rows of L >= 8 elements, two rows per phase -- the rows j*R0 and (j+1)*R0 written by the first stage would
    // share banks, so the low bit of the row index is flipped by bit log2(R0) (consecutive rows stay a pair).
    template <int S> NRB_DEVM static int phys(int l, int n)
    {
        if (LAYOUT == LAYOUT_ROW) {
            constexpr int NS = stage_ns(LOG2N, S);
            const int f = NS == 1 ? ((n >> 4) & 7) : NS == 2 ? (((n >> 4) & 7) << 1) : NS == 4 ? (((n >> 5) & 3) << 2)
                        : NS == 8 ? (((n >> 6) & 1) << 3) : 0;
            return l * LP + (n ^ f);
        }
        if (VARIANT == VAR_XPOSE) return n * CP + l;
        return ((n ^ ((n >> LOG2R0) & 1)) * L) + l;
    }
};

template <class G, int S> struct Stage2 {
    static constexpr int R = radix_plan(G::LOG2N).r[S];
    static constexpr int NS = stage_ns(G::LOG2N, S);
    static constexpr int NB = G::N / R;
    static constexpr int BPT = G::PPT / R;
    static_assert(BPT >= 1, "radix larger than points per thread");
    NRB_DEVM static void coords(int tid, int i, int &ln, int &jj)
    {
        if (G::LAYOUT == LAYOUT_COL) {
            const int b = tid + i * G::NT;
            ln = b & (G::L - 1);
            jj = b / G::L;
        } else {
            constexpr int TPL = G::N / G::PPT;      // threads per line
            ln = tid / TPL;
            jj = (tid & (TPL - 1)) + i * TPL;
        }
    }
};

// ---- landing zone: slot s of thread tid is S[s * NT + tid]; slot i*R0 + r = element jj_i + r*NB of line ln_i ----
template <class G>
NRB_DEV void v2_prefetch(const PassParams &P, double2 *S, unsigned tile, int tid)
{
    typedef Stage2<G, 0> T;
#pragma unroll
    for (int i = 0; i < T::BPT; ++i) {
        int ln, jj;
        T::coords(tid, i, ln, jj);
        const u64 q = P.q_begin + (u64)tile * G::L + (u64)ln;
        if (q < P.q_end) {
            const double2 *src = P.in + line_base(q, P.in_s0, P.in_s1, P.in_s2, P.logA, P.logB);
#pragma unroll
            for (int r = 0; r < T::R; ++r)
                NRB_CP_ASYNC16(S + (i * T::R + r) * G::NT + tid, src + elem_off(jj + r * T::NB, P.in_es, P.in_eshift, P.in_es_hi));
        }
    }
}

template <class G, int DIR>
NRB_DEV void v2_read_landing(const PassParams &P, const double2 *S, unsigned tile, int tid, double2 *v)
{
    typedef Stage2<G, 0> T;
#pragma unroll
    for (int i = 0; i < T::BPT; ++i) {
        int ln, jj;
        T::coords(tid, i, ln, jj);
        const u64 q = P.q_begin + (u64)tile * G::L + (u64)ln;
        const bool ok = q < P.q_end;
#pragma unroll
        for (int r = 0; r < T::R; ++r) {
            double2 x = make_double2(0.0, 0.0);
            if (ok) x = S[(i * T::R + r) * G::NT + tid];
            v[i * T::R + r] = io_swap<DIR>(x);
        }
    }
}

template <class G, int DIR>
NRB_DEV void v2_load_direct(const PassParams &P, unsigned tile, int tid, double2 *v)
{
    typedef Stage2<G, 0> T;
#pragma unroll
    for (int i = 0; i < T::BPT; ++i) {
        int ln, jj;
        T::coords(tid, i, ln, jj);
        const u64 q = P.q_begin + (u64)tile * G::L + (u64)ln;
        const bool ok = q < P.q_end;
        const double2 *src = P.in + line_base(q, P.in_s0, P.in_s1, P.in_s2, P.logA, P.logB);
#pragma unroll
        for (int r = 0; r < T::R; ++r) {
            double2 x = make_double2(0.0, 0.0);
            if (ok) x = NRB_LDS(src + elem_off(jj + r * T::NB, P.in_es, P.in_eshift, P.in_es_hi));
            v[i * T::R + r] = io_swap<DIR>(x);
        }
    }
}

// ---- twiddle + butterfly of stage S on the thread's registers (same arithmetic as fft_stage) ----
template <class G, int S, class PP>
NRB_DEV void v2_compute(const PP &P, int tid, double2 *v)
{
    typedef Stage2<G, S> T;
    constexpr int R = T::R;
#pragma unroll
    for (int i = 0; i < T::BPT; ++i) {
        if (T::NS > 1) {
            int ln, jj;
            T::coords(tid, i, ln, jj);
            const int jm = jj & (T::NS - 1);
            const double2 *tp = P.tw + stage_tw_off(G::LOG2N, S) + jm * (R - 1);
            if (R >= 4) {
                double2 w[R];
                w[1] = NRB_LDG(tp);
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
            } else {
#pragma unroll
                for (int r = 1; r < R; ++r) v[i * R + r] = cmul(v[i * R + r], NRB_LDG(tp + (r - 1)));
            }
        }
        Bfly<R>::run(v + i * R);
    }
}

// ---- exchange between stage S and stage S+1: re parts, then im parts, through the half-size buffer ----
template <class G, int S>
NRB_DEV void v2_exchange(double *E, int tid, double2 *v)
{
    constexpr int S_ = S;
    typedef Stage2<G, S> A;        // scatter pattern of the stage that produced v
    typedef Stage2<G, S + 1> B;    // gather pattern of the stage that consumes it
    double2 w[G::PPT];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int i = 0; i < A::BPT; ++i) {
            int ln, jj;
            A::coords(tid, i, ln, jj);
            const int jm = jj & (A::NS - 1);
            const int kb = (jj - jm) * A::R + jm;
#pragma unroll
            for (int r = 0; r < A::R; ++r) E[G::template phys<S_>(ln, kb + r * A::NS)] = half == 0 ? v[i * A::R + r].x : v[i * A::R + r].y;
        }
        NRB_SYNC();
#pragma unroll
        for (int i = 0; i < B::BPT; ++i) {
            int ln, jj;
            B::coords(tid, i, ln, jj);
#pragma unroll
            for (int r = 0; r < B::R; ++r) {
                const double x = E[G::template phys<S>(ln, jj + r * B::NB)];
                if (half == 0) w[i * B::R + r].x = x; else w[i * B::R + r].y = x;
            }
        }
        NRB_SYNC();   // everyone has read before anyone overwrites (next half / next exchange / next tile)
    }
#pragma unroll
    for (int k = 0; k < G::PPT; ++k) v[k] = w[k];
}

// ---- last stage -> global memory (PLAIN): natural order, optional four-step twiddle ----
template <class G, int DIR>
NRB_DEV void v2_store_plain(const PassParams &P, unsigned tile, int tid, const double2 *v)
{
    typedef Stage2<G, G::NST - 1> T;
#pragma unroll
    for (int i = 0; i < T::BPT; ++i) {
        int ln, jj;
        T::coords(tid, i, ln, jj);
        const int jm = jj & (T::NS - 1);
        const int kb = (jj - jm) * T::R + jm;
        const u64 q = P.q_begin + (u64)tile * G::L + (u64)ln;
        if (q < P.q_end) {
            double2 *dst = P.out + line_base(q, P.out_s0, P.out_s1, P.out_s2, P.logA, P.logB);
            double2 tw = make_double2(1.0, 0.0), tw_step = make_double2(1.0, 0.0);
            if (P.tw_on) {
                const unsigned q1 = line_q1(P, q);
                tw = fourstep_tw_m(P, q1 * (unsigned)kb);
                tw_step = fourstep_tw_m(P, q1 * (unsigned)T::NS);
            }
#pragma unroll
            for (int r = 0; r < T::R; ++r) {
                const int k = kb + r * T::NS;
                double2 y = v[i * T::R + r];
                if (P.tw_on) { y = cmul(y, tw); tw = cmul(tw, tw_step); }
                NRB_STS(dst + elem_off(k, P.out_es, P.out_eshift, P.out_es_hi), io_swap<DIR>(y));
            }
        }
    }
}

// ---- last stage -> exchange buffer -> line-contiguous store with the four-step twiddle (XPOSE, COL only) ----
template <class G, int DIR>
NRB_DEV void v2_store_xpose(const PassParams &P, double *E, unsigned tile, int tid, double2 *v)
{
    typedef Stage2<G, G::NST - 1> A;
    constexpr int S_ = G::NST - 1;
    constexpr int C = (G::N > G::NT) ? G::N / G::NT : 1;
    constexpr int D = (G::NT >= G::N) ? G::NT / G::N : 1;
    constexpr int EPC = G::PPT / C;
    double2 w[G::PPT];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int i = 0; i < A::BPT; ++i) {
            int ln, jj;
            A::coords(tid, i, ln, jj);
            const int jm = jj & (A::NS - 1);
            const int kb = (jj - jm) * A::R + jm;
#pragma unroll
            for (int r = 0; r < A::R; ++r) E[G::template phys<S_>(ln, kb + r * A::NS)] = half == 0 ? v[i * A::R + r].x : v[i * A::R + r].y;
        }
        NRB_SYNC();
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int idx0 = tid + c * G::NT;
            const int k = idx0 & (G::N - 1), l0 = idx0 >> G::LOG2N;
#pragma unroll
            for (int e = 0; e < EPC; ++e) {
                const double x = E[G::template phys<G::NST - 1>(l0 + e * D, k)];
                if (half == 0) w[c * EPC + e].x = x; else w[c * EPC + e].y = x;
            }
        }
        NRB_SYNC();
    }
    const u64 q_tile = P.q_begin + (u64)tile * G::L;
    const bool geometric = P.tw_on && P.logB == 0 && ((1ull << P.logA) >= (u64)G::L);   // no wrap of q1 inside the tile
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int idx0 = tid + c * G::NT;
        const int k = idx0 & (G::N - 1), l0 = idx0 >> G::LOG2N;
        double2 tw = make_double2(1.0, 0.0), tw_step = make_double2(1.0, 0.0);
        if (geometric) {
            tw = fourstep_tw_m(P, line_q1(P, q_tile + (u64)l0) * (unsigned)k);
            tw_step = fourstep_tw_m(P, (unsigned)(D * k));
        }
#pragma unroll
        for (int e = 0; e < EPC; ++e) {
            const int l = l0 + e * D;
            const u64 q = q_tile + (u64)l;
            if (q < P.q_end) {
                double2 y = w[c * EPC + e];
                if (geometric) y = cmul(y, tw);
                else if (P.tw_on) y = cmul(y, fourstep_tw(P, q, (unsigned)k));
                double2 *dst = P.out + line_base(q, P.out_s0, P.out_s1, P.out_s2, P.logA, P.logB);
                NRB_STS(dst + (i64)k * P.out_es, io_swap<DIR>(y));
            }
            tw = cmul(tw, tw_step);
        }
    }
}

// stages S..NST-1 on registers: compute, exchange, compute, ...
template <class G, int S> struct Chain2 {
    NRB_DEVM static void run(const PassParams &P, double *E, int tid, double2 *v)
    {
        v2_compute<G, S>(P, tid, v);
        if (S + 1 < G::NST) {
            v2_exchange<G, (S + 1 < G::NST ? S : 0)>(E, tid, v);
            Chain2<G, (S + 1 < G::NST ? S + 1 : -1)>::run(P, E, tid, v);
        }
    }
};
template <class G> struct Chain2<G, -1> {
    NRB_DEVM static void run(const PassParams &, double *, int, double2 *) {}
};

// The whole persistent CTA: tiles first, first + stride, ... < ntiles.
template <int LOG2N, int LAYOUT, int DIR, int VARIANT>
NRB_DEV void fft_pass2_cta(const PassParams &P, double2 *sm, unsigned first, unsigned stride, unsigned ntiles, int tid)
{
    typedef Geo2<LOG2N, LAYOUT, VARIANT> G;
    static_assert(G::NST >= 2, "big-tile pass needs at least two stages");
    if (NRB_V2_DIRECT) {
        double *E0 = reinterpret_cast<double *>(sm);
        for (unsigned tile = first; tile < ntiles; tile += stride) {
            double2 v[G::PPT];
            v2_load_direct<G, DIR>(P, tile, tid, v);
            Chain2<G, 0>::run(P, E0, tid, v);
            if (VARIANT == VAR_XPOSE) v2_store_xpose<G, DIR>(P, E0, tile, tid, v);
            else v2_store_plain<G, DIR>(P, tile, tid, v);
        }
        return;
    }
    double2 *S = sm;
    double *E = reinterpret_cast<double *>(sm + G::TILE);
    if (first < ntiles) v2_prefetch<G>(P, S, first, tid);
    NRB_CP_ASYNC_COMMIT();
    for (unsigned tile = first; tile < ntiles; tile += stride) {
        double2 v[G::PPT];
        NRB_CP_ASYNC_WAIT_ALL();
        v2_read_landing<G, DIR>(P, S, tile, tid, v);
        const unsigned long long next = (unsigned long long)tile + stride;
        if (next < ntiles) v2_prefetch<G>(P, S, (unsigned)next, tid);
        NRB_CP_ASYNC_COMMIT();
        Chain2<G, 0>::run(P, E, tid, v);
        if (VARIANT == VAR_XPOSE) v2_store_xpose<G, DIR>(P, E, tile, tid, v);
        else v2_store_plain<G, DIR>(P, tile, tid, v);
    }
}

} // namespace nrb
