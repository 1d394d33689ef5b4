// nrb_common.h -- shared definitions for the numrs_b200 device code and host planner.
//
// The device code in fft_pass.cuh / aux_kernels.cuh is written against a tiny portability
// layer (NRB_DEV, NRB_SYNC, NRB_LDG, double2) so the very same source can also be compiled
// with g++ under -DNRB_EMU, where a CTA is emulated by a pool of host threads.  The emulation
// exists only so that tests/ can validate the index arithmetic of every kernel in a container
// without a GPU; it is never built into libnumrs_b200.so and the product has no CPU path.
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(NRB_EMU)
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
namespace nrb_emu { void barrier(); }
#define NRB_DEV inline
#define NRB_DEVM inline
#define NRB_HD inline
#define NRB_SYNC() nrb_emu::barrier()
#define NRB_SYNCWARP() nrb_emu::barrier()
#define NRB_LDG(p) (*(p))
#define NRB_LDS(p) (*(p))
#define NRB_STS(p, v) (*(p) = (v))
namespace nrb_emu { double shfl(double v, int src_lane); }
#define NRB_SHFL(v, lane) nrb_emu::shfl((v), (lane))
#define NRB_PREFETCH_L2(p) ((void)(p))
#else
#include <cuda_runtime.h>
#define NRB_DEV __device__ __forceinline__
#define NRB_DEVM __device__ __forceinline__
#define NRB_HD __host__ __device__ __forceinline__
#define NRB_SYNC() __syncthreads()
#define NRB_SYNCWARP() __syncwarp()
#define NRB_LDG(p) __ldg(p)
// streaming data: bypass L1 (ld.global.cg) -- every element is touched once per pass, and keeping it
// out of L1 leaves the cache to the twiddle tables (measured: 5.2 -> 6.3 TB/s on the strided pattern)
#ifndef NRB_NO_STREAM_LD
#define NRB_LDS(p) __ldcg(p)
#else
#define NRB_LDS(p) (*(p))
#endif
#define NRB_SHFL(v, lane) __shfl_sync(0xffffffffu, (v), (lane))
#define NRB_PREFETCH_L2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#ifdef NRB_STREAM_ST
#define NRB_STS(p, v) __stcg((p), (v))
#else
#define NRB_STS(p, v) (*(p) = (v))
#endif
#endif

namespace nrb {

typedef long long i64;
typedef unsigned long long u64;

// ---- kernel selection keys ----
enum Layout { LAYOUT_ROW = 0, LAYOUT_COL = 1 };
// PLAIN: first stage loads from global, last stage stores to global (both fused).
// REAL : ROW only. dir=+1: c2c then real untangle (post) ; dir=-1: untangle (pre) then c2c.
// XPOSE: COL only. last stage goes through shared memory and is stored row-like
//        (each line contiguous), with the four-step twiddle applied on the way out.
enum Variant { VAR_PLAIN = 0, VAR_REAL = 1, VAR_XPOSE = 2 };

enum RealMode { REAL_NONE = 0, REAL_PACKED = 1, REAL_SPEQ = 2 };

// geometry of the block-cooperative running sum (aux_scan_cta)
constexpr int kScanThreads = 256;
constexpr int kScanPer = 8;                                  // consecutive positions per thread in phase 2
constexpr int kScanChunk = kScanThreads * kScanPer;          // positions per CTA
constexpr int kScanSmem = kScanChunk + kScanChunk / 8 + kScanThreads;   // double2 elements (padded tile + thread sums)
constexpr int kSlabMaxChunks = 16;  // z-chunks of the pipelined slab exchange (flag slots per receive buffer)
constexpr int kMaxLog2N = 13;       // longest line one CTA transforms in shared memory
// points each thread owns per stage (compile-time): 16 -> TILE/16 threads and two radix-8 butterflies
// per thread (128 registers, 2 CTAs/SM), 8 -> TILE/8 threads and one (64 registers, 2 CTAs/SM, twice
// the warps).  Measured on B200: 8 wins for COL kernels and short ROW lines, 16 for ROW lines >= 2048.
#ifndef NRB_PPT_ROW_SMALL
#define NRB_PPT_ROW_SMALL 8
#endif
#ifndef NRB_PPT_ROW_LARGE
#define NRB_PPT_ROW_LARGE 16
#endif
#ifndef NRB_PPT_ROW_SPLIT
#define NRB_PPT_ROW_SPLIT 10     /* log2n <= this uses NRB_PPT_ROW_SMALL */
#endif
#ifndef NRB_PPT_COL
#define NRB_PPT_COL 8
#endif
// COL line lengths (bit log2n) that run 16 points/thread instead (256 threads, 128 registers, 2 CTAs/SM): measured
// +2.5 % at N = 512 and +2 % at N = 64, -3 % at N = 128 and N = 1024 (profiles/r01_tuning.md #23)
#ifndef NRB_PPT_COL16_MASK
#define NRB_PPT_COL16_MASK ((1 << 9) | (1 << 6))
#endif
NRB_HD constexpr int points_per_thread(int layout, int log2n)
{
    return layout == 0 /* LAYOUT_ROW */ ? (log2n <= NRB_PPT_ROW_SPLIT ? NRB_PPT_ROW_SMALL : NRB_PPT_ROW_LARGE)
                                        : (((NRB_PPT_COL16_MASK >> log2n) & 1) ? 16 : NRB_PPT_COL);
}

// line lengths (bit log2n) whose PLAIN pass is also built with the cheap addressing path (fft_stage, SIMPLE) and takes
// it whenever the pass's element index is not split.  Measured on B200 (profiles/r01_simple_addr_ab.txt): +6 % on
// contiguous lines of 8192, +2.4 ... +6.6 % on strided lines of 512, +3 % on 1024; neutral on contiguous 4096;
// -3 % on strided lines of 64 / 128 (so those are left out).
#ifndef NRB_SIMPLE_ROW_MASK
#define NRB_SIMPLE_ROW_MASK (1 << 13)
#endif
#ifndef NRB_SIMPLE_COL_MASK
#define NRB_SIMPLE_COL_MASK ((1 << 9) | (1 << 10))
#endif
// the transposing (XPOSE) passes: cheap addressing for the strided loads and the line-contiguous store; measured on the
// first pass of a 2^20 transform: fft_col_xpose_n1024 3 608 -> 3 839 GB/s (profiles/r02_tuning.md #40)
#ifndef NRB_SIMPLE_XPOSE_MASK
#define NRB_SIMPLE_XPOSE_MASK (1 << 10)
#endif
NRB_HD constexpr bool simple_built(int log2n, int layout, int variant = 0 /* VAR_PLAIN */)
{
    return variant == 2 /* VAR_XPOSE */ ? ((NRB_SIMPLE_XPOSE_MASK >> log2n) & 1) != 0
         : variant == 0 ? (((layout == 0 /* LAYOUT_ROW */ ? NRB_SIMPLE_ROW_MASK : NRB_SIMPLE_COL_MASK) >> log2n) & 1) != 0
                        : false;
}

// ---- radix plan per log2(N): stage radices, first stage first ----
constexpr int kMaxStages = 5;
struct RadixPlan { int nst; int r[kMaxStages]; };

#ifndef NRB_RADIX16
#define NRB_RADIX16 0
#endif

NRB_HD constexpr RadixPlan radix_plan(int log2n)
{
#if NRB_RADIX16
    return log2n == 1 ? RadixPlan{1, {2, 1, 1, 1, 1}}
         : log2n == 2 ? RadixPlan{1, {4, 1, 1, 1, 1}}
         : log2n == 3 ? RadixPlan{1, {8, 1, 1, 1, 1}}
         : log2n == 4 ? RadixPlan{1, {16, 1, 1, 1, 1}}
         : log2n == 5 ? RadixPlan{2, {4, 8, 1, 1, 1}}
         : log2n == 6 ? RadixPlan{2, {8, 8, 1, 1, 1}}
         : log2n == 7 ? RadixPlan{2, {8, 16, 1, 1, 1}}
         : log2n == 8 ? RadixPlan{2, {16, 16, 1, 1, 1}}
         : log2n == 9 ? RadixPlan{3, {8, 8, 8, 1, 1}}
         : log2n == 10 ? RadixPlan{3, {8, 8, 16, 1, 1}}
         : log2n == 11 ? RadixPlan{3, {8, 16, 16, 1, 1}}
         : log2n == 12 ? RadixPlan{3, {16, 16, 16, 1, 1}}
         : RadixPlan{4, {8, 8, 8, 16, 1}};
#else
    return log2n == 1 ? RadixPlan{1, {2, 1, 1, 1, 1}}
         : log2n == 2 ? RadixPlan{1, {4, 1, 1, 1, 1}}
         : log2n == 3 ? RadixPlan{1, {8, 1, 1, 1, 1}}
         : log2n == 4 ? RadixPlan{2, {2, 8, 1, 1, 1}}
         : log2n == 5 ? RadixPlan{2, {4, 8, 1, 1, 1}}
         : log2n == 6 ? RadixPlan{2, {8, 8, 1, 1, 1}}
         : log2n == 7 ? RadixPlan{3, {2, 8, 8, 1, 1}}
         : log2n == 8 ? RadixPlan{3, {4, 8, 8, 1, 1}}
         : log2n == 9 ? RadixPlan{3, {8, 8, 8, 1, 1}}
         : log2n == 10 ? RadixPlan{4, {2, 8, 8, 8, 1}}
         : log2n == 11 ? RadixPlan{4, {4, 8, 8, 8, 1}}
         : log2n == 12 ? RadixPlan{4, {8, 8, 8, 8, 1}}
         : RadixPlan{5, {2, 8, 8, 8, 8}};
#endif
}

// Ns of stage s = product of the radices of the stages before it
NRB_HD constexpr int stage_ns(int log2n, int s)
{
    int ns = 1;
    for (int t = 0; t < s; ++t) ns *= radix_plan(log2n).r[t];
    return ns;
}
// offset (in double2) of stage s inside the packed per-size stage-twiddle table:
// stage t >= 1 stores Ns(t) * (R(t)-1) entries, entry [jm*(R-1) + (r-1)] = exp(-2 pi i jm r / (Ns R))
NRB_HD constexpr int stage_tw_off(int log2n, int s)
{
    int off = 0;
    for (int t = 1; t < s; ++t) off += stage_ns(log2n, t) * (radix_plan(log2n).r[t] - 1);
    return off;
}
NRB_HD constexpr int stage_tw_total(int log2n)
{
    return stage_tw_off(log2n, radix_plan(log2n).nst);
}

// tile geometry: a CTA transforms L = TILE/N lines of N points
// ROW lines are contiguous, so a ROW tile may be smaller than a COL tile (whose line count sets the
// length of every global access run): smaller CTAs, more of them per SM, cheaper barriers.
#ifndef NRB_TL_ROW
#define NRB_TL_ROW 9
#endif
// COL line lengths (bit log2n) that take 8192-point tiles (twice the lines per tile: global runs of 2 L x 16 bytes, one CTA
// per SM) instead of 4096-point ones
#ifndef NRB_TL_COL13_MASK
#define NRB_TL_COL13_MASK 0
#endif
NRB_HD constexpr int tile_log2(int log2n, int layout)
{
    return layout == 0 /* ROW */ ? (log2n > NRB_TL_ROW ? log2n : NRB_TL_ROW)
                                 : (log2n > 12 ? log2n : (((NRB_TL_COL13_MASK >> log2n) & 1) ? 13 : 12));
}
NRB_HD constexpr int cta_threads(int log2n, int layout) { return (1 << tile_log2(log2n, layout)) / points_per_thread(layout, log2n); }

// shared-memory footprint in double2 elements
NRB_HD constexpr int row_line_pitch(int log2n) { return (1 << log2n) + ((1 << log2n) >> 3); }
NRB_HD constexpr int lines_per_tile(int log2n, int layout) { return (1 << tile_log2(log2n, layout)) >> log2n; }
NRB_HD constexpr int col_line_count(int log2n) { return lines_per_tile(log2n, 1); }
NRB_HD constexpr int col_pitch(int log2n, int variant)
{
    // XPOSE reads the tile back line-contiguously: pitch L+1 keeps that conflict-free for L >= 8;
    // for L = 2, 4 an XOR swizzle is used instead (Geo::phys), no padding
    return col_line_count(log2n) + ((variant == VAR_XPOSE && col_line_count(log2n) >= 8) ? 1 : 0);
}
NRB_HD constexpr size_t smem_elems(int log2n, int layout, int variant)
{
    return layout == LAYOUT_ROW
               ? (size_t)lines_per_tile(log2n, 0) * (size_t)row_line_pitch(log2n)
               : (size_t)(1 << log2n) * (size_t)col_pitch(log2n, variant);
}

// ---- parameters of one FFT pass (one kernel launch) ----
// A pass transforms lines q in [q_begin, q_end).  Line q decomposes as
//   q2 = q & (2^logB - 1), q1 = (q >> logB) & (2^logA - 1), q0 = q >> (logA + logB)
// and its element n lives at  base + q0*s0 + q1*s1 + q2*s2 + n*es  (units: complex elements).
struct PassParams {
    const double2 *in;
    double2 *out;
    const double2 *tw;      // packed stage twiddles of this log2n
    const double2 *tw_lo;   // four-step twiddle, exp(-2 pi i m / M): low / high tables
    const double2 *tw_hi;
    const double2 *rtw;     // VAR_REAL: exp(-i pi k / N), k < N
    double2 *speq;          // VAR_REAL + REAL_SPEQ: Nyquist plane, element q
    i64 in_s0, in_s1, in_s2, in_es;
    i64 out_s0, out_s1, out_s2, out_es;
    // two-level element index (slab exchange layouts): element n sits at
    //   (n & (2^eshift - 1)) * es + (n >> eshift) * es_hi ; eshift = 31 disables the split
    i64 in_es_hi, out_es_hi;
    int in_eshift, out_eshift;
    // fused exchange (slab transforms over NVLink): when out_peer_on, the high part of the output element
    // index selects an exchange block through a pointer table instead of a stride:
    //   address = out_peer[8 * zsel + (n >> out_eshift)] + out_peer_off + line offset + (n & mask) * out_es
    // and the same on the input side (in_peer_on / in_peer).  A table entry is either local memory or a PEER GPU's buffer
    // (mapped through CUDA IPC or peer access).  zsel = ((q & peer_zmask) >= peer_zthr) picks the second half of the
    // tables for the lines whose z index is past a threshold: the "push + pull" split of the exchange -- stage 0 PUSHES the
    // low-z part of every block into the consumer's receive buffer and writes the high-z part into its own send buffer,
    // from which the consumer's stage 1 PULLS it, so the link carries half the bytes during each of the two passes instead
    // of all of them during the first (plan.cpp exec_slab_stage).
    double2 *out_peer[16];
    const double2 *in_peer[16];
    i64 out_peer_off;
    int out_peer_on, in_peer_on;
    unsigned peer_zmask, peer_zthr;
    // tile subset (slab stages with the z pass cut into y-chunks, plan.cpp): the launch covers only the tiles whose index
    // has the bits [tile_run, tile_run + tile_nsel) equal to tile_sel; CTA b works on tile
    //   (((b >> tile_run) << tile_nsel) | tile_sel) << tile_run | (b & (2^tile_run - 1)).   tile_nsel = 0: all tiles.
    int tile_run, tile_nsel, tile_sel;
    int prefetch_dist;      // > 0: every CTA first asks L2 for the input of tile (own + prefetch_dist), so DRAM keeps
                            // streaming while the CTAs of an SM are in their shared-memory stages
    int grid_cap;           // > 0: launch at most this many CTAs (they loop over the tiles); used to leave SM slots
                            // to the local pass that runs beside an NVLink-bound exchange pass
    u64 q_begin, q_end;
    int logA, logB;
    int tw_on;              // multiply output k of line q by exp(-/+ 2 pi i q1 k / M)
    int tw_h;               // m = (hi << tw_h) | lo
    int real_mode;
    int simple;             // set by the backend (pass_is_simple): no split element index, ROW element stride 1 -> the kernel
                            // takes the cheap addressing path (fft_stage, SIMPLE)
};

// ---- two dependent passes fused into one persistent launch (rlft3 z + y through L2) ----
// Work is cut into `units` (x-planes).  Pass A's tiles of unit u must all finish before any tile
// of pass B on unit u starts.  CTAs pull tickets in order; the ticket sequence is
//   A(0) .. A(lag-1), then A(i) B(i-lag) for i = lag .. units-1, then B(units-lag) .. B(units-1)
// so B runs `lag` units behind A: its inputs were produced recently (still in L2) and long enough
// ago that the wait is normally already satisfied.  Earlier tickets never wait on later ones, so
// the scheme cannot deadlock whatever the number of resident CTAs.
struct FuseSched {
    unsigned long long *ticket;   // zeroed before every launch
    unsigned *done;               // [units] completed pass-A tiles, zeroed before every launch
    unsigned units, ta, tb, lag;  // ta / tb = tiles of pass A / B per unit
};

// ---- elementwise kernels ----
enum AuxKind {
    AUX_UNTANGLE = 0,     // standalone real untangle (large lines / N == 1)
    AUX_SPECTRAL = 1,     // convlv multiply / divide, correl conj-multiply on packed spectra
    AUX_PAD_RESPONSE = 2, // Convolve.rs:41-63 response placement
    AUX_CORREL_DIRECT = 3,// Correlation.rs:37-50, n <= 32
    AUX_FILL = 4,         // synthetic input generator: n doubles, seed in m, offset in count
    AUX_SPECTRAL_Z = 5,   // untangle + spectral op + inverse untangle in one pass on raw c2c outputs
    AUX_SIGNAL = 6,       // slab exchange barrier: publish `epoch` (in m) to slot `rank` (in n) of every peer's flag array
    AUX_WAIT = 7,         // slab exchange barrier: spin until the first `count` local flags are >= epoch (in m)
    // ---- "next" rows of SURVEY.md 8f (callers on either side of the hot path) ----
    AUX_REDUCE = 8,       // strided partial sums: m accumulators per signal (see aux_reduce)
    AUX_STATS_FINAL = 9,  // partial sums -> (mean, std) per signal (Correlation.rs:200-212 / :236-244)
    AUX_NORMALIZE = 10,   // (x - mean) / std (Correlation.rs:219-220 / :265-266)
    AUX_POWER = 11,       // |z|^2 or |z| per complex point (FFT_1.rs:206-228)
    AUX_PACK2 = 12,       // twofft: two real signals -> one complex signal (FFT_2.rs:33-37)
    AUX_TWOFFT_SPLIT = 13,// twofft: separate the two spectra (FFT_2.rs:53-90, mirror n-k)
    AUX_SCALE = 14,       // out[i] *= 1/n (correl_normalized_fast direct branch, Correlation.rs:252)
    AUX_COSFT = 15,       // cosft1 / cosft2 / sinft pre- and post-processing around realft (Cos_FT.rs, Cos_FT2.rs)
    AUX_SCAN = 16,        // running sums of the odd (cosft1) / even-from-the-top (cosft2) outputs, three-phase
    AUX_CMUL = 17,        // a[i] = a[i] * b[i] (op 0) or a[i] * conj(b[i]) (op 1), times a real scale (bits in m)
    AUX_SPECTRAL_ZT = 18, // AUX_SPECTRAL_Z on spectra kept in the transposed order of a two-pass transform (see aux_spectral_zt)
    AUX_KIND_COUNT = 19
};
enum SpectralOp { SPEC_CONV_MUL = 0, SPEC_CONV_DIV = 1, SPEC_CORREL = 2, SPEC_AUTOCORREL = 3 };
enum ReduceMode { RED_SUM_SQ = 0, RED_CENTERED_SQ = 1, RED_PARTIALS = 2 };
enum StatsMode { STATS_FAST = 0, STATS_MEAN = 1, STATS_STD = 2, STATS_SUM = 3 };
// AUX_COSFT / AUX_SCAN transform selector (AuxParams::dir)
enum CosMode { COS1 = 0, COS2F = 1, SINFT = 2, COS2I_PRE = 3, COS2I_POST = 4 };

struct AuxParams {
    int kind;
    int op;               // SpectralOp / pad_mode / real_mode
    int dir;              // untangle: +1 forward (post), -1 inverse (pre)
    const double2 *a;     // primary input
    const double2 *b;     // second operand
    double2 *out;
    double2 *speq;
    const double2 *rtw_lo; // untangle twiddle exp(-i pi k / N) two-level tables
    const double2 *rtw_hi;
    int rtw_h;
    u64 n;                // untangle: complex line length N ; spectral: real length n ; pad: n
    u64 m;                // pad: taps
    u64 count;            // lines / signals
    i64 a_stride, b_stride, out_stride; // per line / signal, in complex elements (doubles for pad/direct)
    unsigned long long *peer_flags[8];  // AUX_SIGNAL: every rank's flag array (IPC-mapped); AUX_WAIT: [0] = local
};

// ---- fused middle of the long-line convlv / correl pipeline (conv_mid.cuh) ----
struct ConvMidParams {
    double2 *data;               // [count][F][REST] the strided pass's output (twiddled), transformed in place
    const double2 *b;            // second operand: finished transposed spectrum (unused for SPEC_AUTOCORREL)
    i64 data_stride, b_stride;   // per signal, in complex elements; b_stride = 0: one spectrum for the whole batch
    u64 count;
    int f;                       // log2 F
    int op;                      // SpectralOp
    const double2 *tw;           // stage twiddles of the REST-point transform
    const double2 *fs_lo, *fs_hi;   // four-step twiddle exp(-2 pi i m / N), two-level
    int fs_h;
    const double2 *rtw_lo, *rtw_hi; // untangle twiddle exp(-i pi k / N), two-level
    int rtw_h;
    int prefetch_dist;           // > 0: every CTA first asks L2 for the rows of tile (own + prefetch_dist): one CTA per SM has
                                 // nothing else to overlap its serial load / transform / store phases with
};

// ---- cosft1 / cosft2 / sinft and twofft in one kernel for lines that fit on chip (trig_fused.cuh) ----
// complex points per line: 8 .. 8192 (real lines of 16 .. 16384 points; twofft lines of 8 .. 8192 points)
constexpr int kTrigMinLog2 = 3, kTrigMaxLog2 = 13;
struct TrigParams {
    double *io;                      // lines of the reference's 1-based arrays (element 0 of a line unused), transformed in place
    i64 ld;                          // doubles per line
    u64 count;                       // lines
    int mode;                        // COS1, COS2F, SINFT (forward) or COS2I_PRE (= the whole inverse cosft2)
    const double2 *tw;               // stage twiddles of the N = n/2 point transform
    const double2 *rtw;              // exp(-i pi k / N), k < N (realft untangling)
    const double2 *ctw_lo, *ctw_hi;  // exp(-2 pi i m / M), two-level: M = 2n (cosft1, sinft) or 4n (cosft2)
    int ctw_h;
};

struct TwoFFTParams {
    const double *d1, *d2;           // [count][n] real signals
    double2 *f1, *f2;                // [count][n + 1] complex spectra (element n: the reference's two extra doubles, set to 0)
    u64 count;
    const double2 *tw;               // stage twiddles of the n-point transform
};

} // namespace nrb
