// plan.h -- host-side planner: turns a transform request into a list of kernel launches.
// Backend-agnostic (the CUDA backend lives in kernels.cu; tests/emu provides a host-thread
// emulation of the same kernels for index-math validation only).
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "nrb_common.h"

namespace nrb {

// ---- backend interface ----
struct KernelKey { int log2n, layout, dir, variant; };
int be_launch_pass(const KernelKey &key, const PassParams &p, u64 ntiles, void *stream);
int be_launch_aux(const AuxParams &a, void *stream);
// two dependent passes in one persistent launch; be_fused_available tells the planner whether the
// backend has that kernel pair built
bool be_fused_available(const KernelKey &a, const KernelKey &b);
int be_launch_fused(const KernelKey &ka, const PassParams &pa, const KernelKey &kb, const PassParams &pb, const FuseSched &fs,
                    void *stream);
// fused middle of the long-line convlv / correl pipeline (rows of 2^log2rest points); available for the built lengths
bool be_conv_mid_available(int log2rest);
int be_launch_conv_mid(int log2rest, const ConvMidParams &m, u64 ntiles, void *stream);
// cosft1 / cosft2 / sinft and twofft in one kernel (trig_fused.cuh): lines of 2^log2n complex points, kTrigMinLog2 <= log2n <= kTrigMaxLog2
bool be_trig_available(int log2n);
int be_launch_trig(int log2n, const TrigParams &t, void *stream);
int be_launch_twofft(int log2n, const TwoFFTParams &t, void *stream);
int be_malloc(void **p, size_t bytes);
int be_free(void *p);
int be_memset(void *p, int value, size_t bytes, void *stream);
int be_ipc_export(void *dptr, unsigned char handle[64]);
int be_ipc_import(const unsigned char handle[64], void **dptr);
int be_ipc_release(void *dptr);
int be_h2d(void *dst, const void *src, size_t bytes, void *stream);
int be_d2h(void *dst, const void *src, size_t bytes, void *stream);
int be_d2d(void *dst, const void *src, size_t bytes, void *stream);
// strided copies (rows of `width` bytes): host [rows][spitch] -> device [rows][dpitch] and back
int be_h2d_2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows, void *stream);
int be_d2h_2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows, void *stream);
// lets kernels on `dev` (made current by the call) store into `peer`'s memory; 0 also when it was enabled before
int be_enable_peer(int dev, int peer);
int be_sync(void *stream);
int be_current_device();
int be_device_count();
int be_set_device(int dev);
int be_stream_create(void **stream);
int be_stream_destroy(void *stream);
void *be_host_alloc(size_t bytes);
void be_host_free(void *p);
void *be_event_record(void *stream);
void *be_event_create();                          // reusable, no timing
int be_event_record_on(void *event, void *stream);
int be_stream_wait(void *stream, void *event);
int be_stream_create_prio(void **stream, int high_priority);
float be_event_elapsed_ms(void *a, void *b);
void be_event_destroy(void *e);
const char *be_last_error();

void set_error(const std::string &msg);   // thread-local message behind nrb_last_error()
const std::string &get_error();

// ---- buffers a program can refer to ----
enum BufId { BUF_IO = 0, BUF_AUX = 1, BUF_OUT = 2, BUF_WS = 3, BUF_NONE = 4 };
struct BufRef {
    int id;
    i64 off;   // in complex (double2) elements
    BufRef() : id(BUF_NONE), off(0) {}
    BufRef(int i, i64 o) : id(i), off(o) {}
    BufRef operator+(i64 d) const { return BufRef(id, off + d); }
    bool same(const BufRef &o) const { return id == o.id && off == o.off; }
};

struct Step {
    bool is_aux;
    KernelKey key;
    PassParams pp;
    AuxParams ap;
    u64 ntiles;
    BufRef in, out, speq, b;
    bool patch_pad_mode;   // AUX_PAD_RESPONSE: op comes from exec's `arg`
    bool is_mid;           // fused conv middle: mp, in = data (in place), b = second operand; key.log2n = log2 REST
    ConvMidParams mp;
    int trig;              // 1: one-kernel cosft1 / cosft2 / sinft (tp, in = io); 2: one-kernel twofft (fp, in = data1, b = data2,
                           // out = fft1, speq = fft2); key.log2n = log2 of the complex points per line
    TrigParams tp;
    TwoFFTParams fp;
    bool join_side;        // the side lane must have finished before this (lane 0) step starts
    bool side_after_main;  // this lane-1 step must not start before what the caller's stream holds so far
    bool zsplit;           // slab exchange data pass: its lines' low bits are the z index, so the push + pull split applies
    int lane;              // 0 = the caller's stream; 1 = the plan's side stream (small independent work, see SideLane)
    // fused pair: (key, pp, in/out/speq) is pass A, the *2 members are pass B
    bool is_fused;
    KernelKey key2;
    PassParams pp2;
    BufRef in2, out2, speq2;
    FuseSched fs;          // counters live in the plan's scheduler scratch (sched_off = element offset)
    size_t sched_off;
    Step() : is_aux(false), key{0, 0, 0, 0}, pp(), ap(), ntiles(0), patch_pad_mode(false), is_mid(false), mp(), trig(0), tp(), fp(), join_side(false), side_after_main(false), zsplit(false), lane(0), is_fused(false),
             key2{0, 0, 0, 0}, pp2(), fs{nullptr, nullptr, 0, 0, 0, 0}, sched_off(0) {}
};

struct Program {
    std::vector<Step> steps;
};

// device-resident twiddle tables, cached per device and shared between plans.  A table is reference-counted: the
// cache holds one reference (dropped by release_tables(), i.e. nrb_shutdown) and every plan built while a TableScope is
// open holds one for each table its programs point into, so a plan that outlives nrb_shutdown keeps its tables alive.
typedef std::shared_ptr<void> TableRef;
struct TableScope {
    explicit TableScope(std::vector<TableRef> *sink);
    ~TableScope();
    std::vector<TableRef> *prev;
};
struct FourStepTable { const double2 *lo, *hi; int h; };
const double2 *stage_twiddles(int log2n);          // packed per-stage tables (nrb_common.h layout)
FourStepTable fourstep_table(int log2m);           // exp(-2 pi i m / 2^log2m), two-level
const double2 *real_twiddles(int log2n);           // exp(-i pi k / 2^log2n), k < N
void release_tables();

// tunables (environment overrides, read once)
struct Tunables {
    int col_max_log2;      // longest strided-axis FFT done in one pass (NRB_COL_MAX_LOG2, default 10)
    int row_max_log2;      // longest contiguous FFT done in one pass   (NRB_ROW_MAX_LOG2, default 13)
    u64 l2_group_bytes;    // rlft3: bytes of x-planes handled per z/y launch pair (NRB_L2_GROUP_MB; default: whole volume)
    int fuse_zy;           // rlft3: fuse the z and y passes of each x-plane through L2 (NRB_FUSE_ZY, default 0: measured slower)
    int fuse_lag;          // planes pass B runs behind pass A (NRB_FUSE_LAG, default 16)
    u64 batch_group_bytes; // convlv/correl: bytes of signals handled per launch group (NRB_BATCH_GROUP_MB, default 512)
    int conv_transposed;   // convlv/correl with lines longer than a tile: two passes per transform and the spectrum in
                           // transposed order instead of three natural-order passes (NRB_CONV_TRANSPOSED, default 1)
    int prefetch_dist;     // tiles ahead whose input every CTA prefetches into L2 (NRB_PREFETCH_DIST; 0 = off, -1 = per-kernel policy, default)
    int conv_fused_mid;    // long-line convlv / correl: contiguous forward pass + spectral step + contiguous inverse pass in one kernel
                           // (NRB_CONV_FUSED_MID, default 1: convlv -3 %, correl -9.5 %, autocorrel_fast -5 % at n = 2^22, profiles/r02_tuning.md #39)
    int mid_prefetch;      // fused conv middle: tiles ahead whose rows every CTA prefetches into L2 (NRB_MID_PREFETCH, default 0)
    int conv_rest_log2;    // long-line convlv / correl: log2 of the contiguous rows of the two-pass split (NRB_CONV_REST_LOG2, default 12:
                           // 4096-point rows; 11 halves the fused middle kernel's CTA so that two fit an SM, at the price of a
                           // 1024-point strided pass)
    int speq_side;         // rlft3: run the speq-plane passes on the plan's side stream (NRB_SPEQ_SIDE, default 1: -0.5 % of the 1-GPU step, profiles/r02_tuning.md #38)
    int simple_addr;       // 1: passes whose element index is not split use the cheap addressing path (NRB_SIMPLE_ADDR, default 1)
    int big_row_mask;      // bit log2n set: contiguous lines of 2^log2n points use the big-tile pass of fft_pass2.cuh (NRB_BIG_ROW_MASK)
    int big_col_mask;      // the same for strided lines (NRB_BIG_COL_MASK)
    int xchg_grid_cap;     // pipelined slab exchange: CTAs of an exchange (peer-store) pass, 0 = one per tile (NRB_XCHG_GRID_CAP)
    int num_devices;       // GPUs the host-slice entry points spread one call over (NRB_NUM_DEVICES; 1 = the calling thread's
                           // device only (default), 0 = every visible device, n = devices 0 .. n-1): multi.cpp
    int z_chunks;          // slab stages: y-chunks the z pass and the exchange pass beside it are cut into, so that the z pass of
                           // chunk c + 1 (c - 1) runs on the side stream under the NVLink-bound pass of chunk c
                           // (NRB_Z_CHUNKS, 1 = off, default; 2 or 4)
    int pull_eighths;      // push + pull exchange: eighths of the z range that stage 1 pulls instead of stage 0 pushing them
                           // (NRB_PULL_EIGHTHS, 0 .. 8, default 4: half and half)
    int dma_streams;       // DMA slab exchange: copy streams the pieces of a chunk are spread over (NRB_DMA_STREAMS, 1 .. 4, default 1)
    int tma_col_mask;      // bit log2n set: eligible strided PLAIN passes of 2^log2n points use the TMA-fed kernel of fft_tma.cuh
                           // (NRB_TMA_COL_MASK, default 512 | 1024: +2.3 % / +5.7 % on the y / x pass of rlft3 512^3, +6 % on the second
                           // pass of a 2^20 transform, profiles/r02_tuning.md #47); tma_persist = 1: persistent
                           // CTAs with two tile buffers (the next tile's bulk load in flight during the stages)
    int tma_persist;
    int tma_in_mask;       // bit log2n set: strided PLAIN passes of 2^log2n points whose input side is regular load their tile by TMA and
                           // store from registers (NRB_TMA_IN_MASK); takes precedence over tma_col_mask
    int tma_in_ctas;       // resident CTAs per SM the input-only TMA pass of 512-point lines is compiled for: 2 (no spills) or 3 (80 registers)
    int tma_xpose;         // the transposing 1024-point pass loads its tile by TMA (64-byte swizzle) (NRB_TMA_XPOSE, default 1: +4.3 % on the first pass of a 2^20 transform, profiles/r02_tuning.md #56)
    int pipeline_batches;  // host-slice batch calls run in chunks over three streams (H2D | transforms | D2H overlap): 1 (default) / 0
                           // (NRB_PIPELINE_BATCHES)
    int pipeline_min_kb;   // smallest chunk of a pipelined batch call (NRB_PIPELINE_MIN_KB, default 16 MiB; calls under 4 chunks stay one shot)
    int shard_min_kb;      // batches smaller than this stay on one device (NRB_SHARD_MIN_KB, default 16 MiB)
    int trig_fused;        // cosft1 / cosft2 / sinft of 16 .. 16384 points and twofft of 8 .. 8192 points per line: the whole routine in
                           // one kernel, one HBM pass (trig_fused.cuh) instead of 5-7 launches (NRB_TRIG_FUSED, default 1)
};
const Tunables &tunables();
int set_tunable(const char *name, long value);   // returns 0 if the name is known
// can the pass use the cheap addressing path of fft_stage (element offset = n * es, es = 1 for contiguous lines)?
inline bool pass_is_simple(const KernelKey &key, const PassParams &p)
{
    if (!tunables().simple_addr || p.out_peer_on || p.in_peer_on || !simple_built(key.log2n, key.layout, key.variant)) return false;
    if (p.in_eshift <= kMaxLog2N || p.out_eshift <= kMaxLog2N) return false;   // the element index is split
    if (key.layout == LAYOUT_ROW) return p.in_es == 1 && p.out_es == 1;
    return true;
}
// does this launch take the TMA-fed strided kernel of fft_tma.cuh (option tma_col_mask)?  Pure function of the launch
// geometry (the backend additionally needs 16-byte aligned base pointers and the driver's tensor-map encoder): LAYOUT_COL,
// VAR_PLAIN, no four-step twiddle, no split element index or exchange tables, lines = [outer][inner] with inner a multiple
// of the tile's line count, the same geometry on both sides.
// the transposing 1024-point pass with TMA loads (fft_xpose_tma_kernel): tma_xpose option, geometry of emit_axis's first
// multi-step pass (lines = [outer][rest] with the rest index contiguous, element stride rest)
inline bool pass_takes_tma_xpose(const KernelKey &key, const PassParams &p)
{
    if (!tunables().tma_xpose || key.layout != LAYOUT_COL || key.variant != VAR_XPOSE || key.log2n != 10) return false;
    if (p.out_peer_on || p.in_peer_on || p.grid_cap > 0 || p.tile_nsel > 0 || p.in_eshift <= kMaxLog2N) return false;
    if (p.logB != 0 || p.logA < 2 || p.logA > 28 || p.in_s1 != 1) return false;
    const u64 rest = 1ull << p.logA, L = 4;
    if (p.in_es != (i64)rest || p.in_s0 != (i64)(rest << 10)) return false;
    return (p.q_begin % L) == 0 && ((p.q_end - p.q_begin) % L) == 0 && p.q_end > p.q_begin;
}
// TMA on the input side only (fft_col_tma_in_kernel, option tma_in_mask): any strided PLAIN pass whose INPUT is the regular
// [outer][N][inner] geometry -- expressed either through (logB, in_s2) as emit_axis does or through (logA, in_s1) as the
// conv passes do; *log2_inner tells which.  The output side is unconstrained.
inline bool pass_takes_tma_in(const KernelKey &key, const PassParams &p, int *log2_inner = nullptr)
{
    if (!((tunables().tma_in_mask >> key.log2n) & 1)) return false;
    if (key.layout != LAYOUT_COL || key.variant != VAR_PLAIN || p.in_peer_on || p.grid_cap > 0 || p.tile_nsel > 0) return false;
    if (key.log2n < 7 || key.log2n > 10 || p.in_eshift <= kMaxLog2N) return false;
    int li = -1;
    if (p.logA == 0 && p.in_s2 == 1 && p.logB <= 28) li = p.logB;
    else if (p.logB == 0 && p.in_s1 == 1 && p.logA <= 28) li = p.logA;
    if (li < 0) return false;
    const u64 inner = 1ull << li, N = 1ull << key.log2n, L = (u64)lines_per_tile(key.log2n, LAYOUT_COL);
    if (inner < L || (inner % L) != 0 || p.in_es != (i64)inner || p.in_s0 != (i64)(N * inner)) return false;
    if ((p.q_begin % L) != 0 || ((p.q_end - p.q_begin) % L) != 0 || p.q_end <= p.q_begin) return false;
    if (log2_inner) *log2_inner = li;
    return true;
}
inline bool pass_takes_tma(const KernelKey &key, const PassParams &p)
{
    if (pass_takes_tma_xpose(key, p) || pass_takes_tma_in(key, p)) return true;
    if (!((tunables().tma_col_mask >> key.log2n) & 1)) return false;
    if (key.layout != LAYOUT_COL || key.variant != VAR_PLAIN || p.tw_on || p.out_peer_on || p.in_peer_on || p.grid_cap > 0) return false;
    if (key.log2n < 7 || key.log2n > 10) return false;
    if (p.in_eshift <= kMaxLog2N || p.out_eshift <= kMaxLog2N || p.logA != 0 || p.in_s2 != 1 || p.out_s2 != 1) return false;
    const u64 inner = 1ull << p.logB, N = 1ull << key.log2n, L = (u64)lines_per_tile(key.log2n, LAYOUT_COL);
    if (p.logB > 28 || inner < L || (inner % L) != 0) return false;
    if (p.in_es != (i64)inner || p.out_es != (i64)inner || p.in_s0 != (i64)(N * inner) || p.out_s0 != p.in_s0) return false;
    return (p.q_begin % L) == 0 && ((p.q_end - p.q_begin) % L) == 0 && p.q_end > p.q_begin;
}
// does this launch go to the big-tile kernel (if the backend has one for the key)?
inline bool use_big_tiles(const KernelKey &key, const PassParams &p)
{
    if (key.variant == VAR_REAL || p.out_peer_on || p.in_peer_on || p.grid_cap > 0) return false;
    const int mask = key.layout == LAYOUT_ROW ? tunables().big_row_mask : tunables().big_col_mask;
    return ((mask >> key.log2n) & 1) != 0;
}

// Side lane of a plan: steps tagged lane 1 (the speq-plane passes of rlft3: 4 launches of ~10 us that depend only on
// the z pass) run on a second stream, forked from the caller's stream at the first of them and joined before the first
// lane-0 step that touches the speq plane again, or at the end of the program.  Created on first use.
struct SideLane {
    void *stream, *ev_fork, *ev_join;
    SideLane() : stream(nullptr), ev_fork(nullptr), ev_join(nullptr) {}
};

struct Plan {
    int kind;
    std::vector<size_t> dims;
    size_t batch;
    Program prog[2];       // [0]: isign = +1, [1]: isign = -1
    size_t ws_elems;       // workspace size in complex elements
    void *ws;              // device workspace (owned)
    void *sched;           // device scratch for fused-launch counters (owned)
    size_t sched_bytes;
    int device;
    SideLane side;
    std::vector<TableRef> tables;   // twiddle tables the programs point into
    Plan() : kind(0), batch(1), ws_elems(0), ws(nullptr), sched(nullptr), sched_bytes(0), device(0) {}
};
void release_side_lane(SideLane &sl);

// builders; return NRB_* codes
int build_plan(Plan &pl, int kind, const size_t *dims, size_t ndim, size_t batch);
int exec_plan(Plan &pl, double *d_io, double *d_aux, double *d_out, int isign, int arg, void *stream);
// same as exec_plan but brackets every launch with events; ms[i] = duration of launch i (blocks)
int profile_plan(Plan &pl, double *d_io, double *d_aux, double *d_out, int isign, int arg, void *stream, float *ms,
                 int cap);
// description of launch i: kernel name and the bytes it must read + write (algorithmic)
int describe_launch(const Plan &pl, int isign, int idx, char *name, size_t cap, double *bytes);
int fill_uniform_device(double *d_out, u64 seed, u64 offset, u64 count, void *stream);
int complex_multiply_device(double *d_a, const double *d_b, u64 ncomplex, int conj_b, double scale, void *stream);

// slab-decomposed rlft3 (one rank's share)
struct SlabPlan {
    size_t nn1, nn2, nn3;
    bool real;             // true: rlft3 (nn3 real points per line, speq plane); false: 3-D complex fourn (nn3 complex points)
    int nranks, rank;
    Program prog[2][2];    // [isign index][stage]
    // pipelined exchange: the volume is cut into `chunks` z-ranges; part[isign][0] = work that precedes the
    // chunks (forward: z pass), part[isign][1 + 2c] / [2 + 2c] = stage 0 / stage 1 of chunk c,
    // part[isign][1 + 2*chunks] = work that follows them (inverse: z pass)
    int chunks;
    std::vector<Program> part[2];
    // DMA exchange (slab_set_dma): stage 0 writes a local send buffer whose blocks are laid out chunk-major, so the
    // piece (peer, chunk) is contiguous and copy engines push it into the peer's receive buffer; per-chunk flags
    bool dma;
    void *send;                  // owned, xchg size
    void *copy_stream, *side_stream, *ev_go, *ev_side, *ev_copy, *ev_s0[16];
    // extra copy streams (tunable dma_streams > 1): the pieces for different peers go to different streams so that
    // several copy engines work at once; copy_stream waits for them before it publishes the chunk's flag
    void *copy_extra[3], *ev_extra[3];
    bool timeline;               // diagnostics: timing events around every piece of the last exec_slab_dma
    std::vector<void *> tl_events;
    std::vector<std::string> tl_names;
    size_t ws_elems;
    void *ws;
    SideLane side;         // speq_side: the small speq-plane passes of a stage run beside its data pass
    double2 *peers[8];     // peer receive buffers (fused exchange); peers[rank] is the local one
    double2 *sends[8];     // peer SEND buffers (push + pull exchange: stage 1 pulls the high-z part of every block from the
                           // producer's send buffer); sends[rank] is the local one; all null = everything is pushed
    bool fused;
    std::vector<TableRef> tables;   // twiddle tables the programs point into
    // N3 = complex points per z line; BLK = complex elements per exchange block (the speq part rides along for rlft3)
    size_t n3c() const { return real ? nn3 / 2 : nn3; }
    size_t blk() const { const size_t G = (size_t)nranks; return (nn1 / G) * (nn2 / G) * (n3c() + (real ? 1 : 0)); }
    size_t xchg_elems() const { return (size_t)nranks * blk(); }
    SlabPlan() : nn1(0), nn2(0), nn3(0), real(true), nranks(1), rank(0), chunks(1), dma(false), send(nullptr), copy_stream(nullptr), side_stream(nullptr),
                 ev_go(nullptr), ev_side(nullptr), ev_copy(nullptr), ev_s0{}, copy_extra{}, ev_extra{}, timeline(false), ws_elems(0), ws(nullptr), peers{}, sends{}, fused(false) {}
};
int build_slab_plan(SlabPlan &sp, size_t nn1, size_t nn2, size_t nn3, int nranks, int rank, bool real = true);
// one whole direction of the fused exchange on `stream`: stage 0 (stores go to the peers), epoch-flag barrier, stage 1
int exec_slab_fused(SlabPlan &sp, int isign, double *d_slab, double *d_speq, unsigned long long epoch, void *stream);
// cut the exchange into `chunks` z-ranges (1 = off); allocates the out-of-place work slab
int slab_set_chunks(SlabPlan &sp, int chunks);
// part: -1 = before the chunks, chunks = after them, else stage `stage` of chunk `part` (fused exchange only)
int exec_slab_part(SlabPlan &sp, int stage, int part, int isign, double *d_slab, double *d_speq, void *stream,
                   double *d_xchg = nullptr);
// DMA exchange: chunk-major block layout + plan-owned send buffer, streams and events (chunks >= 1)
int slab_set_dma(SlabPlan &sp, int chunks);
// one direction of the DMA-pipelined transform, enqueued on `stream` (plus the plan's copy and side streams)
int exec_slab_dma(SlabPlan &sp, int isign, double *d_slab, double *d_speq, unsigned long long epoch, void *stream);
void slab_release(SlabPlan &sp);
// diagnostics: after a synchronise, "name=ms since the start" of every piece of the last exec_slab_dma
std::string slab_dma_timeline(SlabPlan &sp);
int slab_barrier_chunk(SlabPlan &sp, int phase, int chunk, unsigned long long epoch, void *stream);
int slab_set_peers(SlabPlan &sp, void *const *peer_recv, int count);
// push + pull exchange: every rank's send buffer (xchg size), mapped like the receive buffers; nullptr switches it off
int slab_set_send_peers(SlabPlan &sp, void *const *peer_send, int count);
// flag barrier of the fused exchange (flags live right after the exchange area of each receive buffer)
int slab_barrier(SlabPlan &sp, int phase /*0 = signal, 1 = wait*/, unsigned long long epoch, void *stream);
int exec_slab_stage(SlabPlan &sp, int stage, int isign, double *d_slab, double *d_speq, double *d_send,
                    double *d_recv, void *stream);

inline bool is_pow2(size_t n) { return n && !(n & (n - 1)); }
inline int ilog2(size_t n) { int l = 0; while ((size_t(1) << l) < n) ++l; return l; }

} // namespace nrb
