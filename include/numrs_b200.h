/*
 * numrs_b200.h -- C ABI of libnumrs_b200.so: the B200-native (sm_100a) implementation of the
 * numrs FFT hot path (four1 / fourn / realft / rlft3 / convlv / correl).
 *
 * This header is the drop-in boundary.  The reference
 * (SciRustaceans/numrs) is a pure-Rust crate with no FFI of its own; each entry point below
 * is what the reference's public Rust function of the same name would bind through
 * `extern "C"` (see INTEGRATION.md for the Rust shim).  Signatures use plain pointers and
 * sizes only.  All citations are file:line in /root/reference/src.
 *
 * Conventions (identical to the reference):
 *   - complex data is interleaved f64: data[2k] = Re, data[2k+1] = Im;
 *   - isign = +1 uses exp(+2*pi*i*j*k/N), isign = -1 uses exp(-2*pi*i*j*k/N);
 *   - no normalisation in either direction (four1/fourn round trip = N*x,
 *     realft round trip = (n/2)*x, rlft3 round trip = (nn1*nn2*nn3/2)*x);
 *   - sizes are powers of two (the reference never validates this; this library returns
 *     NRB_ERR_NOT_POW2 instead of computing garbage).
 *
 * Host-slice entry points take HOST pointers (caller-owned, not retained), run the transform
 * on the current device (nrb_set_device) and block until the result is back in host memory.
 * They are thread-safe (per-thread stream and staging; plan cache behind a mutex).
 * There is no CPU fallback: without a CUDA device every compute entry point returns
 * NRB_ERR_CUDA.
 */
#ifndef NUMRS_B200_H
#define NUMRS_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes (mapped by the Rust shim to ConvlvError / CorrelError /
 *      io::ErrorKind::InvalidInput / panic) ---- */
#define NRB_OK                     0
#define NRB_ERR_EMPTY_INPUT       -1   /* Convolve.rs:13-15, Correlation.rs:11-13 */
#define NRB_ERR_RESPONSE_TOO_LONG -2   /* Convolve.rs:16-18 */
#define NRB_ERR_INVALID_ISIGN     -3   /* Convolve.rs:19-21, Fourn.rs:371-373, Real_FT3.rs:17 */
#define NRB_ERR_LENGTH_MISMATCH   -4   /* Correlation.rs:14-16 */
#define NRB_ERR_INVALID_DIMS      -5   /* Fourn.rs:368-370,374-376, Real_FT.rs:5 (n odd) */
#define NRB_ERR_NOT_POW2          -6   /* not validated by the reference */
#define NRB_ERR_UNSUPPORTED       -7   /* shape outside what this build implements */
#define NRB_ERR_ZERO_STDDEV       -8   /* Correlation.rs:214-216,246-248 CorrelError::ZeroStdDev */
#define NRB_ERR_CUDA              -10
#define NRB_ERR_NCCL              -11
#define NRB_ERR_OOM               -12

/* convlv response padding (SURVEY.md 8c ledger L1) */
#define NRB_PAD_LITERAL 0              /* Convolve.rs:41-63 as written (default) */
#define NRB_PAD_NR      1              /* Numerical Recipes convlv wrap-around */

/* ---- library ---- */
const char *nrb_version(void);
const char *nrb_last_error(void);      /* thread-local, never NULL */
int  nrb_device_count(void);           /* number of CUDA devices, 0 if none */
int  nrb_set_device(int device);       /* device used by the calling thread */
int  nrb_shutdown(void);               /* frees cached plans, twiddle tables no live plan uses and the staging buffers /
                                          streams of every thread that made host-slice calls; no other nrb_* call may be
                                          in flight.  Plans from nrb_plan_create / nrb_slab_create stay valid: each holds a
                                          reference on the twiddle tables it uses, released by its own destroy call. */
/* Planner tunables (affect plans created afterwards; cached host-call plans are dropped):
 *   "col_max_log2"   longest strided-axis FFT done in one pass          (default 10)
 *   "row_max_log2"   longest contiguous FFT done in one pass            (default 13)
 *   "l2_group_bytes"    rlft3: bytes of x-planes per z/y launch pair    (default: whole volume)
 *   "batch_group_bytes" convlv/correl: bytes of signals per launch group (default 512 MiB)
 *   "fuse_zy"           rlft3: run the z and y passes of every x-plane in one persistent launch so the
 *                       y pass reads the z pass's output from L2 (default 0: measured 3-6 % slower on B200); "fuse_lag" = planes the y
 *                       tiles trail behind the z tiles (default 16)
 *   "conv_transposed"   convlv/correl with lines longer than a tile: two passes per transform, spectrum kept in transposed
 *                       order (default 1); 0 = three natural-order passes per transform
 *   "prefetch_dist"     tiles ahead whose input every CTA prefetches into L2 (-1 = per-kernel policy, 0 = off)
 *   "num_devices"       GPUs ONE host-slice call is spread over, inside this process: 1 = the calling thread's device
 *                       (default), 0 = every visible device, n = devices 0 .. n-1 (rounded down to a power of two, <= 8).
 *                       nrb_rlft3 and 3-D nrb_fourn scatter slabs of the host volume over the devices' PCIe links, exchange
 *                       over NVLink peer memory and gather the result; the *_batch calls shard contiguous batch ranges.
 *   "pipeline_batches"  1 (default): a host-slice batch call of >= 64 MiB runs in chunks over three streams, so the H2D copy
 *                       of one chunk, the transforms of the previous and the D2H copy of the one before overlap; 0 = one shot
 *   "pipeline_min_kb"   smallest chunk of a pipelined batch call in KiB (default 16384; a call needs at least four of them)
 *   "shard_min_kb"      batches smaller than this stay on one device (default 16384)
 *   "z_chunks"          slab stages: 2 = the z pass and the exchange pass beside it run as two halves of the local y rows,
 *                       the z pass of one half on a side stream under the exchange pass of the other (default 1 = off)
 *   "pull_eighths"      push + pull slab exchange: eighths of the z range pulled by stage 1 (0 .. 8, default 4)
 *   "dma_streams"       DMA slab exchange: copy streams the pieces of a chunk are spread over (1 .. 4, default 1)
 *   "tma_xpose"         1 (default): the transposing 1024-point pass of a multi-step transform loads its tile by TMA
 *   "tma_col_mask"      bit log2 N set: eligible strided passes of N points use the TMA-fed kernel (default 512 | 1024)
 *   "tma_persist"       1: the TMA-fed strided pass as a persistent CTA with two tile buffers (default 0: measured no faster)
 *   "tma_in_mask" / "tma_in_ctas"  bit log2 N set: strided passes of N points load their tile by TMA and store from registers
 *                       (also eligible with a four-step twiddle or a transposed output), 2 or 3 CTAs per SM (default 0 / 2:
 *                       measured within noise of the fully TMA-fed pass, kept for A/B)
 *   "xchg_grid_cap"     pipelined slab exchange: CTAs of a peer-store pass (0 = one per tile); fewer CTAs leave SM
 *                       slots to the local pass running beside it on the second stream
 *   "simple_addr"       1 (default): passes whose element index is not split take the cheap addressing code path where it
 *                       is built (contiguous 8192-point lines, strided 512 / 1024-point lines); 0 = general path (A/B)
 *   "conv_fused_mid"    long-line convlv / correl / autocorrel_fast: contiguous forward pass + spectral step + contiguous
 *                       inverse pass of a row pair in ONE kernel, 5 -> 3 passes per signal (default 1)
 *   "conv_rest_log2"    log2 of the contiguous row length of the long-line convlv / correl pipeline (default 12 = 4096-point
 *                       rows, one fused-middle CTA per SM; 11 = 2048-point rows, two CTAs per SM: measured slower overall)
 *   "mid_prefetch"      fused conv middle: tiles ahead whose rows a CTA prefetches into L2 (default 0 = off: measured slower)
 *   "speq_side"         rlft3: the four small speq-plane launches run on a second stream beside the data passes
 *                       (default 1)
 *   "trig_fused"        1 (default): cosft1 / cosft2 / sinft of 16 .. 16384 points and twofft of 8 .. 8192 points per line run
 *                       as ONE kernel and one HBM pass (pre-processing, realft, running sums / packing, four1, separation
 *                       all on chip); 0 = the multi-launch programs (A/B, tests)
 *   "big_row_mask" / "big_col_mask"  bit log2(n) set: lines of n points use the big-tile prefetching pass (default 0:
 *                       measured no faster, kept as an experiment)
 * Environment overrides at load time: NRB_COL_MAX_LOG2, NRB_ROW_MAX_LOG2, NRB_L2_GROUP_MB, NRB_BATCH_GROUP_MB,
 * NRB_SIMPLE_ADDR, NRB_CONV_FUSED_MID, NRB_SPEQ_SIDE, NRB_BIG_ROW_MASK, NRB_BIG_COL_MASK. */
int  nrb_set_option(const char *name, long value);
/* pinned host memory, so host-slice calls copy at full PCIe rate (optional) */
/* Multi-device introspection: devices one host-slice call is spread over under the current "num_devices" option, and the
 * number of calls that really took the multi-device path (which = 0: 3-D transforms, 1: sharded batches) or the chunked
 * three-stream pipeline of the batch calls (which = 2). */
int   nrb_num_devices_in_use(void);
long  nrb_multi_device_calls(int which);
void *nrb_host_alloc(size_t bytes);
void  nrb_host_free(void *p);

/* ---- host-slice drop-in entry points ---- */

/* FFT_1.rs:5  pub fn four1(data: &mut [f64], nn: usize, isign: i32); data has 2*nn doubles */
int nrb_four1(double *data, size_t nn, int isign);
/* FFT_1.rs:185 FFTProcessor::fft_batch(&self, batches: &mut [&mut [f64]], isign): `count`
 * independent transforms, slice b has 2*nn[b] doubles; slices need not be contiguous. */
int nrb_four1_batch(double *const *ptrs, const size_t *nn, size_t count, int isign);
/* Real_FT3.rs:35 call shape `Fourn(&mut flat, &nn, ndim, isign)` (NR in-memory fourn, the
 * reference's Fourn.rs has no in-memory body); validation as Fourn.rs:367-378.
 * data has 2*prod(nn[0..ndim]) doubles, row-major, nn[ndim-1] fastest. */
int nrb_fourn(double *data, const size_t *nn, size_t ndim, int isign);
/* Real_FT.rs:4  pub fn realft(data: &mut [f64], n: usize, isign: i32); isign == 1 forward,
 * anything else inverse (Real_FT.rs:10,15).  Packed output: data[0]=F_0, data[1]=F_{n/2}. */
int nrb_realft(double *data, size_t n, int isign);
/* Real_FT.rs:365 RealFTProcessor::process_batch: `count` real transforms of length n each */
int nrb_realft_batch(double *const *ptrs, size_t n, size_t count, int isign);
/* Real_FT3.rs:8 rlft3(data: Array3 [nn1][nn2][nn3], speq: Array2 [nn1][2*nn2], .., isign) */
int nrb_rlft3(double *data, double *speq, size_t nn1, size_t nn2, size_t nn3, int isign);
/* Convolve.rs:8 convlv(data, respns, isign) -> Array1 of length n, written to ans[0..n) */
int nrb_convlv(const double *data, size_t n, const double *respns, size_t m, int isign,
               int pad_mode, double *ans);
/* Convolve.rs:241 convlv_batch: `count` signals of length n, one response */
int nrb_convlv_batch(const double *const *data, size_t count, size_t n, const double *respns,
                     size_t m, int isign, int pad_mode, double *const *ans);
/* Correlation.rs:8 correl(data1, data2) -> Array1 of length n1, written to ans[0..n1) */
int nrb_correl(const double *data1, size_t n1, const double *data2, size_t n2, double *ans);
/* Correlation.rs:273 correl_batch: `count` pairs, each of length n */
int nrb_correl_batch(const double *const *data1, const double *const *data2, size_t count,
                     size_t n, double *const *ans);

/* ---- callers on either side of the hot path (SURVEY.md 8f "next" rows N2, N4) ---- */
/* Correlation.rs:189 correl_normalized (fast = 0) / :226 correl_normalized_fast (fast = 1): population mean / std
 * of both inputs (two-pass resp. single-pass formulas), NRB_ERR_ZERO_STDDEV when either std is 0, then correl of
 * the normalised signals (the fast variant's n <= 32 branch also divides by n, Correlation.rs:251-263). */
int nrb_correl_normalized(const double *data1, size_t n1, const double *data2, size_t n2, int fast, double *ans);
/* Correlation.rs:286 autocorrel_fast: n <= 32 direct lags, else one forward realft, |F|^2, inverse realft */
int nrb_autocorrel_fast(const double *data, size_t n, double *ans);
/* FFT_2.rs:3 twofft(data1, data2, fft1, fft2): spectra of two real signals from one complex four1(+1);
 * fft1 / fft2 have 2n + 2 doubles (FFT_2.rs:6-7): n complex bins followed by two zero doubles. */
int nrb_twofft(const double *data1, const double *data2, size_t n, double *fft1, double *fft2);
/* FFT_2.rs:258 TwoFFTProcessor::process_batch: `count` independent signal pairs of the same length n in one batched
 * plan (the reference loops over the tuples and calls twofft on each); fft1[b] / fft2[b] have 2n + 2 doubles. */
int nrb_twofft_batch(const double *const *data1, const double *const *data2, size_t count, size_t n,
                     double *const *fft1, double *const *fft2);
/* FFT_1.rs:218 power_spectrum (take_sqrt = 0) / :206 magnitude_spectrum (take_sqrt = 1) of npoints complex points */
int nrb_power_spectrum(const double *complex_data, size_t npoints, int take_sqrt, double *out);

/* N3: cosine / sine transforms around realft (NR semantics; the reference's bodies stop at `unimplemented!()`,
 * Cos_FT.rs:70-74).  The reference's 1-based calling convention is kept: y[0] is unused.
 * Cos_FT.rs:7   cosft1(y, n): y has n + 2 doubles, data y[1..=n+1];  F_k = f_0/2 + (-1)^k f_n/2 + sum f_j cos(pi j k / n)
 * Cos_FT2.rs:7  cosft2(y, n, isign): y has n + 1 doubles, data y[1..=n]; isign = 1: F_k = sum f_j cos(pi k (j + 1/2) / n);
 *               isign = -1: the inverse up to the factor 2/n; any other isign -> NRB_ERR_INVALID_ISIGN (Cos_FT2.rs:11 panics)
 * README.md:72  sinft(y, n): y has n + 1 doubles, data y[1..=n]; y[1] is taken as 0; F_k = sum f_j sin(pi j k / n) */
int nrb_cosft1(double *y, size_t n);
int nrb_cosft2(double *y, size_t n, int isign);
int nrb_sinft(double *y, size_t n);

/* ---- device-resident plan API (what the benchmark times; pointers are DEVICE memory) ---- */
typedef struct nrb_plan_s *nrb_plan_t;

#define NRB_KIND_FOUR1  1   /* dims = {nn};            io = batch x 2*nn doubles            */
#define NRB_KIND_FOURN  2   /* dims = nn[0..ndim);     io = batch x 2*prod(nn) doubles      */
#define NRB_KIND_REALFT 3   /* dims = {n};             io = batch x n doubles               */
#define NRB_KIND_RLFT3  4   /* dims = {nn1,nn2,nn3};   io = data, aux = speq                */
#define NRB_KIND_CONVLV 5   /* dims = {n, m};          io = batch x n signals (read only),
                               aux = m response taps, out = batch x n doubles              */
#define NRB_KIND_CORREL 6   /* dims = {n};             io = data1, aux = data2 (batch x n,
                               read only), out = batch x n doubles; n > 32                 */

#define NRB_KIND_CORREL_NORM      7  /* dims = {n}; io = data1, aux = data2; out = 2*batch (mean, std) pairs
                                        (data1's signals, then data2's) followed by batch x n answers    */
#define NRB_KIND_CORREL_NORM_FAST 8  /* same, single-pass statistics (Correlation.rs:226)                 */
#define NRB_KIND_AUTOCORREL_FAST  9  /* dims = {n}; io = data, out = batch x n doubles                    */
#define NRB_KIND_TWOFFT          10  /* dims = {n}; io = data1, aux = data2 (batch x n doubles); out = fft1
                                        [batch][n+1] complex followed by fft2 [batch][n+1] complex       */
#define NRB_KIND_POWER           11  /* dims = {npoints}; io = complex points, out = doubles; arg = sqrt  */
#define NRB_KIND_COSFT1          12  /* dims = {n}; io = batch x (n + 2) doubles (1-based lines), in place */
#define NRB_KIND_COSFT2          13  /* dims = {n}; io = batch x (n + 1) doubles, in place; isign = +-1    */
#define NRB_KIND_SINFT           14  /* dims = {n}; io = batch x (n + 1) doubles, in place                 */

int    nrb_plan_create(int kind, const size_t *dims, size_t ndim, size_t batch, nrb_plan_t *plan);
size_t nrb_plan_workspace_bytes(nrb_plan_t plan);
int    nrb_plan_num_launches(nrb_plan_t plan, int isign);  /* kernels per exec */
/* Enqueue on `stream` (a cudaStream_t, NULL = default stream); does not synchronise.
 * `arg` is pad_mode for CONVLV, take_sqrt for POWER and ignored otherwise. */
int    nrb_plan_exec(nrb_plan_t plan, double *d_io, double *d_aux, double *d_out, int isign,
                     int arg, void *stream);
int    nrb_plan_destroy(nrb_plan_t plan);
/* Measurement helpers.  nrb_plan_profile runs the plan once with CUDA events around every
 * launch and returns ms[i] for launch i (blocks until done); nrb_plan_describe_launch gives
 * launch i's kernel name and the bytes it has to read + write (its algorithmic traffic). */
int    nrb_plan_profile(nrb_plan_t plan, double *d_io, double *d_aux, double *d_out, int isign,
                        int arg, void *stream, float *ms, int cap);
int    nrb_plan_describe_launch(nrb_plan_t plan, int isign, int idx, char *name, size_t cap,
                                double *bytes);
/* Synthetic input: out[i] = uniform[-1,1) from splitmix64(seed*phi+offset+i),
 * written to DEVICE memory; bit-identical to the generator in oracle/ and in Python. */
int    nrb_fill_uniform_device(double *d_out, unsigned long long seed, unsigned long long offset,
                               size_t count, void *stream);

/* ---- device-resident buffers: chains such as rlft3 -> pointwise product -> rlft3^-1 stay in HBM.
 * Buffers come from nrb_device_alloc / nrb_device_free (below); copies are asynchronous on `stream` (NULL = the
 * default stream) and full speed only from / to pinned host memory (nrb_host_alloc); nrb_stream_synchronize blocks
 * until everything enqueued on `stream` is done.  nrb_complex_multiply_device: a[i] = a[i] * b[i] * scale, or
 * a[i] * conj(b[i]) * scale when conj_b != 0, over ncomplex interleaved complex elements (what NR's 3-D convolution
 * example does between rlft3 and its inverse, applied to `data` and to `speq`). */
int    nrb_upload(void *d_dst, const void *h_src, size_t bytes, void *stream);
int    nrb_download(void *h_dst, const void *d_src, size_t bytes, void *stream);
int    nrb_stream_synchronize(void *stream);
int    nrb_complex_multiply_device(double *d_a, const double *d_b, size_t ncomplex, int conj_b, double scale, void *stream);

/* ---- slab-decomposed rlft3 across the GPUs of one box (one process per GPU) ----
 * Forward: rank r holds the nn2-slab data[:, r*nn2/G:(r+1)*nn2/G, :] as a contiguous
 * [nn1][nn2/G][nn3] array.  stage 0 = z real transform + x transform on the slab (output
 * in exchange layout: G blocks [nn1/G][nn2/G][nn3/2] complex, block p goes to rank p);
 * the caller exchanges blocks (all-to-all over NCCL/NVLink); stage 1 = y transform reading
 * the received blocks and writing the nn1-slab [nn1/G][nn2][nn3] plus speq [nn1/G][2*nn2].
 * Inverse (isign=-1) runs the mirror image: stage 0 on the nn1-slab, exchange, stage 1. */
typedef struct nrb_slab_s *nrb_slab_t;
int    nrb_slab_create(size_t nn1, size_t nn2, size_t nn3, int nranks, int rank, nrb_slab_t *plan);
/* The same decomposition for the in-memory 3-D complex Fourn(data, [nn1, nn2, nn3], 3, isign) (call shape Real_FT3.rs:35):
 * slabs of nn3 complex points per line, no speq plane (nrb_slab_speq_doubles = 0, d_speq may be NULL).  isign = +1 takes
 * nn2-slabs [nn1][nn2/G][nn3] and returns nn1-slabs [nn1/G][nn2][nn3]; isign = -1 is the mirror image. */
int    nrb_slab_create_fourn(size_t nn1, size_t nn2, size_t nn3, int nranks, int rank, nrb_slab_t *plan);
size_t nrb_slab_local_doubles(nrb_slab_t plan);   /* doubles in one slab (data)           */
size_t nrb_slab_speq_doubles(nrb_slab_t plan);    /* doubles in the local speq part        */
size_t nrb_slab_xchg_doubles(nrb_slab_t plan);    /* doubles in the exchange buffer        */
int    nrb_slab_stage(nrb_slab_t plan, int stage, int isign, double *d_slab, double *d_speq,
                      double *d_send, double *d_recv, void *stream);
/* Fused exchange over NVLink peer memory: give the plan every rank's receive buffer (each at least
 * nrb_slab_xchg_doubles() doubles, allocated with nrb_device_alloc and mapped into this process with
 * nrb_ipc_export / nrb_ipc_import; peer_recv[rank] is the local one).  Stage 0 then stores its output
 * DIRECTLY into the owning peer's buffer from the FFT kernel's epilogue (no send buffer, no NCCL
 * all-to-all: d_send / d_recv are ignored); the caller only has to put a cross-rank barrier between
 * stage 0 and stage 1.  Passing NULL returns the plan to the explicit send/recv mode. */
int    nrb_slab_set_peers(nrb_slab_t plan, void *const *peer_recv, int count);
/* Push + pull exchange (on top of nrb_slab_set_peers): peer_send[i] = rank i's SEND buffer (nrb_slab_xchg_doubles()
 * doubles, mapped like the receive buffers).  Stage 0 then pushes only the low-z part of every block and leaves the rest
 * (option "pull_eighths" / 8 of the z range, default half) in its own send buffer, from which the consumer's stage 1 reads
 * it over NVLink: the link works during both passes of a direction.  NULL switches back to pushing everything.  Like the
 * receive buffers, send buffers must be double buffered by the caller across calls (dist_rlft3.py). */
int    nrb_slab_set_send_peers(nrb_slab_t plan, void *const *peer_send, int count);
/* Size of one receive buffer for the fused mode: the exchange area plus a small flag array used by
 * nrb_slab_barrier (nrb_device_alloc returns zeroed memory). */
size_t nrb_slab_recv_bytes(nrb_slab_t plan);
/* Collective-free barrier of the fused mode, enqueued on `stream`: phase 0 (after stage 0) publishes
 * `epoch` into this rank's slot of every peer's flag array; phase 1 (before stage 1) spins on the
 * device until all ranks have published `epoch` locally; phase 2 does both in one launch (what nrb_slab_exec uses).
 * Epochs must increase per receive buffer. */
int    nrb_slab_barrier(nrb_slab_t plan, int phase, unsigned long long epoch, void *stream);
/* One whole direction of the fused exchange in one call: stage 0, signal + wait on epoch `epoch`, stage 1. */
int    nrb_slab_exec(nrb_slab_t plan, int isign, double *d_slab, double *d_speq, unsigned long long epoch, void *stream);
int    nrb_slab_num_launches(nrb_slab_t plan, int isign);   /* kernels nrb_slab_exec launches (both stages + the flag barrier) */
/* Pipelined exchange (fused mode only): cut the volume into `chunks` z-ranges (a power of two, at most 16;
 * 1 = off) so that stage 1 of chunk c can run -- on a second stream -- under the NVLink-bound stores of chunk
 * c + 1.  nrb_slab_stage_part runs one piece: part = -1 is the work before the chunks (forward: the z pass),
 * part = chunks the work after them (inverse: the z pass), otherwise stage `stage` (0 or 1) of chunk `part`.
 * Order per direction:  part -1;  for every c: [stage 0 of c, barrier_chunk(0, c)] on the main stream and
 * [barrier_chunk(1, c), stage 1 of c] on the side stream;  join the streams;  part `chunks`.
 * Stage 1 no longer runs in place behind stage 0, so the plan owns one extra slab of workspace. */
int    nrb_slab_set_chunks(nrb_slab_t plan, int chunks);
int    nrb_slab_stage_part(nrb_slab_t plan, int stage, int part, int isign, double *d_slab, double *d_speq, void *stream);
int    nrb_slab_barrier_chunk(nrb_slab_t plan, int phase, int chunk, unsigned long long epoch, void *stream);
/* DMA-pipelined exchange: stage 0 writes a plan-owned send buffer whose blocks are laid out chunk-major
 * ([chunk][nn1/G][nn2/G][nn3/2/chunks]), so the piece (peer, chunk) is contiguous; copy engines -- not SMs -- push
 * the pieces into the peers' receive buffers (nrb_slab_set_peers) and a one-warp kernel then publishes the chunk's
 * epoch flag; stage 1 of the chunk runs on the plan's high-priority side stream as soon as every rank's flag is in.
 * nrb_slab_exec_dma enqueues one whole direction (work before the chunks, all chunks, work after them) and makes
 * `stream` wait for the plan's internal streams at the end.  `chunks` is a power of two, 1 ... 16.
 * nrb_slab_stage_part_xchg is the step-wise form (tests): stage 0 of a chunk writes d_xchg, stage 1 reads it. */
int    nrb_slab_set_dma(nrb_slab_t plan, int chunks);
int    nrb_slab_exec_dma(nrb_slab_t plan, int isign, double *d_slab, double *d_speq, unsigned long long epoch, void *stream);
int    nrb_slab_stage_part_xchg(nrb_slab_t plan, int stage, int part, int isign, double *d_slab, double *d_speq, double *d_xchg,
                                void *stream);
/* diagnostics: (after a synchronise) writes "piece=ms since the start" for every piece of the last nrb_slab_exec_dma
 * into `text`, then switches the recording of those timing events on or off for the next calls */
int    nrb_slab_dma_timeline(nrb_slab_t plan, int enable, char *text, size_t cap);
int    nrb_slab_destroy(nrb_slab_t plan);
/* raw (zero-initialised) device memory + CUDA IPC plumbing for the above */
int    nrb_device_alloc(size_t bytes, void **dptr);
int    nrb_device_free(void *dptr);
int    nrb_ipc_export(void *dptr, unsigned char handle[64]);
int    nrb_ipc_import(const unsigned char handle[64], void **dptr);
int    nrb_ipc_release(void *dptr);

#ifdef __cplusplus
}
#endif
#endif /* NUMRS_B200_H */
