// plan.cpp -- host planner for the numrs_b200 FFT hot path (see plan.h).
//
// Every transform is expressed as a sequence of launches of one generic kernel family
// (fft_pass.cuh) plus a few elementwise kernels (aux_kernels.cuh):
//   * an FFT along one axis of a row-major view [outer][N][inner] is one pass when N fits a CTA
//     tile, otherwise a multi-step ("four-step") decomposition N = F1*F2*..: pass t transforms
//     the F_t-long sub-axis, multiplies by exp(-+2 pi i n_rest k / N_cur) and stores
//     transposed, so the remaining sub-axis becomes an ordinary strided axis again;
//   * real transforms (realft, the z axis of rlft3) are a c2c pass with NR's untangling fused
//     in (VAR_REAL), or c2c passes + a standalone untangle kernel for lines longer than a tile;
//   * rlft3 is evaluated separably (z real pass, then y and x complex passes on data and on
//     the speq plane), which is algebraically identical to NR's "fourn then untangle"
//     (Real_FT3.rs:31-127) because the untangling is linear and the mirror in (i1,i2) commutes
//     with the y/x transforms.  z and y passes run on groups of x-planes sized to stay
//     L2-resident so the pair costs one HBM round trip.
#include "plan.h"

#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <mutex>

#include "../../include/numrs_b200.h"

namespace nrb {

// ------------------------------------------------------------------ errors / tunables
static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
const std::string &get_error() { return g_err; }

static int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

static Tunables &tunables_mut()
{
    static Tunables t = [] {
        Tunables x;
        x.col_max_log2 = env_int("NRB_COL_MAX_LOG2", 10);
        x.row_max_log2 = env_int("NRB_ROW_MAX_LOG2", 13);
        x.l2_group_bytes = (u64)env_int("NRB_L2_GROUP_MB", 1 << 20) << 20;
        x.batch_group_bytes = (u64)env_int("NRB_BATCH_GROUP_MB", 512) << 20;
        x.fuse_zy = env_int("NRB_FUSE_ZY", 0);
        x.fuse_lag = env_int("NRB_FUSE_LAG", 16);
        x.xchg_grid_cap = env_int("NRB_XCHG_GRID_CAP", 0);
        x.prefetch_dist = env_int("NRB_PREFETCH_DIST", -1);
        x.conv_transposed = env_int("NRB_CONV_TRANSPOSED", 1);
        x.simple_addr = env_int("NRB_SIMPLE_ADDR", 1);
        x.speq_side = env_int("NRB_SPEQ_SIDE", 1);
        x.conv_fused_mid = env_int("NRB_CONV_FUSED_MID", 1);
        x.conv_rest_log2 = env_int("NRB_CONV_REST_LOG2", 12);
        x.mid_prefetch = env_int("NRB_MID_PREFETCH", 0);
        x.big_row_mask = env_int("NRB_BIG_ROW_MASK", 0);
        x.big_col_mask = env_int("NRB_BIG_COL_MASK", 0);
        x.dma_streams = env_int("NRB_DMA_STREAMS", 1);
        x.pull_eighths = env_int("NRB_PULL_EIGHTHS", 4);
        x.z_chunks = env_int("NRB_Z_CHUNKS", 1);
        x.tma_col_mask = env_int("NRB_TMA_COL_MASK", (1 << 9) | (1 << 10));
        x.tma_persist = env_int("NRB_TMA_PERSIST", 0);
        x.tma_xpose = env_int("NRB_TMA_XPOSE", 1);
        x.tma_in_mask = env_int("NRB_TMA_IN_MASK", 0);
        x.tma_in_ctas = env_int("NRB_TMA_IN_CTAS", 2);
        x.num_devices = env_int("NRB_NUM_DEVICES", 1);
        x.shard_min_kb = env_int("NRB_SHARD_MIN_KB", 16384);
        x.pipeline_batches = env_int("NRB_PIPELINE_BATCHES", 1);
        x.pipeline_min_kb = env_int("NRB_PIPELINE_MIN_KB", 16384);
        x.trig_fused = env_int("NRB_TRIG_FUSED", 1);
        return x;
    }();
    if (t.col_max_log2 < 1) t.col_max_log2 = 1;
    if (t.col_max_log2 > 12) t.col_max_log2 = 12;
    if (t.row_max_log2 < 1) t.row_max_log2 = 1;
    if (t.row_max_log2 > kMaxLog2N) t.row_max_log2 = kMaxLog2N;
    if (t.l2_group_bytes < 1024) t.l2_group_bytes = 1024;
    if (t.batch_group_bytes < 1024) t.batch_group_bytes = 1024;
    if (t.fuse_lag < 1) t.fuse_lag = 1;
    return t;
}
const Tunables &tunables() { return tunables_mut(); }

int set_tunable(const char *name, long value)
{
    Tunables &t = tunables_mut();
    const std::string n(name ? name : "");
    if (n == "col_max_log2") t.col_max_log2 = (int)value;
    else if (n == "row_max_log2") t.row_max_log2 = (int)value;
    else if (n == "l2_group_bytes") t.l2_group_bytes = (u64)value;
    else if (n == "batch_group_bytes") t.batch_group_bytes = (u64)value;
    else if (n == "fuse_zy") t.fuse_zy = (int)value;
    else if (n == "fuse_lag") t.fuse_lag = (int)value;
    else if (n == "xchg_grid_cap") t.xchg_grid_cap = (int)value;
    else if (n == "prefetch_dist") t.prefetch_dist = (int)value;
    else if (n == "conv_transposed") t.conv_transposed = (int)value;
    else if (n == "simple_addr") t.simple_addr = (int)value;
    else if (n == "speq_side") t.speq_side = (int)value;
    else if (n == "conv_fused_mid") t.conv_fused_mid = (int)value;
    else if (n == "mid_prefetch") t.mid_prefetch = value < 0 ? 0 : (int)value;
    else if (n == "conv_rest_log2") t.conv_rest_log2 = value < 1 ? 1 : value > 12 ? 12 : (int)value;
    else if (n == "big_row_mask") t.big_row_mask = (int)value;
    else if (n == "big_col_mask") t.big_col_mask = (int)value;
    else if (n == "z_chunks") t.z_chunks = (value == 2 || value == 4) ? (int)value : 1;
    else if (n == "pull_eighths") t.pull_eighths = value < 0 ? 0 : value > 8 ? 8 : (int)value;
    else if (n == "dma_streams") t.dma_streams = value < 1 ? 1 : value > 4 ? 4 : (int)value;
    else if (n == "tma_col_mask") t.tma_col_mask = (int)value;
    else if (n == "tma_persist") t.tma_persist = (int)value;
    else if (n == "tma_xpose") t.tma_xpose = value != 0;
    else if (n == "tma_in_mask") t.tma_in_mask = (int)value;
    else if (n == "tma_in_ctas") t.tma_in_ctas = value >= 3 ? 3 : 2;
    else if (n == "num_devices") t.num_devices = value < 0 ? 1 : (int)value;
    else if (n == "pipeline_batches") t.pipeline_batches = value != 0;
    else if (n == "pipeline_min_kb") t.pipeline_min_kb = value < 1 ? 1 : (int)value;
    else if (n == "shard_min_kb") t.shard_min_kb = value < 0 ? 0 : (int)value;
    else if (n == "trig_fused") t.trig_fused = value != 0;
    else return -1;
    tunables_mut();   // re-clamp
    return 0;
}

// ------------------------------------------------------------------ twiddle tables
namespace {
struct TableKey {
    int dev, kind, log;
    bool operator<(const TableKey &o) const
    {
        if (dev != o.dev) return dev < o.dev;
        if (kind != o.kind) return kind < o.kind;
        return log < o.log;
    }
};
std::mutex g_tab_mu;
// never destroyed: at process exit the CUDA context may already be gone, so nothing is freed from a static destructor
std::map<TableKey, TableRef> &g_tables = *new std::map<TableKey, TableRef>();
thread_local std::vector<TableRef> *t_table_sink = nullptr;   // the plan being built on this thread (TableScope)

// exp(-2 pi i num / den) with long double trigonometry
double2 unit_root(u64 num, u64 den)
{
    const long double pi2 = 6.283185307179586476925286766559005768L;
    // reduce to the first octant for accuracy
    num %= den;
    const long double a = pi2 * (long double)num / (long double)den;
    double2 r;
    r.x = (double)cosl(a);
    r.y = (double)(-sinl(a));
    // exact values on the axes
    if (num == 0) { r.x = 1.0; r.y = 0.0; }
    else if (num * 4 == den) { r.x = 0.0; r.y = -1.0; }
    else if (num * 2 == den) { r.x = -1.0; r.y = 0.0; }
    else if (num * 4 == den * 3) { r.x = 0.0; r.y = 1.0; }
    return r;
}

const void *get_table(int kind, int log, std::vector<double2> (*gen)(int))
{
    std::lock_guard<std::mutex> lk(g_tab_mu);
    TableKey key{be_current_device(), kind, log};
    auto it = g_tables.find(key);
    if (it != g_tables.end()) {
        if (t_table_sink) t_table_sink->push_back(it->second);
        return it->second.get();
    }
    std::vector<double2> host = gen(log);
    if (host.empty()) host.push_back(unit_root(0, 1));
    void *d = nullptr;
    if (be_malloc(&d, host.size() * sizeof(double2)) != 0) return nullptr;
    if (be_h2d(d, host.data(), host.size() * sizeof(double2), nullptr) != 0 || be_sync(nullptr) != 0) { be_free(d); return nullptr; }
    const int dev = key.dev;
    TableRef ref(d, [dev](void *p) {          // freed when the cache and the last plan using it have let go
        const int cur = be_current_device();
        if (cur != dev) be_set_device(dev);
        be_free(p);
        if (cur != dev && cur >= 0) be_set_device(cur);
    });
    g_tables[key] = ref;
    if (t_table_sink) t_table_sink->push_back(ref);
    return d;
}

std::vector<double2> gen_stage(int log2n)
{
    const RadixPlan rp = radix_plan(log2n);
    std::vector<double2> t((size_t)stage_tw_total(log2n));
    for (int s = 1; s < rp.nst; ++s) {
        const int R = rp.r[s], NS = stage_ns(log2n, s), off = stage_tw_off(log2n, s);
        for (int jm = 0; jm < NS; ++jm)
            for (int r = 1; r < R; ++r)
                t[(size_t)off + (size_t)jm * (R - 1) + (r - 1)] = unit_root((u64)jm * r, (u64)NS * R);
    }
    return t;
}
int fourstep_h(int log2m) { return (log2m + 1) / 2; }
std::vector<double2> gen_fs_lo(int log2m)
{
    const int h = fourstep_h(log2m);
    std::vector<double2> t((size_t)1 << h);
    for (u64 i = 0; i < t.size(); ++i) t[i] = unit_root(i, 1ull << log2m);
    return t;
}
std::vector<double2> gen_fs_hi(int log2m)
{
    const int h = fourstep_h(log2m);
    std::vector<double2> t((size_t)1 << (log2m - h));
    for (u64 i = 0; i < t.size(); ++i) t[i] = unit_root(i << h, 1ull << log2m);
    return t;
}
std::vector<double2> gen_real(int log2n)
{
    const u64 N = 1ull << log2n;
    std::vector<double2> t((size_t)N);
    for (u64 k = 0; k < t.size(); ++k) t[k] = unit_root(k, 2 * N);
    return t;
}
} // namespace

const double2 *stage_twiddles(int log2n) { return (const double2 *)get_table(0, log2n, gen_stage); }
FourStepTable fourstep_table(int log2m)
{
    FourStepTable f;
    f.lo = (const double2 *)get_table(1, log2m, gen_fs_lo);
    f.hi = (const double2 *)get_table(2, log2m, gen_fs_hi);
    f.h = fourstep_h(log2m);
    return f;
}
const double2 *real_twiddles(int log2n) { return (const double2 *)get_table(3, log2n, gen_real); }

TableScope::TableScope(std::vector<TableRef> *sink) : prev(t_table_sink) { t_table_sink = sink; }
TableScope::~TableScope() { t_table_sink = prev; }

// drops the cache's references: tables no live plan points into are freed now, the others with their last plan
void release_tables()
{
    std::lock_guard<std::mutex> lk(g_tab_mu);
    g_tables.clear();
}

// ------------------------------------------------------------------ program builder
namespace {

struct AxisMap {        // custom addressing for one side of a single-pass axis
    bool on;
    i64 s0;             // stride of the outer index
    i64 es;             // stride of the low part of the element index
    int eshift;         // element index split
    i64 es_hi;          // stride of the high part
    AxisMap() : on(false), s0(0), es(0), eshift(30), es_hi(0) {}
};

struct Builder {
    Program *prog;
    size_t ws_used;     // workspace high-water mark (complex elements)
    size_t sched_used;  // fused-launch counter scratch (bytes)
    int rc;
    explicit Builder(Program *p) : prog(p), ws_used(0), sched_used(0), rc(NRB_OK) {}
    void need_ws(size_t end) { if (end > ws_used) ws_used = end; }
};

void init_pass(PassParams &pp)
{
    memset(&pp, 0, sizeof(pp));
    pp.in_eshift = 30;
    pp.out_eshift = 30;
    pp.logA = 0;
    pp.logB = 40;
    pp.peer_zthr = 1;      // with peer_zmask = 0: always the first half of the exchange pointer tables
}

u64 tiles_for(int log2n, int layout, u64 lines)
{
    const u64 L = (u64)lines_per_tile(log2n, layout);
    return (lines + L - 1) / L;
}

// split log2(N) into per-pass factors
std::vector<int> choose_factors(int p, u64 inner)
{
    const Tunables &T = tunables();
    std::vector<int> f;
    const int single_max = (inner == 1) ? T.row_max_log2 : T.col_max_log2;
    if (p <= single_max) { f.push_back(p); return f; }
    const int cm = T.col_max_log2;
    const int m = (p + cm - 1) / cm;
    const int base = p / m, rem = p % m;
    for (int i = 0; i < m; ++i) f.push_back(base + (i < rem ? 1 : 0));
    return f;
}

// complex FFT of length 2^p along the middle axis of the view [outer][2^p][inner] living in
// `src`, result in `dst` (same layout; src may equal dst).  Only outer indices
// [o_begin, o_end) are processed.  `tmp` is scratch for multi-step axes (must not alias).
void emit_axis(Builder &B, BufRef src, BufRef dst, BufRef tmp, u64 outer, u64 o_begin, u64 o_end, int p,
               u64 inner, int dir, const AxisMap *in_map = nullptr, const AxisMap *out_map = nullptr)
{
    (void)outer;
    if (p == 0 || o_begin >= o_end) return;
    const u64 N = 1ull << p;
    const std::vector<int> fac = choose_factors(p, inner);
    const int m = (int)fac.size();
    const int logI = ilog2((size_t)inner);

    if (m == 1) {
        Step st;
        init_pass(st.pp);
        PassParams &pp = st.pp;
        if (inner == 1) {
            st.key = KernelKey{p, LAYOUT_ROW, dir, VAR_PLAIN};
            pp.q_begin = o_begin; pp.q_end = o_end;
            pp.logA = 0; pp.logB = 40;
            pp.in_s2 = (i64)N; pp.in_es = 1;
            pp.out_s2 = (i64)N; pp.out_es = 1;
        } else {
            st.key = KernelKey{p, LAYOUT_COL, dir, VAR_PLAIN};
            pp.q_begin = o_begin * inner; pp.q_end = o_end * inner;
            pp.logA = 0; pp.logB = logI;
            pp.in_s0 = (i64)(N * inner); pp.in_s2 = 1; pp.in_es = (i64)inner;
            pp.out_s0 = (i64)(N * inner); pp.out_s2 = 1; pp.out_es = (i64)inner;
        }
        if (in_map && in_map->on) {
            if (inner == 1) pp.in_s2 = in_map->s0; else pp.in_s0 = in_map->s0;
            pp.in_es = in_map->es; pp.in_eshift = in_map->eshift; pp.in_es_hi = in_map->es_hi;
        }
        if (out_map && out_map->on) {
            if (inner == 1) pp.out_s2 = out_map->s0; else pp.out_s0 = out_map->s0;
            pp.out_es = out_map->es; pp.out_eshift = out_map->eshift; pp.out_es_hi = out_map->es_hi;
        }
        pp.tw = stage_twiddles(p);
        st.in = src; st.out = dst;
        st.ntiles = tiles_for(p, st.key.layout, pp.q_end - pp.q_begin);
        B.prog->steps.push_back(st);
        return;
    }

    if ((in_map && in_map->on) || (out_map && out_map->on)) { B.rc = NRB_ERR_UNSUPPORTED; return; }

    // multi-step: transposing passes 0..m-2, then a final strided pass
    const u64 seg = N * inner;                       // elements per outer index
    B.need_ws((size_t)(tmp.off + (i64)(o_end * seg)));  // tmp mirrors the layout of src/dst
    int pcur = p;
    u64 inn = inner;
    BufRef cur = src;
    for (int t = 0; t < m - 1; ++t) {
        const int f = fac[t];
        const u64 F = 1ull << f, rest = 1ull << (pcur - f);
        const BufRef nxt = (t % 2 == 0) ? tmp : dst;
        Step st;
        init_pass(st.pp);
        PassParams &pp = st.pp;
        const bool xpose = (inn == 1);
        st.key = KernelKey{f, LAYOUT_COL, dir, xpose ? VAR_XPOSE : VAR_PLAIN};
        pp.logB = ilog2((size_t)inn);
        pp.logA = pcur - f;
        pp.q_begin = o_begin * rest * inn; pp.q_end = o_end * rest * inn;
        pp.in_s0 = (i64)seg; pp.in_s1 = (i64)inn; pp.in_s2 = 1; pp.in_es = (i64)(rest * inn);
        pp.out_s0 = (i64)seg; pp.out_s1 = (i64)(F * inn); pp.out_s2 = 1; pp.out_es = (i64)inn;
        const FourStepTable fs = fourstep_table(pcur);
        pp.tw_on = 1; pp.tw_lo = fs.lo; pp.tw_hi = fs.hi; pp.tw_h = fs.h;
        pp.tw = stage_twiddles(f);
        st.in = cur; st.out = nxt;
        st.ntiles = tiles_for(f, st.key.layout, pp.q_end - pp.q_begin);
        B.prog->steps.push_back(st);
        cur = nxt;
        pcur -= f;
        inn *= F;
    }
    {
        Step st;
        init_pass(st.pp);
        PassParams &pp = st.pp;
        st.key = KernelKey{pcur, LAYOUT_COL, dir, VAR_PLAIN};
        const u64 Nl = 1ull << pcur;
        pp.logA = 0; pp.logB = ilog2((size_t)inn);
        // view [outer * (N / Nl / ... )]: after the transposes the layout is [o][Nl][inn] with inn = seg / Nl
        pp.q_begin = o_begin * inn; pp.q_end = o_end * inn;
        pp.in_s0 = (i64)seg; pp.in_s2 = 1; pp.in_es = (i64)inn;
        pp.out_s0 = (i64)seg; pp.out_s2 = 1; pp.out_es = (i64)inn;
        (void)Nl;
        pp.tw = stage_twiddles(pcur);
        st.in = cur; st.out = dst;
        st.ntiles = tiles_for(pcur, st.key.layout, pp.q_end - pp.q_begin);
        B.prog->steps.push_back(st);
    }
}

// real transform of `lines` lines of n = 2N real points (N = 2^p complex), line l at
// src + l*N.  dir=+1: forward (NR realft isign=1), dir=-1: inverse.  real_mode selects where
// the Nyquist bin lives (packed into element 0, or the speq plane).
void emit_real(Builder &B, BufRef src, BufRef dst, BufRef tmp, u64 l_begin, u64 l_end, int p, int dir,
               int real_mode, BufRef speq)
{
    if (l_begin >= l_end) return;
    const u64 N = 1ull << p;
    if (p >= 1 && p <= tunables().row_max_log2) {
        Step st;
        init_pass(st.pp);
        PassParams &pp = st.pp;
        st.key = KernelKey{p, LAYOUT_ROW, dir, VAR_REAL};
        pp.q_begin = l_begin; pp.q_end = l_end;
        pp.logA = 0; pp.logB = 40;
        pp.in_s2 = (i64)N; pp.in_es = 1;
        pp.out_s2 = (i64)N; pp.out_es = 1;
        pp.tw = stage_twiddles(p);
        pp.rtw = real_twiddles(p);
        pp.real_mode = real_mode;
        st.in = src; st.out = dst; st.speq = speq;
        st.ntiles = tiles_for(p, st.key.layout, l_end - l_begin);
        B.prog->steps.push_back(st);
        return;
    }
    // standalone untangle (+ c2c passes when N > 1)
    Step un;
    un.is_aux = true;
    memset(&un.ap, 0, sizeof(un.ap));
    un.ap.kind = AUX_UNTANGLE;
    un.ap.op = real_mode;
    un.ap.dir = dir;
    un.ap.n = N;
    un.ap.count = l_end - l_begin;
    un.ap.a_stride = un.ap.out_stride = (i64)N;
    const FourStepTable rt = fourstep_table(p + 1);   // exp(-2 pi i k / 2N) = exp(-i pi k / N)
    un.ap.rtw_lo = rt.lo; un.ap.rtw_hi = rt.hi; un.ap.rtw_h = rt.h;
    un.speq = (speq.id == BUF_NONE) ? speq : speq + (i64)l_begin;
    if (dir > 0) {
        emit_axis(B, src, dst, tmp, l_end, l_begin, l_end, p, 1, +1);
        un.in = (p == 0 ? src : dst) + (i64)(l_begin * N);
        un.out = dst + (i64)(l_begin * N);
        B.prog->steps.push_back(un);
    } else {
        un.in = src + (i64)(l_begin * N);
        un.out = dst + (i64)(l_begin * N);
        B.prog->steps.push_back(un);
        emit_axis(B, dst, dst, tmp, l_end, l_begin, l_end, p, 1, -1);
    }
}

// c2c part only of a long real transform (lines longer than a CTA tile): the untangling is left to
// the fused spectral kernel (AUX_SPECTRAL_Z)
bool real_needs_separate_untangle(int p) { return p > tunables().row_max_log2; }

Step make_spectral_z(int op, BufRef a, BufRef b, BufRef out, u64 n, u64 count, i64 a_stride, i64 b_stride, i64 out_stride)
{
    Step st;
    st.is_aux = true;
    memset(&st.ap, 0, sizeof(st.ap));
    st.ap.kind = AUX_SPECTRAL_Z;
    st.ap.op = op;
    st.ap.n = n;
    st.ap.count = count;
    st.ap.a_stride = a_stride; st.ap.b_stride = b_stride; st.ap.out_stride = out_stride;
    const FourStepTable rt = fourstep_table(ilog2((size_t)n));   // exp(-2 pi i k / n) = exp(-i pi k / N)
    st.ap.rtw_lo = rt.lo; st.ap.rtw_hi = rt.hi; st.ap.rtw_h = rt.h;
    st.in = a; st.b = b; st.out = out;
    return st;
}

// ---- two-pass transforms with the spectrum left in transposed order (convlv / correl intermediates) ----
// N = 2^p = F * REST, F = 2^f.  Forward: strided F-point pass over n = j*REST + r for every r (COL, lines adjacent),
// output kf multiplied by exp(-+2 pi i r kf / N), stored in place; then contiguous REST-point pass per kf (ROW).
// Bin kf + F*kr ends at position kf*REST + kr.  Inverse: ROW pass per kf with the twiddle on its output, then the
// COL pass; the result is in natural order again.  No scratch, no transposing stores.
int conv_split(int p)
{
    const Tunables &T = tunables();
    if (!T.conv_transposed || p < 2) return -1;
    const int rmax = T.conv_rest_log2;
    int rest = p - 1 < rmax ? p - 1 : rmax;             // contiguous pass: 4096 points at most (the fast ROW kernels)
    if (rest > T.row_max_log2) rest = T.row_max_log2;
    int f = p - rest;
    if (f > T.col_max_log2) { f = T.col_max_log2; rest = p - f; }
    if (f < 1 || rest < 1 || rest > T.row_max_log2) return -1;
    return f;
}

void emit_conv_forward(Builder &B, BufRef src, BufRef dst, u64 cnt, int p, int f, bool with_row = true)
{
    const u64 N = 1ull << p, F = 1ull << f, REST = N >> f;
    {
        Step st;
        init_pass(st.pp);
        PassParams &pp = st.pp;
        st.key = KernelKey{f, LAYOUT_COL, +1, VAR_PLAIN};
        pp.logB = 0; pp.logA = p - f;                   // line q = b*REST + r: q1 = r, q0 = b
        pp.q_begin = 0; pp.q_end = cnt * REST;
        pp.in_s0 = (i64)N; pp.in_s1 = 1; pp.in_s2 = 0; pp.in_es = (i64)REST;
        pp.out_s0 = (i64)N; pp.out_s1 = 1; pp.out_s2 = 0; pp.out_es = (i64)REST;
        const FourStepTable fs = fourstep_table(p);
        pp.tw_on = 1; pp.tw_lo = fs.lo; pp.tw_hi = fs.hi; pp.tw_h = fs.h;
        pp.tw = stage_twiddles(f);
        st.in = src; st.out = dst;
        st.ntiles = tiles_for(f, LAYOUT_COL, pp.q_end);
        B.prog->steps.push_back(st);
    }
    if (with_row) emit_axis(B, dst, dst, BufRef(), cnt * F, 0, cnt * F, p - f, 1, +1);
}

void emit_conv_inverse(Builder &B, BufRef buf, u64 cnt, int p, int f, bool with_row = true)
{
    const u64 N = 1ull << p, F = 1ull << f, REST = N >> f;
    if (with_row) {
        Step st;
        init_pass(st.pp);
        PassParams &pp = st.pp;
        st.key = KernelKey{p - f, LAYOUT_ROW, -1, VAR_PLAIN};
        pp.logB = 0; pp.logA = f;                       // line q = b*F + kf: q1 = kf, q0 = b
        pp.q_begin = 0; pp.q_end = cnt * F;
        pp.in_s0 = (i64)N; pp.in_s1 = (i64)REST; pp.in_s2 = 0; pp.in_es = 1;
        pp.out_s0 = (i64)N; pp.out_s1 = (i64)REST; pp.out_s2 = 0; pp.out_es = 1;
        const FourStepTable fs = fourstep_table(p);
        pp.tw_on = 1; pp.tw_lo = fs.lo; pp.tw_hi = fs.hi; pp.tw_h = fs.h;
        pp.tw = stage_twiddles(p - f);
        st.in = buf; st.out = buf;
        st.ntiles = tiles_for(p - f, LAYOUT_ROW, pp.q_end);
        B.prog->steps.push_back(st);
    }
    // COL pass over kf for every r: view [cnt][F][REST]
    emit_axis(B, buf, buf, BufRef(), cnt, 0, cnt, f, REST, -1);
}

// fused middle (conv_mid.cuh): rows of REST = 2^(p-f) points, `cnt` signals in `data` (in place), second operand `b`
bool conv_mid_usable(int p, int f)
{
    return tunables().conv_fused_mid && f >= 1 && be_conv_mid_available(p - f);
}
Step make_conv_mid(int op, BufRef data, BufRef b, u64 cnt, int p, int f, i64 b_stride)
{
    Step st;
    st.is_mid = true;
    memset(&st.mp, 0, sizeof(st.mp));
    st.key = KernelKey{p - f, LAYOUT_ROW, 0, VAR_PLAIN};
    st.mp.data_stride = (i64)1 << p; st.mp.b_stride = b_stride;
    st.mp.count = cnt; st.mp.f = f; st.mp.op = op;
    st.mp.tw = stage_twiddles(p - f);
    const FourStepTable fs = fourstep_table(p);
    st.mp.fs_lo = fs.lo; st.mp.fs_hi = fs.hi; st.mp.fs_h = fs.h;
    const FourStepTable rt = fourstep_table(p + 1);          // exp(-2 pi i k / n) = exp(-i pi k / N)
    st.mp.rtw_lo = rt.lo; st.mp.rtw_hi = rt.hi; st.mp.rtw_h = rt.h;
    st.in = data; st.out = data; st.b = b;
    st.ntiles = cnt * ((1ull << f) / 2);
    return st;
}

Step make_spectral_zt(int op, BufRef a, BufRef b, BufRef out, u64 n, int f, u64 count, i64 a_stride, i64 b_stride, i64 out_stride)
{
    Step st = make_spectral_z(op, a, b, out, n, count, a_stride, b_stride, out_stride);
    st.ap.kind = AUX_SPECTRAL_ZT;
    st.ap.m = (u64)f;
    return st;
}

Step make_spectral(int op, BufRef a, BufRef b, BufRef out, u64 n, u64 count, i64 a_stride, i64 b_stride,
                   i64 out_stride)
{
    Step st;
    st.is_aux = true;
    memset(&st.ap, 0, sizeof(st.ap));
    st.ap.kind = AUX_SPECTRAL;
    st.ap.op = op;
    st.ap.n = n;
    st.ap.count = count;
    st.ap.a_stride = a_stride; st.ap.b_stride = b_stride; st.ap.out_stride = out_stride;
    st.in = a; st.b = b; st.out = out;
    return st;
}

int check_pow2_dims(const size_t *dims, size_t ndim)
{
    for (size_t d = 0; d < ndim; ++d)
        if (!is_pow2(dims[d])) { set_error("dimension is not a power of two"); return NRB_ERR_NOT_POW2; }
    return NRB_OK;
}

// ---- per-kind program construction ----
int build_four1(Plan &pl, Builder &B, int dir)
{
    const int p = ilog2(pl.dims[0]);
    emit_axis(B, BufRef(BUF_IO, 0), BufRef(BUF_IO, 0), BufRef(BUF_WS, 0), pl.batch, 0, pl.batch, p, 1, dir);
    return B.rc;
}

int build_fourn(Plan &pl, Builder &B, int dir)
{
    const size_t nd = pl.dims.size();
    u64 total = 1;
    for (size_t d = 0; d < nd; ++d) total *= pl.dims[d];
    u64 inner = 1;
    for (size_t dd = nd; dd-- > 0;) {
        const u64 N = pl.dims[dd];
        const u64 outer = pl.batch * (total / (N * inner));
        emit_axis(B, BufRef(BUF_IO, 0), BufRef(BUF_IO, 0), BufRef(BUF_WS, 0), outer, 0, outer, ilog2((size_t)N),
                  inner, dir);
        inner *= N;
    }
    return B.rc;
}

int build_realft(Plan &pl, Builder &B, int dir)
{
    const int p = ilog2(pl.dims[0]) - 1;
    emit_real(B, BufRef(BUF_IO, 0), BufRef(BUF_IO, 0), BufRef(BUF_WS, 0), 0, pl.batch, p, dir, REAL_PACKED,
              BufRef());
    return B.rc;
}

int build_rlft3(Plan &pl, Builder &B, int dir)
{
    const u64 nn1 = pl.dims[0], nn2 = pl.dims[1], N3 = pl.dims[2] / 2;
    const int p1 = ilog2((size_t)nn1), p2 = ilog2((size_t)nn2), p3 = ilog2((size_t)N3);
    const BufRef D(BUF_IO, 0), S(BUF_AUX, 0), W(BUF_WS, 0);
    const u64 plane_bytes = nn2 * N3 * 16;
    u64 g = tunables().l2_group_bytes / (plane_bytes ? plane_bytes : 1);
    if (g < 1) g = 1;
    if (g > nn1) g = nn1;
    if (nn1 * plane_bytes <= 2 * tunables().l2_group_bytes) g = nn1;
    // Fused z+y: one persistent launch runs the z pass and the y pass of every x-plane with the
    // y tiles trailing `lag` planes behind, so the y pass reads the z pass's output from L2.
    const bool fusable = tunables().fuse_zy && g == nn1 && nn1 >= 2 && p3 >= 1 && p3 <= tunables().row_max_log2 &&
                         p2 >= 1 && p2 <= tunables().col_max_log2 && N3 >= (u64)col_line_count(p2) &&
                         nn2 >= (u64)lines_per_tile(p3, LAYOUT_ROW);
    auto emit_fused = [&](int d) -> bool {
        const KernelKey kz{p3, LAYOUT_ROW, d, VAR_REAL}, ky{p2, LAYOUT_COL, d, VAR_PLAIN};
        if (!fusable || !be_fused_available(d > 0 ? kz : ky, d > 0 ? ky : kz)) return false;
        Program tmp;
        Builder TB(&tmp);
        emit_real(TB, D, D, W, 0, nn1 * nn2, p3, d, REAL_SPEQ, S);
        emit_axis(TB, D, D, W, nn1, 0, nn1, p2, N3, d);
        if (tmp.steps.size() != 2 || TB.rc != NRB_OK) return false;
        const Step &z = tmp.steps[0], &y = tmp.steps[1];
        const Step &a = d > 0 ? z : y, &b = d > 0 ? y : z;
        Step st = a;
        st.is_fused = true;
        st.key2 = b.key; st.pp2 = b.pp; st.in2 = b.in; st.out2 = b.out; st.speq2 = b.speq;
        const u64 tz = nn2 / (u64)lines_per_tile(p3, LAYOUT_ROW), ty = N3 / (u64)col_line_count(p2);
        st.fs.units = (unsigned)nn1;
        st.fs.ta = (unsigned)(d > 0 ? tz : ty);
        st.fs.tb = (unsigned)(d > 0 ? ty : tz);
        u64 lag = (u64)tunables().fuse_lag;
        if (lag > nn1) lag = nn1;
        st.fs.lag = (unsigned)lag;
        st.sched_off = B.sched_used;
        B.sched_used += 16 + 4 * ((nn1 + 3) & ~(u64)3);
        st.ntiles = nn1 * (tz + ty);
        B.prog->steps.push_back(st);
        return true;
    };
    // speq-plane passes: single-pass axes only touch the speq buffer, so they may run beside the data passes (lane 1)
    const bool side = tunables().speq_side && g == nn1 && !fusable && p1 <= tunables().col_max_log2 && p2 <= tunables().row_max_log2;
    auto emit_speq = [&](int axis, int d) {
        const size_t first = B.prog->steps.size();
        if (axis == 2) emit_axis(B, S, S, W, nn1, 0, nn1, p2, 1, d);
        else emit_axis(B, S, S, W, 1, 0, 1, p1, nn2, d);
        if (side) for (size_t i = first; i < B.prog->steps.size(); ++i) B.prog->steps[i].lane = 1;
    };
    if (dir > 0 && side) {
        // z pass (writes data and speq), then the two speq passes on the side lane beside the y and x passes
        emit_real(B, D, D, W, 0, nn1 * nn2, p3, +1, REAL_SPEQ, S);
        emit_speq(2, +1);
        emit_speq(1, +1);
        emit_axis(B, D, D, W, nn1, 0, nn1, p2, N3, +1);
        emit_axis(B, D, D, W, 1, 0, 1, p1, nn2 * N3, +1);
    } else if (dir > 0) {
        if (!emit_fused(+1)) {
            for (u64 x0 = 0; x0 < nn1; x0 += g) {
                const u64 x1 = (x0 + g < nn1) ? x0 + g : nn1;
                emit_real(B, D, D, W, x0 * nn2, x1 * nn2, p3, +1, REAL_SPEQ, S);
                emit_axis(B, D, D, W, nn1, x0, x1, p2, N3, +1);
            }
        }
        emit_speq(2, +1);
        emit_axis(B, D, D, W, 1, 0, 1, p1, nn2 * N3, +1);
        emit_speq(1, +1);
    } else if (side) {
        emit_speq(1, -1);
        emit_speq(2, -1);
        emit_axis(B, D, D, W, 1, 0, 1, p1, nn2 * N3, -1);
        emit_axis(B, D, D, W, nn1, 0, nn1, p2, N3, -1);
        emit_real(B, D, D, W, 0, nn1 * nn2, p3, -1, REAL_SPEQ, S);   // reads speq: joins the side lane first
    } else {
        emit_speq(1, -1);
        emit_axis(B, D, D, W, 1, 0, 1, p1, nn2 * N3, -1);
        emit_speq(2, -1);
        if (!emit_fused(-1)) {
            for (u64 x0 = 0; x0 < nn1; x0 += g) {
                const u64 x1 = (x0 + g < nn1) ? x0 + g : nn1;
                emit_axis(B, D, D, W, nn1, x0, x1, p2, N3, -1);
                emit_real(B, D, D, W, x0 * nn2, x1 * nn2, p3, -1, REAL_SPEQ, S);
            }
        }
    }
    return B.rc;
}

// convlv: dims = {n, m}.  io = signals (read only), aux = response taps, out = answers.
// workspace: [0, n/2) response spectrum, [n/2, n/2 + gs*n/2) multi-step scratch.
int build_convlv(Plan &pl, Builder &B, int dir)
{
    const u64 n = pl.dims[0], m = pl.dims[1];
    const u64 N = n / 2;
    const int p = ilog2((size_t)N);
    const BufRef R(BUF_WS, 0), T(BUF_WS, (i64)N);
    B.need_ws((size_t)N);
    // response: pad (Convolve.rs:41-63) then forward realft, once per batch
    {
        Step st;
        st.is_aux = true;
        memset(&st.ap, 0, sizeof(st.ap));
        st.ap.kind = AUX_PAD_RESPONSE;
        st.ap.n = n; st.ap.m = m; st.ap.count = 1;
        st.in = BufRef(BUF_AUX, 0); st.out = R;
        st.patch_pad_mode = true;
        B.prog->steps.push_back(st);
    }
    const int fsplit = real_needs_separate_untangle(p) ? conv_split(p) : -1;
    // (two-pass plan: the response spectrum stays raw and transposed like the signals', untangled on the fly)
    if (fsplit > 0) emit_conv_forward(B, R, R, 1, p, fsplit);
    else emit_real(B, R, R, T, 0, 1, p, +1, REAL_PACKED, BufRef());
    // signals in L2-sized groups: forward, spectral op, inverse
    u64 gs = tunables().batch_group_bytes / (n * 8);
    if (gs < 1) gs = 1;
    if (gs > pl.batch) gs = pl.batch;
    for (u64 b0 = 0; b0 < pl.batch; b0 += gs) {
        const u64 b1 = (b0 + gs < pl.batch) ? b0 + gs : pl.batch;
        const BufRef in(BUF_IO, (i64)(b0 * N)), out(BUF_OUT, (i64)(b0 * N));
        if (fsplit > 0) {
            // two passes, spectrum in transposed order, [untangle + multiply + inverse untangle], two passes back
            if (conv_mid_usable(p, fsplit)) {
                // strided pass | contiguous forward pass + spectral step + contiguous inverse pass in one kernel | strided pass
                emit_conv_forward(B, in, out, b1 - b0, p, fsplit, false);
                B.prog->steps.push_back(make_conv_mid(dir > 0 ? SPEC_CONV_MUL : SPEC_CONV_DIV, out, R, b1 - b0, p, fsplit, 0));
                emit_conv_inverse(B, out, b1 - b0, p, fsplit, false);
                continue;
            }
            emit_conv_forward(B, in, out, b1 - b0, p, fsplit);
            Step sp = make_spectral_zt(dir > 0 ? SPEC_CONV_MUL : SPEC_CONV_DIV, out, R, out, n, fsplit, b1 - b0, (i64)N, 0, (i64)N);
            sp.ap.dir = 1;      // b = raw transposed response
            B.prog->steps.push_back(sp);
            emit_conv_inverse(B, out, b1 - b0, p, fsplit);
        } else if (real_needs_separate_untangle(p)) {
            // c2c -> [untangle + multiply + inverse untangle in one pass] -> inverse c2c
            emit_axis(B, in, out, T, b1 - b0, 0, b1 - b0, p, 1, +1);
            B.prog->steps.push_back(make_spectral_z(dir > 0 ? SPEC_CONV_MUL : SPEC_CONV_DIV, out, R, out, n, b1 - b0,
                                                    (i64)N, 0, (i64)N));
            emit_axis(B, out, out, T, b1 - b0, 0, b1 - b0, p, 1, -1);
        } else {
            emit_real(B, in, out, T, 0, b1 - b0, p, +1, REAL_PACKED, BufRef());
            B.prog->steps.push_back(make_spectral(dir > 0 ? SPEC_CONV_MUL : SPEC_CONV_DIV, out, R, out, n, b1 - b0,
                                                  (i64)N, 0, (i64)N));
            emit_real(B, out, out, T, 0, b1 - b0, p, -1, REAL_PACKED, BufRef());
        }
    }
    return B.rc;
}

Step make_aux(int kind, int op, BufRef a, BufRef b, BufRef out, BufRef stats, u64 n, u64 m, u64 count, i64 a_stride,
              i64 b_stride, i64 out_stride)
{
    Step st;
    st.is_aux = true;
    memset(&st.ap, 0, sizeof(st.ap));
    st.ap.kind = kind; st.ap.op = op;
    st.ap.n = n; st.ap.m = m; st.ap.count = count;
    st.ap.a_stride = a_stride; st.ap.b_stride = b_stride; st.ap.out_stride = out_stride;
    st.in = a; st.b = b; st.out = out; st.speq = stats;
    return st;
}

Step make_correl_direct(BufRef a, BufRef b, BufRef out, u64 n, u64 count)
{
    // Correlation.rs:19-21 direct branch (linear lags); strides in doubles
    return make_aux(AUX_CORREL_DIRECT, 0, a, b, out, BufRef(), n, 0, count, (i64)n, (i64)n, (i64)n);
}

// FFT correlation (n > 32) of `cnt` signal pairs: a, b read only (b ignored for SPEC_AUTOCORREL), F2 and T scratch
// of cnt * n/2 complex each.  Correlation.rs:23-33 with NR's correl for the placeholder (ledger D7).
void emit_correl_group(Builder &B, BufRef a, BufRef b, BufRef out, BufRef F2, BufRef T, u64 cnt, u64 n, int op)
{
    const u64 N = n / 2;
    const int p = ilog2((size_t)N);
    const bool two = op != SPEC_AUTOCORREL;
    const int fsplit = real_needs_separate_untangle(p) ? conv_split(p) : -1;
    if (fsplit > 0 && conv_mid_usable(p, fsplit)) {
        if (two) emit_conv_forward(B, b, F2, cnt, p, fsplit);          // the second operand's finished spectrum
        emit_conv_forward(B, a, out, cnt, p, fsplit, false);
        B.prog->steps.push_back(make_conv_mid(op, out, two ? F2 : out, cnt, p, fsplit, (i64)N));
        emit_conv_inverse(B, out, cnt, p, fsplit, false);
    } else if (fsplit > 0) {
        emit_conv_forward(B, a, out, cnt, p, fsplit);
        if (two) emit_conv_forward(B, b, F2, cnt, p, fsplit);
        B.prog->steps.push_back(make_spectral_zt(op, out, two ? F2 : out, out, n, fsplit, cnt, (i64)N, (i64)N, (i64)N));
        emit_conv_inverse(B, out, cnt, p, fsplit);
    } else if (real_needs_separate_untangle(p)) {
        emit_axis(B, a, out, T, cnt, 0, cnt, p, 1, +1);
        if (two) emit_axis(B, b, F2, T, cnt, 0, cnt, p, 1, +1);
        B.prog->steps.push_back(make_spectral_z(op, out, two ? F2 : out, out, n, cnt, (i64)N, (i64)N, (i64)N));
        emit_axis(B, out, out, T, cnt, 0, cnt, p, 1, -1);
    } else {
        emit_real(B, a, out, T, 0, cnt, p, +1, REAL_PACKED, BufRef());
        if (two) emit_real(B, b, F2, T, 0, cnt, p, +1, REAL_PACKED, BufRef());
        B.prog->steps.push_back(make_spectral(op, out, two ? F2 : out, out, n, cnt, (i64)N, (i64)N, (i64)N));
        emit_real(B, out, out, T, 0, cnt, p, -1, REAL_PACKED, BufRef());
    }
}

u64 correl_group_size(const Plan &pl, u64 n)
{
    u64 gs = tunables().batch_group_bytes / (n * 8);
    if (gs < 1) gs = 1;
    if (gs > pl.batch) gs = pl.batch;
    return gs;
}

// correl: dims = {n}.  io = data1, aux = data2 (read only), out = answers.
// autocorrel_fast (Correlation.rs:286-323, ledger D10): the same with one forward transform and |F|^2.
int build_correl(Plan &pl, Builder &B, int op)
{
    const u64 n = pl.dims[0];
    const BufRef IO(BUF_IO, 0), AUXB(op == SPEC_AUTOCORREL ? BUF_IO : BUF_AUX, 0), OUT(BUF_OUT, 0);
    if (n <= 32) {
        B.prog->steps.push_back(make_correl_direct(IO, AUXB, OUT, n, pl.batch));
        return B.rc;
    }
    const u64 N = n / 2, gs = correl_group_size(pl, n);
    const BufRef F2(BUF_WS, 0), T(BUF_WS, (i64)(gs * N));
    B.need_ws((size_t)(gs * N));
    for (u64 b0 = 0; b0 < pl.batch; b0 += gs) {
        const u64 b1 = (b0 + gs < pl.batch) ? b0 + gs : pl.batch;
        emit_correl_group(B, IO + (i64)(b0 * N), AUXB + (i64)(b0 * N), OUT + (i64)(b0 * N), F2, T, b1 - b0, n, op);
    }
    return B.rc;
}

// (mean, std) of S signals of n doubles (`a`: the first S/2 or all of them, `b`: the second half or none)
// into stats[S] (double2), by deterministic strided partial sums in two or three levels.
void emit_stats(Builder &B, BufRef a, BufRef b, BufRef stats, BufRef P0, BufRef P1, u64 S, u64 n, u64 C0, u64 C1, bool fast)
{
    auto level = [&](int red_op, int stats_op) {
        B.prog->steps.push_back(make_aux(AUX_REDUCE, red_op, a, b, P0, stats, n, C0, S, (i64)n, (i64)n, 0));
        if (C1) {
            B.prog->steps.push_back(make_aux(AUX_REDUCE, RED_PARTIALS, P0, BufRef(), P1, BufRef(), C0, C1, S, 0, 0, 0));
            B.prog->steps.push_back(make_aux(AUX_STATS_FINAL, stats_op, P1, BufRef(), stats, BufRef(), n, C1, S, 0, 0, 0));
        } else {
            B.prog->steps.push_back(make_aux(AUX_STATS_FINAL, stats_op, P0, BufRef(), stats, BufRef(), n, C0, S, 0, 0, 0));
        }
    };
    if (fast) level(RED_SUM_SQ, STATS_FAST);                        // Correlation.rs:236-244 (single pass)
    else { level(RED_SUM_SQ, STATS_MEAN); level(RED_CENTERED_SQ, STATS_STD); }   // Correlation.rs:199-212 (two passes)
}

// correl_normalized (Correlation.rs:189-223) / correl_normalized_fast (:226-270): dims = {n}.
// io = data1, aux = data2 (read only); out = [2*batch] (mean, std) pairs -- data1's signals first, then data2's;
// the host entry point turns a zero std into CorrelError::ZeroStdDev -- followed by batch x n answers.
int build_correl_norm(Plan &pl, Builder &B, bool fast)
{
    const u64 n = pl.dims[0], S = 2 * pl.batch;
    const BufRef IO(BUF_IO, 0), AUXB(BUF_AUX, 0), STATS(BUF_OUT, 0), ANS(BUF_OUT, (i64)S);
    const u64 C0 = n <= 64 ? 1 : (n / 64 < 16384 ? n / 64 : 16384);
    const u64 C1 = C0 > 256 ? 128 : 0;
    i64 off = 0;
    const BufRef P0(BUF_WS, off); off += (i64)(S * C0);
    const BufRef P1(BUF_WS, off); off += (i64)(S * C1);
    emit_stats(B, IO, AUXB, STATS, P0, P1, S, n, C0, C1, fast);
    if (n <= 32) {
        // normalise both inputs (strides in doubles), direct lags; the fast variant's own branch divides by n
        const i64 half = (i64)((pl.batch * n + 1) / 2);
        const BufRef NA(BUF_WS, off), NB = NA + half;
        B.need_ws((size_t)(off + 2 * half));
        B.prog->steps.push_back(make_aux(AUX_NORMALIZE, 0, IO, AUXB, NA, STATS, n, 0, S, (i64)n, (i64)n, (i64)n));
        // the second half of the signals lands at NA + batch*n doubles: keep NB there (may be odd -> address in doubles)
        Step cd = make_correl_direct(NA, NA, ANS, n, pl.batch);
        cd.ap.m = pl.batch * n;     // b = a + m doubles (see aux_correl_direct)
        B.prog->steps.push_back(cd);
        if (fast) B.prog->steps.push_back(make_aux(AUX_SCALE, 0, BufRef(), BufRef(), ANS, BufRef(), pl.batch * n, n, 1, 0, 0, 0));
        (void)NB;
        return B.rc;
    }
    const u64 N = n / 2, gs = correl_group_size(pl, n);
    const BufRef NA(BUF_WS, off); off += (i64)(gs * N);
    const BufRef NB(BUF_WS, off); off += (i64)(gs * N);
    const BufRef F2(BUF_WS, off); off += (i64)(gs * N);
    const BufRef T(BUF_WS, off);
    B.need_ws((size_t)off);
    for (u64 b0 = 0; b0 < pl.batch; b0 += gs) {
        const u64 b1 = (b0 + gs < pl.batch) ? b0 + gs : pl.batch, cnt = b1 - b0;
        // two launches (one per input) so a group's signals and their stats are contiguous ranges
        B.prog->steps.push_back(make_aux(AUX_NORMALIZE, 1, IO + (i64)(b0 * N), BufRef(), NA, STATS + (i64)b0, n, 0, cnt, (i64)n, 0, (i64)n));
        B.prog->steps.push_back(make_aux(AUX_NORMALIZE, 1, AUXB + (i64)(b0 * N), BufRef(), NB, STATS + (i64)(pl.batch + b0), n, 0, cnt, (i64)n, 0, (i64)n));
        emit_correl_group(B, NA, NB, ANS + (i64)(b0 * N), F2, T, cnt, n, SPEC_CORREL);
    }
    return B.rc;
}

// twofft (FFT_2.rs:3-17, ledger D9): dims = {n}.  io = data1, aux = data2 (n doubles each, read only);
// out = fft1 [batch][n + 1] complex followed by fft2 [batch][n + 1] complex.
int build_twofft(Plan &pl, Builder &B)
{
    const u64 n = pl.dims[0], per = n + 1;
    const BufRef F1(BUF_OUT, 0), F2(BUF_OUT, (i64)(pl.batch * per));
    const int p = ilog2((size_t)n);
    if (tunables().trig_fused && p <= tunables().row_max_log2 && be_trig_available(p)) {   // pack, four1 and separation in one kernel (trig_fused.cuh)
        Step st;
        st.trig = 2;
        st.key = KernelKey{p, LAYOUT_ROW, +1, VAR_PLAIN};
        memset(&st.fp, 0, sizeof(st.fp));
        st.fp.count = pl.batch;
        st.fp.tw = stage_twiddles(p);
        st.in = BufRef(BUF_IO, 0); st.b = BufRef(BUF_AUX, 0); st.out = F1; st.speq = F2;
        st.ntiles = pl.batch;
        B.prog->steps.push_back(st);
        return B.rc;
    }
    // short lines: transform in place in fft1 (lines n + 1 complex apart); multi-step transforms need densely
    // packed lines and go through the workspace
    const bool dense = p > tunables().row_max_log2;
    const BufRef D = dense ? BufRef(BUF_WS, 0) : F1, T(BUF_WS, (i64)(pl.batch * n));
    const i64 ds = dense ? (i64)n : (i64)per;
    if (dense) B.need_ws((size_t)(pl.batch * n));
    B.prog->steps.push_back(make_aux(AUX_PACK2, 0, BufRef(BUF_IO, 0), BufRef(BUF_AUX, 0), D, BufRef(), n, 0, pl.batch, (i64)n, (i64)n, ds));
    if (dense) {
        emit_axis(B, D, D, T, pl.batch, 0, pl.batch, p, 1, +1);                  // FFT_2.rs:13 four1(fft1, n, 1)
    } else if (p >= 1) {
        AxisMap map;
        map.on = true; map.s0 = (i64)per; map.es = 1; map.eshift = 30; map.es_hi = 0;
        emit_axis(B, D, D, T, pl.batch, 0, pl.batch, p, 1, +1, &map, &map);
    }
    B.prog->steps.push_back(make_aux(AUX_TWOFFT_SPLIT, 0, D, BufRef(), F1, F2, n, 0, pl.batch, ds, 0, (i64)per));
    return B.rc;
}

// cosft1 (Cos_FT.rs:7), cosft2 (Cos_FT2.rs:7) and sinft (README.md:72) -- NR's O(n) pre/post-processing around
// realft (ledger D11).  dims = {n}; io = batch lines of the reference's 1-based arrays (ld = n + 2 doubles for
// cosft1, n + 1 otherwise; element 0 of a line is unused).  Workspace: G (realft work array, n/2 complex per
// line), the pre-pass partial sums, the chunk sums of the running sum, one (sum, -) pair per line.
int build_cosft(Plan &pl, Builder &B, int kind, int dir)
{
    const u64 n = pl.dims[0], N = n / 2, L = pl.batch;
    const int p = ilog2((size_t)N);
    const i64 ld = (i64)(kind == NRB_KIND_COSFT1 ? n + 2 : n + 1);
    const int mode = kind == NRB_KIND_COSFT1 ? COS1 : kind == NRB_KIND_SINFT ? SINFT : COS2F;
    const bool inverse = kind == NRB_KIND_COSFT2 && dir < 0;
    if (tunables().trig_fused && p <= tunables().row_max_log2 && be_trig_available(p)) {   // the whole routine in one kernel (trig_fused.cuh)
        Step st;
        st.trig = 1;
        st.key = KernelKey{p, LAYOUT_ROW, dir, VAR_REAL};
        memset(&st.tp, 0, sizeof(st.tp));
        st.tp.ld = ld; st.tp.count = L; st.tp.mode = inverse ? COS2I_PRE : mode;
        st.tp.tw = stage_twiddles(p);
        st.tp.rtw = real_twiddles(p);
        const FourStepTable ct = fourstep_table(ilog2((size_t)n) + (kind == NRB_KIND_COSFT2 ? 2 : 1));
        st.tp.ctw_lo = ct.lo; st.tp.ctw_hi = ct.hi; st.tp.ctw_h = ct.h;
        st.in = BufRef(BUF_IO, 0);
        st.ntiles = L;
        B.prog->steps.push_back(st);
        return B.rc;
    }
    const u64 C = N + 1 < 16384 ? N + 1 : 16384;        // pre-pass accumulators per line
    const u64 C1 = C > 256 ? 128 : 0;
    const u64 K = (u64)kScanChunk, nch = (N + K - 1) / K;   // running sum: positions per chunk (one CTA each)
    i64 off = 0;
    const BufRef IO(BUF_IO, 0), G(BUF_WS, off); off += (i64)(L * N);
    const BufRef T(BUF_WS, off); off += real_needs_separate_untangle(p) ? (i64)(L * N) : 0;
    const BufRef P0(BUF_WS, off); off += (i64)(L * C);
    const BufRef P1(BUF_WS, off); off += (i64)(L * C1);
    const BufRef PC(BUF_WS, off); off += (i64)(L * nch);
    const BufRef INIT(BUF_WS, off); off += (i64)L;
    B.need_ws((size_t)off);
    const FourStepTable tw = fourstep_table(ilog2((size_t)n) + (kind == NRB_KIND_COSFT2 ? 2 : 1));
    auto with_tw = [&](Step st, int m) {
        st.ap.dir = m;
        st.ap.rtw_lo = tw.lo; st.ap.rtw_hi = tw.hi; st.ap.rtw_h = tw.h;
        return st;
    };
    if (inverse) {
        B.prog->steps.push_back(with_tw(make_aux(AUX_COSFT, 0, IO, BufRef(), G, BufRef(), n, C, L, ld, 0, 0), COS2I_PRE));
        emit_real(B, G, G, T, 0, L, p, -1, REAL_PACKED, BufRef());
        B.prog->steps.push_back(with_tw(make_aux(AUX_COSFT, 0, G, BufRef(), IO, BufRef(), n, C, L, 0, 0, ld), COS2I_POST));
        return B.rc;
    }
    B.prog->steps.push_back(with_tw(make_aux(AUX_COSFT, 0, IO, mode == COS1 ? P0 : BufRef(), G, BufRef(), n, C, L, ld, 0, 0), mode));
    if (mode == COS1) {
        if (C1) {
            B.prog->steps.push_back(make_aux(AUX_REDUCE, RED_PARTIALS, P0, BufRef(), P1, BufRef(), C, C1, L, 0, 0, 0));
            B.prog->steps.push_back(make_aux(AUX_STATS_FINAL, STATS_SUM, P1, BufRef(), INIT, BufRef(), n, C1, L, 0, 0, 0));
        } else {
            B.prog->steps.push_back(make_aux(AUX_STATS_FINAL, STATS_SUM, P0, BufRef(), INIT, BufRef(), n, C, L, 0, 0, 0));
        }
    }
    emit_real(B, G, G, T, 0, L, p, +1, REAL_PACKED, BufRef());
    for (int phase = 0; phase < 3; ++phase)
        B.prog->steps.push_back(with_tw(make_aux(AUX_SCAN, phase, G, PC, IO, INIT, n, K, L, 0, 0, ld), mode));
    return B.rc;
}

// power / magnitude spectrum (FFT_1.rs:206-228): dims = {npoints}; io = complex points, out = npoints doubles;
// exec's `arg` = 1 takes the square root.
int build_power(Plan &pl, Builder &B)
{
    Step st = make_aux(AUX_POWER, 0, BufRef(BUF_IO, 0), BufRef(), BufRef(BUF_OUT, 0), BufRef(), pl.dims[0] * pl.batch, 0, 1, 0, 0, 0);
    st.patch_pad_mode = true;
    B.prog->steps.push_back(st);
    return B.rc;
}

} // namespace

int build_plan(Plan &pl, int kind, const size_t *dims, size_t ndim, size_t batch)
{
    if (!dims || ndim == 0) { set_error("no dimensions"); return NRB_ERR_INVALID_DIMS; }
    if (batch == 0) { set_error("empty batch"); return NRB_ERR_EMPTY_INPUT; }
    pl.kind = kind;
    pl.dims.assign(dims, dims + ndim);
    pl.batch = batch;
    pl.device = be_current_device();
    TableScope tables(&pl.tables);
    int rc = NRB_OK;
    switch (kind) {
    case NRB_KIND_FOUR1:
        if (ndim != 1 || dims[0] == 0) { set_error("four1: nn must be >= 1"); return NRB_ERR_INVALID_DIMS; }
        if ((rc = check_pow2_dims(dims, 1))) return rc;
        break;
    case NRB_KIND_FOURN:
        for (size_t d = 0; d < ndim; ++d)
            if (dims[d] <= 1) { set_error("Invalid dimension size"); return NRB_ERR_INVALID_DIMS; } // Fourn.rs:374
        if ((rc = check_pow2_dims(dims, ndim))) return rc;
        break;
    case NRB_KIND_REALFT:
        if (ndim != 1 || dims[0] < 2 || (dims[0] & 1)) { set_error("realft: n must be even"); return NRB_ERR_INVALID_DIMS; }
        if ((rc = check_pow2_dims(dims, 1))) return rc;
        break;
    case NRB_KIND_RLFT3:
        if (ndim != 3 || dims[0] == 0 || dims[1] == 0 || dims[2] < 2) { set_error("rlft3: bad dimensions"); return NRB_ERR_INVALID_DIMS; }
        if ((rc = check_pow2_dims(dims, 3))) return rc;
        break;
    case NRB_KIND_CONVLV:
        if (ndim != 2) { set_error("convlv: dims = {n, m}"); return NRB_ERR_INVALID_DIMS; }
        if (dims[0] == 0 || dims[1] == 0) { set_error("Input arrays cannot be empty"); return NRB_ERR_EMPTY_INPUT; }
        if (dims[1] > dims[0]) { set_error("Response function longer than data"); return NRB_ERR_RESPONSE_TOO_LONG; }
        if (dims[0] < 2 || !is_pow2(dims[0])) { set_error("convlv: n must be a power of two >= 2"); return NRB_ERR_NOT_POW2; }
        break;
    case NRB_KIND_CORREL:
        if (ndim != 1 || dims[0] == 0) { set_error("Input arrays cannot be empty"); return NRB_ERR_EMPTY_INPUT; }
        if (dims[0] > 32 && !is_pow2(dims[0])) { set_error("correl: n > 32 must be a power of two"); return NRB_ERR_NOT_POW2; }
        break;
    case NRB_KIND_CORREL_NORM:
    case NRB_KIND_CORREL_NORM_FAST:
    case NRB_KIND_AUTOCORREL_FAST:
        if (ndim != 1 || dims[0] == 0) { set_error("Input arrays cannot be empty"); return NRB_ERR_EMPTY_INPUT; }
        if (dims[0] > 32 && !is_pow2(dims[0])) { set_error("correl: n > 32 must be a power of two"); return NRB_ERR_NOT_POW2; }
        break;
    case NRB_KIND_TWOFFT:
        if (ndim != 1 || dims[0] == 0) { set_error("twofft: n must be >= 1"); return NRB_ERR_INVALID_DIMS; }
        if ((rc = check_pow2_dims(dims, 1))) return rc;
        break;
    case NRB_KIND_POWER:
        if (ndim != 1 || dims[0] == 0) { set_error("power spectrum: empty input"); return NRB_ERR_EMPTY_INPUT; }
        break;
    case NRB_KIND_COSFT1:
    case NRB_KIND_COSFT2:
    case NRB_KIND_SINFT:
        if (ndim != 1 || dims[0] < 2) { set_error("cosft/sinft: n must be >= 2"); return NRB_ERR_INVALID_DIMS; }
        if ((rc = check_pow2_dims(dims, 1))) return rc;
        break;
    default:
        set_error("unknown plan kind");
        return NRB_ERR_INVALID_DIMS;
    }
    size_t ws = 0;
    pl.sched_bytes = 0;
    for (int s = 0; s < 2; ++s) {
        const int dir = s == 0 ? +1 : -1;
        Builder B(&pl.prog[s]);
        switch (kind) {
        case NRB_KIND_FOUR1: rc = build_four1(pl, B, dir); break;
        case NRB_KIND_FOURN: rc = build_fourn(pl, B, dir); break;
        case NRB_KIND_REALFT: rc = build_realft(pl, B, dir); break;
        case NRB_KIND_RLFT3: rc = build_rlft3(pl, B, dir); break;
        case NRB_KIND_CONVLV: rc = build_convlv(pl, B, dir); break;
        case NRB_KIND_CORREL: rc = build_correl(pl, B, SPEC_CORREL); break;
        case NRB_KIND_AUTOCORREL_FAST: rc = build_correl(pl, B, SPEC_AUTOCORREL); break;
        case NRB_KIND_CORREL_NORM: rc = build_correl_norm(pl, B, false); break;
        case NRB_KIND_CORREL_NORM_FAST: rc = build_correl_norm(pl, B, true); break;
        case NRB_KIND_TWOFFT: rc = build_twofft(pl, B); break;
        case NRB_KIND_POWER: rc = build_power(pl, B); break;
        case NRB_KIND_COSFT1: case NRB_KIND_COSFT2: case NRB_KIND_SINFT: rc = build_cosft(pl, B, kind, dir); break;
        }
        if (rc != NRB_OK) { set_error("shape not supported by this build"); return rc; }
        for (const Step &st : pl.prog[s].steps) {
            if (!st.is_aux && (!(st.trig == 1 ? st.tp.tw : st.trig == 2 ? st.fp.tw : st.is_mid ? st.mp.tw : st.pp.tw) && radix_plan(st.key.log2n).nst > 1)) { set_error(std::string("twiddle table allocation failed: ") + be_last_error()); return NRB_ERR_CUDA; }
        }
        if (B.ws_used > ws) ws = B.ws_used;
        if (B.sched_used > pl.sched_bytes) pl.sched_bytes = B.sched_used;
    }
    pl.ws_elems = ws;
    pl.ws = nullptr;
    if (ws) {
        const int mrc = be_malloc(&pl.ws, ws * sizeof(double2));
        if (mrc != 0) { set_error(std::string("workspace allocation failed: ") + be_last_error()); return NRB_ERR_OOM; }
    }
    pl.sched = nullptr;
    if (pl.sched_bytes && be_malloc(&pl.sched, pl.sched_bytes) != 0) { set_error("scheduler scratch allocation failed"); return NRB_ERR_OOM; }
    return NRB_OK;
}

// pointer tables of the fused exchange (PassParams::out_peer / in_peer): [0..8) for the pushed (low-z) part of the blocks,
// [8..16) for the pulled (high-z) part; zthr > zmask = no split (everything through the first half)
struct PeerExchange {
    double2 *out[16];
    const double2 *in[16];
    bool out_on, in_on;
    unsigned zmask, zthr;
    PeerExchange() : out{}, in{}, out_on(false), in_on(false), zmask(0), zthr(1) {}
};

void release_side_lane(SideLane &sl)
{
    if (sl.ev_fork) be_event_destroy(sl.ev_fork);
    if (sl.ev_join) be_event_destroy(sl.ev_join);
    if (sl.stream) be_stream_destroy(sl.stream);
    sl = SideLane();
}

static bool open_side_lane(SideLane &sl)
{
    if (sl.stream) return true;
    if (be_stream_create_prio(&sl.stream, 0) != 0) { sl.stream = nullptr; return false; }
    sl.ev_fork = be_event_create();
    sl.ev_join = be_event_create();
    if (!sl.ev_fork || !sl.ev_join) { release_side_lane(sl); return false; }
    return true;
}

static int run_program(Program &prog, double2 *const base[4], int arg, void *main_stream, std::vector<void *> *events = nullptr,
                       const PeerExchange *px = nullptr, void *sched = nullptr, SideLane *side = nullptr)
{
    void *stream = main_stream;
    if (events) events->push_back(be_event_record(stream));
    bool forked = false;
    // lanes are only honoured for plain executions (profiling times one launch after the other on one stream)
    const bool lanes = side && !events;
    for (Step &st : prog.steps) {
        int rc;
        if (lanes) {
            const bool touches_speq = st.speq.id != BUF_NONE || (st.is_fused && st.speq2.id != BUF_NONE);
            if (st.lane == 1 && (forked || open_side_lane(*side))) {
                if (!forked || st.side_after_main) {   // fork: everything enqueued so far precedes the side work
                    if (be_event_record_on(side->ev_fork, main_stream) != 0 || be_stream_wait(side->stream, side->ev_fork) != 0) {
                        set_error(std::string("side lane fork failed: ") + be_last_error());
                        return NRB_ERR_CUDA;
                    }
                    forked = true;
                }
                stream = side->stream;
            } else {
                if (forked && (touches_speq || st.join_side)) {   // join before the speq plane (or a chunk the side lane produced) is used on the main lane again
                    if (be_event_record_on(side->ev_join, side->stream) != 0 || be_stream_wait(main_stream, side->ev_join) != 0) {
                        set_error(std::string("side lane join failed: ") + be_last_error());
                        return NRB_ERR_CUDA;
                    }
                    forked = false;
                }
                stream = main_stream;
            }
        }
        if (st.is_fused) {
            PassParams pa = st.pp, pb = st.pp2;
            pa.in = base[st.in.id] + st.in.off; pa.out = base[st.out.id] + st.out.off;
            pa.speq = st.speq.id == BUF_NONE ? nullptr : base[st.speq.id] + st.speq.off;
            pb.in = base[st.in2.id] + st.in2.off; pb.out = base[st.out2.id] + st.out2.off;
            pb.speq = st.speq2.id == BUF_NONE ? nullptr : base[st.speq2.id] + st.speq2.off;
            FuseSched fs = st.fs;
            fs.ticket = (unsigned long long *)((char *)sched + st.sched_off);
            fs.done = (unsigned *)((char *)sched + st.sched_off + 16);
            rc = be_launch_fused(st.key, pa, st.key2, pb, fs, stream);
        } else if (st.trig == 1) {
            TrigParams tp = st.tp;
            tp.io = reinterpret_cast<double *>(base[st.in.id] + st.in.off);
            rc = be_launch_trig(st.key.log2n, tp, stream);
        } else if (st.trig == 2) {
            TwoFFTParams fp = st.fp;
            fp.d1 = reinterpret_cast<const double *>(base[st.in.id] + st.in.off);
            fp.d2 = reinterpret_cast<const double *>(base[st.b.id] + st.b.off);
            fp.f1 = base[st.out.id] + st.out.off;
            fp.f2 = base[st.speq.id] + st.speq.off;
            rc = be_launch_twofft(st.key.log2n, fp, stream);
        } else if (st.is_mid) {
            ConvMidParams mp = st.mp;
            mp.prefetch_dist = tunables().prefetch_dist >= 0 ? tunables().prefetch_dist : tunables().mid_prefetch;
            mp.data = base[st.in.id] + st.in.off;
            mp.b = st.b.id == BUF_NONE ? nullptr : base[st.b.id] + st.b.off;
            rc = be_launch_conv_mid(st.key.log2n, mp, st.ntiles, stream);
        } else if (st.is_aux) {
            AuxParams ap = st.ap;
            ap.a = st.in.id == BUF_NONE ? nullptr : base[st.in.id] + st.in.off;
            ap.b = st.b.id == BUF_NONE ? nullptr : base[st.b.id] + st.b.off;
            ap.out = st.out.id == BUF_NONE ? nullptr : base[st.out.id] + st.out.off;
            ap.speq = st.speq.id == BUF_NONE ? nullptr : base[st.speq.id] + st.speq.off;
            if (st.patch_pad_mode) ap.op = arg;
            rc = be_launch_aux(ap, stream);
        } else {
            PassParams pp = st.pp;
            pp.in = base[st.in.id] + st.in.off;
            pp.out = base[st.out.id] + st.out.off;
            // L2 prefetch of the tile one wave ahead: pays where an SM holds too few CTAs to overlap its own loads
            // with another CTA's shared-memory stages (ROW lines >= 4096: +3 ... +7 %) and for COL N = 128 (+3 %);
            // it costs 3 - 6 % on the other COL kernels (profiles/r01_tuning.md #26).  -1 = this policy.
            pp.prefetch_dist = tunables().prefetch_dist >= 0 ? tunables().prefetch_dist
                             : ((st.key.layout == LAYOUT_ROW && st.key.log2n >= 12) || (st.key.layout == LAYOUT_COL && st.key.log2n == 7 && st.key.variant == VAR_PLAIN)) ? 148 : 0;
            pp.speq = st.speq.id == BUF_NONE ? nullptr : base[st.speq.id] + st.speq.off;
            if (px && px->out_on && st.out.id == BUF_OUT) {   // exchange output: the peers' receive buffers / the local send buffer
                pp.out_peer_on = 1;
                pp.out_peer_off = 0;
                for (int i = 0; i < 16; ++i) pp.out_peer[i] = px->out[i] ? px->out[i] + st.out.off : nullptr;
                pp.peer_zmask = st.zsplit ? px->zmask : 0u;
                pp.peer_zthr = st.zsplit ? px->zthr : 1u;
            }
            if (px && px->in_on && st.in.id == BUF_OUT) {     // exchange input: the local receive buffer / the peers' send buffers
                pp.in_peer_on = 1;
                for (int i = 0; i < 16; ++i) pp.in_peer[i] = px->in[i] ? px->in[i] + st.in.off : nullptr;
                pp.peer_zmask = st.zsplit ? px->zmask : 0u;
                pp.peer_zthr = st.zsplit ? px->zthr : 1u;
                pp.prefetch_dist = 0;
            }
            rc = be_launch_pass(st.key, pp, st.ntiles, stream);
        }
        if (rc != 0) { set_error(std::string("kernel launch failed: ") + be_last_error()); return NRB_ERR_CUDA; }
        if (events) events->push_back(be_event_record(stream));
    }
    if (forked) {   // the program ends with side work in flight: later work on the caller's stream must see it
        if (be_event_record_on(side->ev_join, side->stream) != 0 || be_stream_wait(main_stream, side->ev_join) != 0) {
            set_error(std::string("side lane join failed: ") + be_last_error());
            return NRB_ERR_CUDA;
        }
    }
    return NRB_OK;
}

// stage 0 of the DMA exchange: like run_program, but block p of the exchange output lives at table[p] + p*BLK
static int run_program_dma(Program &prog, double2 *const base[4], void *stream, double2 *const table[8], i64 BLK)
{
    for (Step &st : prog.steps) {
        if (st.is_aux || st.is_fused) { set_error("slab: unexpected step in an exchange program"); return NRB_ERR_UNSUPPORTED; }
        PassParams pp = st.pp;
        pp.in = base[st.in.id] + st.in.off;
        pp.out = base[st.out.id] + st.out.off;
        pp.speq = st.speq.id == BUF_NONE ? nullptr : base[st.speq.id] + st.speq.off;
        pp.prefetch_dist = 0;
        if (st.out.id == BUF_OUT) {
            pp.out_peer_on = 1;
            pp.out_peer_off = 0;
            for (int i = 0; i < 8; ++i) pp.out_peer[i] = table[i] ? table[i] + (i64)i * BLK + st.out.off : nullptr;
        }
        if (be_launch_pass(st.key, pp, st.ntiles, stream) != 0) { set_error(std::string("kernel launch failed: ") + be_last_error()); return NRB_ERR_CUDA; }
    }
    return NRB_OK;
}

int exec_plan(Plan &pl, double *d_io, double *d_aux, double *d_out, int isign, int arg, void *stream)
{
    if (isign != 1 && isign != -1) { set_error("isign must be 1 or -1"); return NRB_ERR_INVALID_ISIGN; }
    double2 *const base[4] = {(double2 *)d_io, (double2 *)d_aux, (double2 *)d_out, (double2 *)pl.ws};
    Program &prog = pl.prog[isign == 1 ? 0 : 1];
    bool any_side = false;
    for (const Step &st : prog.steps) any_side = any_side || st.lane != 0;
    return run_program(prog, base, arg, stream, nullptr, nullptr, pl.sched, any_side ? &pl.side : nullptr);
}

int profile_plan(Plan &pl, double *d_io, double *d_aux, double *d_out, int isign, int arg, void *stream, float *ms,
                 int cap)
{
    if (isign != 1 && isign != -1) { set_error("isign must be 1 or -1"); return NRB_ERR_INVALID_ISIGN; }
    double2 *const base[4] = {(double2 *)d_io, (double2 *)d_aux, (double2 *)d_out, (double2 *)pl.ws};
    std::vector<void *> ev;
    const int rc = run_program(pl.prog[isign == 1 ? 0 : 1], base, arg, stream, &ev, nullptr, pl.sched);
    if (be_sync(stream) != 0 && rc == NRB_OK) { set_error(std::string("kernel execution failed: ") + be_last_error()); return NRB_ERR_CUDA; }
    for (size_t i = 0; i + 1 < ev.size(); ++i)
        if ((int)i < cap && ev[i] && ev[i + 1]) ms[i] = be_event_elapsed_ms(ev[i], ev[i + 1]);
    for (void *e : ev) if (e) be_event_destroy(e);
    return rc;
}

int describe_launch(const Plan &pl, int isign, int idx, char *name, size_t cap, double *bytes)
{
    const Program &prog = pl.prog[isign == 1 ? 0 : 1];
    if (idx < 0 || (size_t)idx >= prog.steps.size()) return NRB_ERR_INVALID_DIMS;
    const Step &st = prog.steps[(size_t)idx];
    char buf[128];
    double b = 0.0;
    if (st.is_fused) {
        snprintf(buf, sizeof(buf), "fused_%s_n%d+%s_n%d_%s_U%u", st.key.layout == LAYOUT_ROW ? "row_real" : "col", 1 << st.key.log2n,
                 st.key2.layout == LAYOUT_ROW ? "row_real" : "col", 1 << st.key2.log2n, st.key.dir > 0 ? "p" : "m", st.fs.units);
        // the pair reads the volume once and writes it once (the intermediate stays in L2)
        const double vol = (double)st.fs.units * (double)st.fs.ta * (double)(1 << tile_log2(st.key.log2n, st.key.layout));
        b = 2.0 * 16.0 * vol + 16.0 * (double)(st.key.layout == LAYOUT_ROW ? st.pp.q_end - st.pp.q_begin : st.pp2.q_end - st.pp2.q_begin);
    } else if (st.trig == 1) {
        static const char *tn[] = {"cosft1", "cosft2", "sinft", "cosft2_inv", "?"};
        const double n = (double)(2 << st.key.log2n);
        snprintf(buf, sizeof(buf), "trig_%s_n%d_L%llu", tn[st.tp.mode >= 0 && st.tp.mode <= 3 ? st.tp.mode : 4], 2 << st.key.log2n, (unsigned long long)st.tp.count);
        b = 2.0 * 8.0 * (double)st.tp.count * (st.tp.mode == COS1 ? n + 1.0 : n);      // every real read once and written once
    } else if (st.trig == 2) {
        const double n = (double)(1 << st.key.log2n);
        snprintf(buf, sizeof(buf), "trig_twofft_n%d_L%llu", 1 << st.key.log2n, (unsigned long long)st.fp.count);
        b = (double)st.fp.count * (16.0 * n + 32.0 * (n + 1.0));                         // two real lines in, two spectra out
    } else if (st.is_mid) {
        snprintf(buf, sizeof(buf), "conv_mid_n%d_f%d_op%d", 1 << st.key.log2n, st.mp.f, st.mp.op);
        const double pts = (double)st.mp.count * (double)((u64)1 << (st.key.log2n + st.mp.f));
        // the signals' intermediate is read and written once; the second operand is read once (per signal, or once
        // for the whole batch when it is shared)
        b = 2.0 * 16.0 * pts + (st.mp.op == SPEC_AUTOCORREL ? 0.0 : 16.0 * (st.mp.b_stride ? pts : pts / (double)st.mp.count));
    } else if (st.is_aux) {
        static const char *names[AUX_KIND_COUNT] = {"untangle", "spectral", "pad_response", "correl_direct", "fill", "spectral_z",
                                                    "signal", "wait", "reduce", "stats_final", "normalize", "power", "pack2",
                                                    "twofft_split", "scale", "cosft", "scan", "cmul", "spectral_zt"};
        snprintf(buf, sizeof(buf), "aux_%s", st.ap.kind >= 0 && st.ap.kind < AUX_KIND_COUNT ? names[st.ap.kind] : "unknown");
        switch (st.ap.kind) {
        case AUX_UNTANGLE: b = 2.0 * 16.0 * (double)st.ap.count * (double)st.ap.n; break;
        case AUX_SPECTRAL: b = (double)st.ap.count * (double)st.ap.n * 8.0 * (st.ap.b_stride ? 3.0 : 2.0) + (st.ap.b_stride ? 0.0 : 8.0 * (double)st.ap.n); break;
        case AUX_PAD_RESPONSE: b = 8.0 * ((double)st.ap.n + (double)st.ap.m); break;
        case AUX_SPECTRAL_ZT:
        case AUX_SPECTRAL_Z: b = (double)st.ap.count * (double)st.ap.n * 8.0 * (st.ap.b_stride ? 3.0 : 2.0) + (st.ap.b_stride ? 0.0 : 8.0 * (double)st.ap.n); break;
        case AUX_REDUCE: b = (double)st.ap.count * ((st.ap.op == RED_PARTIALS ? 16.0 : 8.0) * (double)st.ap.n + 16.0 * (double)st.ap.m); break;
        case AUX_STATS_FINAL: b = (double)st.ap.count * 16.0 * ((double)st.ap.m + 1.0); break;
        case AUX_NORMALIZE: b = 2.0 * 8.0 * (double)st.ap.count * (double)st.ap.n; break;
        case AUX_POWER: b = 24.0 * (double)st.ap.n; break;
        case AUX_PACK2: b = 32.0 * (double)st.ap.count * (double)st.ap.n; break;
        case AUX_TWOFFT_SPLIT: b = 48.0 * (double)st.ap.count * (double)st.ap.n; break;
        case AUX_SCALE: b = 16.0 * (double)st.ap.n; break;
        case AUX_COSFT: b = 16.0 * (double)st.ap.count * (double)st.ap.n; break;
        case AUX_SCAN: b = st.ap.op == 1 ? 32.0 * (double)st.ap.count * (double)((st.ap.n / 2 + kScanChunk - 1) / kScanChunk)
                                         : (st.ap.op == 0 ? 8.0 : 16.0) * (double)st.ap.count * (double)st.ap.n; break;
        default: b = 3.0 * 8.0 * (double)st.ap.count * (double)st.ap.n; break;
        }
    } else {
        static const char *var[] = {"plain", "real", "xpose"};
        const double lines = (double)(st.pp.q_end - st.pp.q_begin);
        snprintf(buf, sizeof(buf), "fft_%s_%s_n%d_%s_L%llu", st.key.layout == LAYOUT_ROW ? "row" : "col", pass_takes_tma_xpose(st.key, st.pp) ? "xpose_tma" : pass_takes_tma_in(st.key, st.pp) ? "tma_in" : pass_takes_tma(st.key, st.pp) ? "tma" : var[st.key.variant],
                 1 << st.key.log2n, st.key.dir > 0 ? "p" : "m", (unsigned long long)(st.pp.q_end - st.pp.q_begin));
        b = 2.0 * 16.0 * lines * (double)(1 << st.key.log2n);
        if (st.key.variant == VAR_REAL && st.pp.real_mode == REAL_SPEQ) b += 16.0 * lines;
    }
    if (name && cap) { strncpy(name, buf, cap - 1); name[cap - 1] = 0; }
    if (bytes) *bytes = b;
    return NRB_OK;
}

int complex_multiply_device(double *d_a, const double *d_b, u64 ncomplex, int conj_b, double scale, void *stream)
{
    AuxParams ap;
    memset(&ap, 0, sizeof(ap));
    ap.kind = AUX_CMUL;
    ap.op = conj_b ? 1 : 0;
    ap.a = (const double2 *)d_a; ap.b = (const double2 *)d_b; ap.out = (double2 *)d_a;
    ap.n = ncomplex;
    memcpy(&ap.m, &scale, sizeof(double));
    if (be_launch_aux(ap, stream) != 0) { set_error(std::string("kernel launch failed: ") + be_last_error()); return NRB_ERR_CUDA; }
    return NRB_OK;
}

int fill_uniform_device(double *d_out, u64 seed, u64 offset, u64 count, void *stream)
{
    AuxParams ap;
    memset(&ap, 0, sizeof(ap));
    ap.kind = AUX_FILL;
    ap.out = (double2 *)d_out;
    ap.n = count; ap.m = seed; ap.count = offset;
    if (be_launch_aux(ap, stream) != 0) { set_error(std::string("kernel launch failed: ") + be_last_error()); return NRB_ERR_CUDA; }
    return NRB_OK;
}

// ------------------------------------------------------------------ slab-decomposed rlft3 / 3-D complex fourn
// (fourn: the same programs with a plain complex z pass, N3 = nn3 complex points and no speq part.)
// Rank r of G.  X = nn1/G, Y = nn2/G, N3 = nn3/2, BLK = X*Y*(N3 + 1) complex per exchange
// block (data part X*Y*N3 followed by the speq part X*Y).
//   forward stage 0 : slab [nn1][Y][nn3] real -> z real pass (speq -> ws [nn1][Y]) ->
//                     x pass writing block p = x / X of `send` (data + speq parts)
//   exchange        : all-to-all of BLK-sized blocks (caller; NCCL over NVLink)
//   forward stage 1 : y pass reading the G received blocks (element y -> block y / Y),
//                     writing the nn1-slab [X][nn2][N3] complex and speq [X][nn2]
//   inverse         : mirror image (stage 0 = y pass into blocks, stage 1 = x pass + z c2r).
// buffers: BUF_IO = slab, BUF_AUX = local speq, BUF_OUT = send (stage 0) / recv (stage 1).
int build_slab_plan(SlabPlan &sp, size_t nn1, size_t nn2, size_t nn3, int nranks, int rank, bool real)
{
    if (!is_pow2(nn1) || !is_pow2(nn2) || !is_pow2(nn3) || nn3 < 2) { set_error("slab: dims must be powers of two"); return NRB_ERR_NOT_POW2; }
    if (nranks < 1 || !is_pow2((size_t)nranks) || (size_t)nranks > nn1 || (size_t)nranks > nn2 || rank < 0 || rank >= nranks) {
        set_error("slab: nranks must be a power of two dividing nn1 and nn2");
        return NRB_ERR_INVALID_DIMS;
    }
    sp.nn1 = nn1; sp.nn2 = nn2; sp.nn3 = nn3; sp.nranks = nranks; sp.rank = rank; sp.real = real;
    TableScope tables(&sp.tables);
    const u64 G = (u64)nranks, X = nn1 / G, Y = nn2 / G, N3 = sp.n3c();
    const int p1 = ilog2(nn1), p2 = ilog2(nn2), p3 = ilog2((size_t)N3);
    if (p1 > tunables().col_max_log2 || p2 > tunables().col_max_log2 || p3 > tunables().row_max_log2) {
        set_error("slab: axis too long for a single pass");
        return NRB_ERR_UNSUPPORTED;
    }
    const i64 BLK = (i64)sp.blk();
    const i64 SPQ = (i64)(X * Y * N3);      // offset of the speq part inside a block (rlft3 only)
    const BufRef SLAB(BUF_IO, 0), SPEQ(BUF_AUX, 0), XCH(BUF_OUT, 0), WSPEQ(BUF_WS, 0);
    sp.ws_elems = real ? (size_t)(nn1 * Y) : 0;   // speq of the nn2-slab: [nn1][Y]
    int rc = NRB_OK;

    AxisMap x_blocks;   // x index -> block x / X ; data part, lines (y, z)
    x_blocks.on = true; x_blocks.s0 = 0; x_blocks.es = (i64)(Y * N3); x_blocks.eshift = ilog2((size_t)X); x_blocks.es_hi = BLK;
    AxisMap x_blocks_speq = x_blocks;   // speq part, lines (y)
    x_blocks_speq.es = (i64)Y;
    AxisMap y_blocks;   // y index -> block y / Y ; data part, lines (xl, z)
    y_blocks.on = true; y_blocks.s0 = (i64)(Y * N3); y_blocks.es = (i64)N3; y_blocks.eshift = ilog2((size_t)Y); y_blocks.es_hi = BLK;
    AxisMap y_blocks_speq;   // speq part: lines xl, contiguous yl
    y_blocks_speq.on = true; y_blocks_speq.s0 = (i64)Y; y_blocks_speq.es = 1; y_blocks_speq.eshift = ilog2((size_t)Y); y_blocks_speq.es_hi = BLK;

    // z lines of the local slab: real transform with the Nyquist plane in WSPEQ (rlft3) or a plain complex pass (fourn)
    auto emit_z = [&](Builder &B, int dir) {
        if (real) emit_real(B, SLAB, SLAB, BufRef(), 0, nn1 * Y, p3, dir, REAL_SPEQ, WSPEQ);
        else emit_axis(B, SLAB, SLAB, BufRef(), nn1 * Y, 0, nn1 * Y, p3, 1, dir);
    };
    // speq_side: the speq-plane pass of a stage goes first in program order, on lane 1 (it only depends on what
    // precedes the stage, or on the z pass), so it runs beside the stage's data pass; run_program joins the lanes
    // before the z pass of the inverse reads the plane and at the end of every stage (before the exchange barrier)
    const bool side = real && tunables().speq_side != 0;
    auto lane1 = [&](Builder &B, size_t first) { if (side) for (size_t i = first; i < B.prog->steps.size(); ++i) B.prog->steps[i].lane = 1; };
    // z_chunks = 2: the z pass and the exchange pass beside it are cut into two halves of the local y rows; the z pass of
    // the second half (forward) / first half (inverse) runs on the side lane under the NVLink-bound exchange pass of the
    // other half instead of in front of (behind) the whole exchange.  A z launch covers the tiles of its y rows only
    // (PassParams::tile_run: lines are (x, y), so the rows of a half are every other run of Y/2 lines).
    const u64 Lz = (u64)lines_per_tile(p3, LAYOUT_ROW);
    const int ZC = (tunables().z_chunks == 2 && G > 1 && Y >= 2 * Lz && (Y / 2) % Lz == 0) ? 2 : 1;
    const u64 Yc = Y / (u64)ZC;
    auto z_chunk = [&](Builder &B, int dir, int c) {
        emit_z(B, dir);
        Step &st = B.prog->steps.back();
        st.pp.tile_run = ilog2((size_t)(Yc / Lz));
        st.pp.tile_nsel = 1;
        st.pp.tile_sel = c;
        st.ntiles /= 2;
    };
    auto x_chunk = [&](Builder &B, int dir, int c) {
        if (dir > 0) emit_axis(B, SLAB, XCH, BufRef(), 1, 0, 1, p1, Y * N3, +1, nullptr, &x_blocks);
        else emit_axis(B, XCH, SLAB, BufRef(), 1, 0, 1, p1, Y * N3, -1, &x_blocks, nullptr);
        Step &st = B.prog->steps.back();
        st.zsplit = true;
        st.pp.q_begin = (u64)c * Yc * N3;
        st.pp.q_end = (u64)(c + 1) * Yc * N3;
        st.ntiles = tiles_for(st.key.log2n, st.key.layout, st.pp.q_end - st.pp.q_begin);
    };
    if (ZC == 2) {   // forward stage 0: z0 | x0 beside (z1, speq x pass) | x1
        Builder B(&sp.prog[0][0]);
        z_chunk(B, +1, 0);
        z_chunk(B, +1, 1);
        B.prog->steps.back().lane = 1;
        if (real) { emit_axis(B, WSPEQ, XCH + SPQ, BufRef(), 1, 0, 1, p1, Y, +1, nullptr, &x_blocks_speq); B.prog->steps.back().lane = 1; }
        x_chunk(B, +1, 0);
        x_chunk(B, +1, 1);
        B.prog->steps.back().join_side = true;
        rc = B.rc ? B.rc : rc;
    } else
    {   // forward stage 0
        Builder B(&sp.prog[0][0]);
        emit_z(B, +1);
        if (side) { const size_t f0 = B.prog->steps.size(); emit_axis(B, WSPEQ, XCH + SPQ, BufRef(), 1, 0, 1, p1, Y, +1, nullptr, &x_blocks_speq); lane1(B, f0); }
        emit_axis(B, SLAB, XCH, BufRef(), 1, 0, 1, p1, Y * N3, +1, nullptr, &x_blocks);
        B.prog->steps.back().zsplit = true;
        if (real && !side) emit_axis(B, WSPEQ, XCH + SPQ, BufRef(), 1, 0, 1, p1, Y, +1, nullptr, &x_blocks_speq);
        rc = B.rc ? B.rc : rc;
    }
    {   // forward stage 1
        Builder B(&sp.prog[0][1]);
        if (side) { emit_axis(B, XCH + SPQ, SPEQ, BufRef(), X, 0, X, p2, 1, +1, &y_blocks_speq, nullptr); lane1(B, 0); }
        emit_axis(B, XCH, SLAB, BufRef(), X, 0, X, p2, N3, +1, &y_blocks, nullptr);
        B.prog->steps.back().zsplit = true;
        if (real && !side) emit_axis(B, XCH + SPQ, SPEQ, BufRef(), X, 0, X, p2, 1, +1, &y_blocks_speq, nullptr);
        rc = B.rc ? B.rc : rc;
    }
    {   // inverse stage 0
        Builder B(&sp.prog[1][0]);
        if (side) { emit_axis(B, SPEQ, XCH + SPQ, BufRef(), X, 0, X, p2, 1, -1, nullptr, &y_blocks_speq); lane1(B, 0); }
        emit_axis(B, SLAB, XCH, BufRef(), X, 0, X, p2, N3, -1, nullptr, &y_blocks);
        B.prog->steps.back().zsplit = true;
        if (real && !side) emit_axis(B, SPEQ, XCH + SPQ, BufRef(), X, 0, X, p2, 1, -1, nullptr, &y_blocks_speq);
        rc = B.rc ? B.rc : rc;
    }
    if (ZC == 2) {   // inverse stage 1: speq x pass on the side lane | x0 | x1 beside z0 | z1
        Builder B(&sp.prog[1][1]);
        if (real) { emit_axis(B, XCH + SPQ, WSPEQ, BufRef(), 1, 0, 1, p1, Y, -1, &x_blocks_speq, nullptr); B.prog->steps.back().lane = 1; }
        x_chunk(B, -1, 0);
        z_chunk(B, -1, 0);
        B.prog->steps.back().lane = 1;
        B.prog->steps.back().side_after_main = true;      // behind x0
        x_chunk(B, -1, 1);
        z_chunk(B, -1, 1);
        B.prog->steps.back().join_side = true;            // (it reads the speq plane, which the side lane produced)
        rc = B.rc ? B.rc : rc;
    } else
    {   // inverse stage 1
        Builder B(&sp.prog[1][1]);
        if (real) { emit_axis(B, XCH + SPQ, WSPEQ, BufRef(), 1, 0, 1, p1, Y, -1, &x_blocks_speq, nullptr); lane1(B, 0); }
        emit_axis(B, XCH, SLAB, BufRef(), 1, 0, 1, p1, Y * N3, -1, &x_blocks, nullptr);
        B.prog->steps.back().zsplit = true;
        emit_z(B, -1);
        rc = B.rc ? B.rc : rc;
    }
    if (rc != NRB_OK) { set_error("slab: shape not supported"); return rc; }
    sp.ws = nullptr;
    if (sp.ws_elems && be_malloc(&sp.ws, sp.ws_elems * sizeof(double2)) != 0) { set_error("slab: workspace allocation failed"); return NRB_ERR_OOM; }
    return NRB_OK;
}

// ---- pipelined exchange -------------------------------------------------------------------------------
// Lines of the x and y passes are (y or x_local, z).  Ordering them (z-chunk, y or x_local, z within chunk)
// makes a z-chunk a contiguous line range, so stage 0 can publish chunk c (all ranks' stores for that
// z-range have landed) while it works on chunk c+1, and stage 1 of chunk c runs on a second stream under the
// NVLink-bound stores of chunk c+1.  Because stage 1 then overlaps stage 0, the two must not share a buffer:
// forward  z pass: slab -> work;  x pass (c): work -> peers;  y pass (c): recv -> slab
// inverse  y pass (c): slab -> peers;  x pass (c): recv -> work;  z pass: work -> slab
// The speq planes travel with chunk 0.  `work` sits in the plan's workspace after the [nn1][Y] speq scratch.
namespace {
struct LineMap { i64 s0, s1; };     // strides of (z-chunk, middle index); the z index inside a chunk has stride 1
void chunk_lines(Step &st, u64 mid, LineMap in, LineMap out, u64 Zc, int c)
{
    PassParams &pp = st.pp;
    pp.logB = ilog2((size_t)Zc);
    pp.logA = ilog2((size_t)mid);
    pp.in_s0 = in.s0; pp.in_s1 = in.s1; pp.in_s2 = 1;
    pp.out_s0 = out.s0; pp.out_s1 = out.s1; pp.out_s2 = 1;
    pp.q_begin = (u64)c * mid * Zc;
    pp.q_end = (u64)(c + 1) * mid * Zc;
    st.ntiles = tiles_for(st.key.log2n, st.key.layout, pp.q_end - pp.q_begin);
}

// dma = false: blocks [X][Y][N3] (peer stores from the FFT epilogue); dma = true: blocks [chunk][X][Y][Zc], so the
// piece (peer, chunk) is one contiguous range for a copy engine
int build_chunk_programs(SlabPlan &sp, int chunks, bool dma)
{
    const u64 G = (u64)sp.nranks, X = sp.nn1 / G, Y = sp.nn2 / G, N3 = sp.n3c();
    const int p1 = ilog2(sp.nn1), p2 = ilog2(sp.nn2), p3 = ilog2((size_t)N3);
    const bool real = sp.real;
    TableScope tables(&sp.tables);
    const u64 Zc = N3 / (u64)chunks;
    if (Zc < 1 || Zc * (u64)chunks != N3) { set_error("slab: too many chunks for this nn3"); return NRB_ERR_INVALID_DIMS; }
    const i64 BLK = (i64)sp.blk(), SPQ = (i64)(X * Y * N3), PIECE = (i64)(X * Y * Zc);
    const size_t speq_ws = real ? (size_t)(sp.nn1 * Y) : 0, work_elems = (size_t)(sp.nn1 * Y * N3);
    if (sp.ws_elems < speq_ws + work_elems) {
        void *nw = nullptr;
        if (be_malloc(&nw, (speq_ws + work_elems) * sizeof(double2)) != 0) { set_error("slab: work buffer allocation failed"); return NRB_ERR_OOM; }
        if (sp.ws) be_free(sp.ws);
        sp.ws = nw;
        sp.ws_elems = speq_ws + work_elems;
    }
    const BufRef SLAB(BUF_IO, 0), SPEQ(BUF_AUX, 0), XCH(BUF_OUT, 0), WSPEQ(BUF_WS, 0), WORK(BUF_WS, (i64)speq_ws);
    // strides inside an exchange block
    const i64 xl_stride = dma ? (i64)(Y * Zc) : (i64)(Y * N3), yl_stride = dma ? (i64)Zc : (i64)N3, zc_stride = dma ? PIECE : (i64)Zc;
    AxisMap x_blocks;
    x_blocks.on = true; x_blocks.s0 = 0; x_blocks.es = xl_stride; x_blocks.eshift = ilog2((size_t)X); x_blocks.es_hi = BLK;
    AxisMap x_blocks_speq = x_blocks;
    x_blocks_speq.es = (i64)Y;
    AxisMap y_blocks;
    y_blocks.on = true; y_blocks.s0 = xl_stride; y_blocks.es = yl_stride; y_blocks.eshift = ilog2((size_t)Y); y_blocks.es_hi = BLK;
    AxisMap y_blocks_speq;
    y_blocks_speq.on = true; y_blocks_speq.s0 = (i64)Y; y_blocks_speq.es = 1; y_blocks_speq.eshift = ilog2((size_t)Y); y_blocks_speq.es_hi = BLK;
    const LineMap work_lines{(i64)Zc, (i64)N3};                      // [nn1][Y][N3]: lines (zc, y, zl)
    const LineMap slab_lines{(i64)Zc, (i64)(sp.nn2 * N3)};           // [X][nn2][N3]: lines (zc, xl, zl)
    const LineMap xch_x_lines{zc_stride, yl_stride};                 // exchange block, lines (zc, y_local, zl)
    const LineMap xch_y_lines{zc_stride, xl_stride};                 // exchange block, lines (zc, x_local, zl)
    int rc = NRB_OK;
    for (int s = 0; s < 2; ++s) {
        const int dir = s == 0 ? +1 : -1;
        sp.part[s].clear();
        sp.part[s].resize((size_t)(2 * chunks + 2));
        {   // before the chunks
            Builder B(&sp.part[s][0]);
            if (dir > 0 && real) emit_real(B, SLAB, WORK, BufRef(), 0, sp.nn1 * Y, p3, +1, REAL_SPEQ, WSPEQ);
            else if (dir > 0) emit_axis(B, SLAB, WORK, BufRef(), sp.nn1 * Y, 0, sp.nn1 * Y, p3, 1, +1);
            rc = B.rc ? B.rc : rc;
        }
        for (int c = 0; c < chunks; ++c) {
            Builder B0(&sp.part[s][(size_t)(1 + 2 * c)]), B1(&sp.part[s][(size_t)(2 + 2 * c)]);
            if (dir > 0) {
                emit_axis(B0, WORK, XCH, BufRef(), 1, 0, 1, p1, Y * N3, +1, nullptr, &x_blocks);
                chunk_lines(B0.prog->steps.back(), Y, work_lines, xch_x_lines, Zc, c);
                if (c == 0 && real) emit_axis(B0, WSPEQ, XCH + SPQ, BufRef(), 1, 0, 1, p1, Y, +1, nullptr, &x_blocks_speq);
                emit_axis(B1, XCH, SLAB, BufRef(), X, 0, X, p2, N3, +1, &y_blocks, nullptr);
                chunk_lines(B1.prog->steps.back(), X, xch_y_lines, slab_lines, Zc, c);
                if (c == 0 && real) emit_axis(B1, XCH + SPQ, SPEQ, BufRef(), X, 0, X, p2, 1, +1, &y_blocks_speq, nullptr);
            } else {
                emit_axis(B0, SLAB, XCH, BufRef(), X, 0, X, p2, N3, -1, nullptr, &y_blocks);
                chunk_lines(B0.prog->steps.back(), X, slab_lines, xch_y_lines, Zc, c);
                if (c == 0 && real) emit_axis(B0, SPEQ, XCH + SPQ, BufRef(), X, 0, X, p2, 1, -1, nullptr, &y_blocks_speq);
                if (c == 0 && real) emit_axis(B1, XCH + SPQ, WSPEQ, BufRef(), 1, 0, 1, p1, Y, -1, &x_blocks_speq, nullptr);
                emit_axis(B1, XCH, WORK, BufRef(), 1, 0, 1, p1, Y * N3, -1, &x_blocks, nullptr);
                chunk_lines(B1.prog->steps.back(), Y, xch_x_lines, work_lines, Zc, c);
            }
            rc = B0.rc ? B0.rc : (B1.rc ? B1.rc : rc);
        }
        {   // after the chunks
            Builder B(&sp.part[s][(size_t)(2 * chunks + 1)]);
            if (dir < 0 && real) emit_real(B, WORK, SLAB, BufRef(), 0, sp.nn1 * Y, p3, -1, REAL_SPEQ, WSPEQ);
            else if (dir < 0) emit_axis(B, WORK, SLAB, BufRef(), sp.nn1 * Y, 0, sp.nn1 * Y, p3, 1, -1);
            rc = B.rc ? B.rc : rc;
        }
    }
    if (rc != NRB_OK) { for (int s = 0; s < 2; ++s) sp.part[s].clear(); set_error("slab: shape not supported"); return rc; }
    return NRB_OK;
}
} // namespace

int slab_set_chunks(SlabPlan &sp, int chunks)
{
    if (chunks < 1 || !is_pow2((size_t)chunks)) { set_error("slab: chunks must be a power of two"); return NRB_ERR_INVALID_DIMS; }
    for (int s = 0; s < 2; ++s) sp.part[s].clear();
    sp.chunks = 1;
    sp.dma = false;
    if (chunks == 1) return NRB_OK;
    const int rc = build_chunk_programs(sp, chunks, false);
    if (rc == NRB_OK) sp.chunks = chunks;
    return rc;
}

int slab_set_dma(SlabPlan &sp, int chunks)
{
    if (chunks < 1 || chunks > kSlabMaxChunks || !is_pow2((size_t)chunks)) { set_error("slab: chunks must be a power of two <= 16"); return NRB_ERR_INVALID_DIMS; }
    for (int s = 0; s < 2; ++s) sp.part[s].clear();
    sp.chunks = 1;
    sp.dma = false;
    const int rc = build_chunk_programs(sp, chunks, true);
    if (rc != NRB_OK) return rc;
    const size_t xchg = sp.xchg_elems();
    if (!sp.send && be_malloc(&sp.send, xchg * sizeof(double2)) != 0) { set_error("slab: send buffer allocation failed"); return NRB_ERR_OOM; }
    if (!sp.copy_stream) {
        if (be_stream_create_prio(&sp.copy_stream, 0) != 0 || be_stream_create_prio(&sp.side_stream, 1) != 0) { set_error("slab: stream creation failed"); return NRB_ERR_CUDA; }
        sp.ev_go = be_event_create(); sp.ev_side = be_event_create(); sp.ev_copy = be_event_create();
        for (int k = 0; k < 3; ++k) {
            if (be_stream_create_prio(&sp.copy_extra[k], 0) != 0) { set_error("slab: stream creation failed"); return NRB_ERR_CUDA; }
            sp.ev_extra[k] = be_event_create();
        }
        for (int c = 0; c < kSlabMaxChunks; ++c) sp.ev_s0[c] = be_event_create();
    }
    sp.chunks = chunks;
    sp.dma = true;
    return NRB_OK;
}

void slab_release(SlabPlan &sp)
{
    release_side_lane(sp.side);
    if (sp.ws) be_free(sp.ws);
    if (sp.send) be_free(sp.send);
    if (sp.copy_stream) be_stream_destroy(sp.copy_stream);
    for (int k = 0; k < 3; ++k) {
        if (sp.copy_extra[k]) be_stream_destroy(sp.copy_extra[k]);
        if (sp.ev_extra[k]) be_event_destroy(sp.ev_extra[k]);
        sp.copy_extra[k] = sp.ev_extra[k] = nullptr;
    }
    if (sp.side_stream) be_stream_destroy(sp.side_stream);
    for (void *e : {sp.ev_go, sp.ev_side, sp.ev_copy}) if (e) be_event_destroy(e);
    for (int c = 0; c < kSlabMaxChunks; ++c) if (sp.ev_s0[c]) be_event_destroy(sp.ev_s0[c]);
    sp.ws = sp.send = sp.copy_stream = sp.side_stream = sp.ev_go = sp.ev_side = sp.ev_copy = nullptr;
    for (int c = 0; c < kSlabMaxChunks; ++c) sp.ev_s0[c] = nullptr;
}

int exec_slab_part(SlabPlan &sp, int stage, int part, int isign, double *d_slab, double *d_speq, void *stream, double *d_xchg)
{
    if (isign != 1 && isign != -1) { set_error("isign must be 1 or -1"); return NRB_ERR_INVALID_ISIGN; }
    if (sp.part[0].empty() || (!sp.dma && !sp.fused)) { set_error("slab: the pipelined exchange needs nrb_slab_set_peers and nrb_slab_set_chunks / nrb_slab_set_dma"); return NRB_ERR_INVALID_DIMS; }
    if (part < -1 || part > sp.chunks || ((part >= 0 && part < sp.chunks) && stage != 0 && stage != 1)) { set_error("slab: bad part"); return NRB_ERR_INVALID_DIMS; }
    const size_t idx = part < 0 ? 0 : part == sp.chunks ? (size_t)(2 * sp.chunks + 1) : (size_t)(1 + 2 * part + stage);
    const i64 BLK = (i64)sp.blk();
    const bool chunk_part = part >= 0 && part < sp.chunks;
    Program &prog = sp.part[isign == 1 ? 0 : 1][idx];
    if (sp.dma) {
        // stage 0 writes the caller's send buffer, stage 1 reads the caller's receive buffer
        if (chunk_part && !d_xchg) { set_error("slab: DMA exchange needs the send / receive buffer"); return NRB_ERR_INVALID_DIMS; }
        double2 *const base[4] = {(double2 *)d_slab, (double2 *)d_speq, (double2 *)d_xchg, (double2 *)sp.ws};
        for (Step &st : prog.steps) if (!st.is_aux) st.pp.grid_cap = 0;
        if (chunk_part && stage == 0 && sp.fused) {
            // blocks for the peers go to the send buffer; the own block goes straight to block `rank` of the own
            // receive buffer (no copy for it).  The pointer table is what the fused exchange uses for peer stores.
            double2 *table[8];
            for (int i = 0; i < 8; ++i) table[i] = i >= sp.nranks ? nullptr : i == sp.rank ? sp.peers[sp.rank] : (double2 *)d_xchg;
            return run_program_dma(prog, base, stream, table, BLK);
        }
        return run_program(prog, base, 0, stream);
    }
    double2 *const base[4] = {(double2 *)d_slab, (double2 *)d_speq, sp.peers[sp.rank], (double2 *)sp.ws};
    PeerExchange px;       // the pipelined exchange pushes everything
    px.out_on = true;
    for (int i = 0; i < 8; ++i) px.out[i] = px.out[8 + i] = sp.peers[i] ? sp.peers[i] + (i64)sp.rank * BLK : nullptr;
    const bool sends = chunk_part && stage == 0;
    for (Step &st : prog.steps) if (!st.is_aux) st.pp.grid_cap = sends ? tunables().xchg_grid_cap : 0;
    return run_program(prog, base, 0, stream, nullptr, sends ? &px : nullptr);
}

// One direction with the DMA-pipelined exchange.  Streams: `stream` runs the work before the chunks, every chunk's
// stage 0 and the work after the chunks; the copy stream pushes chunk c's pieces to the peers (copy engines, no SMs)
// and then publishes chunk c's epoch flag; the side stream waits for chunk c's flags of all ranks and runs stage 1.
int exec_slab_dma(SlabPlan &sp, int isign, double *d_slab, double *d_speq, unsigned long long epoch, void *stream)
{
    if (!sp.dma || !sp.fused) { set_error("slab: nrb_slab_set_dma and nrb_slab_set_peers first"); return NRB_ERR_INVALID_DIMS; }
    const u64 G = (u64)sp.nranks, X = sp.nn1 / G, Y = sp.nn2 / G, N3 = sp.n3c();
    const i64 BLK = (i64)sp.blk(), SPQ = (i64)(X * Y * N3), PIECE = (i64)(X * Y * (N3 / (u64)sp.chunks));
    double2 *send = (double2 *)sp.send, *recv = sp.peers[sp.rank];
    if (sp.timeline) { for (void *e : sp.tl_events) be_event_destroy(e); sp.tl_events.clear(); sp.tl_names.clear(); }
    auto mark = [&](const char *what, int c, void *s) {
        if (!sp.timeline) return;
        sp.tl_events.push_back(be_event_record(s));
        sp.tl_names.push_back(c >= 0 ? std::string(what) + "[" + std::to_string(c) + "]" : std::string(what));
    };
    mark("start", -1, stream);
    int rc = exec_slab_part(sp, 0, -1, isign, d_slab, d_speq, stream);
    if (rc) return rc;
    mark("pre", -1, stream);
    if (be_event_record_on(sp.ev_go, stream) || be_stream_wait(sp.copy_stream, sp.ev_go) || be_stream_wait(sp.side_stream, sp.ev_go)) {
        set_error(std::string("stream ordering failed: ") + be_last_error());
        return NRB_ERR_CUDA;
    }
    for (int c = 0; c < sp.chunks; ++c) {
        if ((rc = exec_slab_part(sp, 0, c, isign, d_slab, d_speq, stream, (double *)send))) return rc;
        mark("s0", c, stream);
        if (be_event_record_on(sp.ev_s0[c], stream) || be_stream_wait(sp.copy_stream, sp.ev_s0[c])) { set_error("stream ordering failed"); return NRB_ERR_CUDA; }
        int ns = tunables().dma_streams;
        if (ns > (int)G - 1) ns = (int)G - 1;
        if (ns < 1) ns = 1;
        for (int k = 1; k < ns; ++k)
            if (be_stream_wait(sp.copy_extra[k - 1], sp.ev_s0[c])) { set_error("stream ordering failed"); return NRB_ERR_CUDA; }
        for (u64 i = 1; i < G; ++i) {                     // rotated so that every rank targets a different peer at a time;
            const u64 p = ((u64)sp.rank + i) % G;         // the own block went straight into the own receive buffer
            const int k = (int)((i - 1) % (u64)ns);
            void *cs = k == 0 ? sp.copy_stream : sp.copy_extra[k - 1];
            if (be_d2d(sp.peers[p] + (i64)sp.rank * BLK + (i64)c * PIECE, send + (i64)p * BLK + (i64)c * PIECE, (size_t)PIECE * sizeof(double2), cs) != 0 ||
                (c == 0 && sp.real && be_d2d(sp.peers[p] + (i64)sp.rank * BLK + SPQ, send + (i64)p * BLK + SPQ, (size_t)(X * Y) * sizeof(double2), cs) != 0)) {
                set_error(std::string("peer copy failed: ") + be_last_error());
                return NRB_ERR_CUDA;
            }
        }
        for (int k = 1; k < ns; ++k)      // the flag of the chunk follows ALL its copies
            if (be_event_record_on(sp.ev_extra[k - 1], sp.copy_extra[k - 1]) || be_stream_wait(sp.copy_stream, sp.ev_extra[k - 1])) { set_error("stream ordering failed"); return NRB_ERR_CUDA; }
        mark("copy", c, sp.copy_stream);
        if ((rc = slab_barrier_chunk(sp, 0, c, epoch, sp.copy_stream))) return rc;     // chunk c of this rank has landed everywhere
        if ((rc = slab_barrier_chunk(sp, 1, c, epoch, sp.side_stream))) return rc;     // chunk c of every rank has landed here
        mark("wait", c, sp.side_stream);
        if ((rc = exec_slab_part(sp, 1, c, isign, d_slab, d_speq, sp.side_stream, (double *)recv))) return rc;
        mark("s1", c, sp.side_stream);
    }
    if (be_event_record_on(sp.ev_side, sp.side_stream) || be_stream_wait(stream, sp.ev_side) ||
        be_event_record_on(sp.ev_copy, sp.copy_stream) || be_stream_wait(stream, sp.ev_copy)) { set_error("stream ordering failed"); return NRB_ERR_CUDA; }
    rc = exec_slab_part(sp, 0, sp.chunks, isign, d_slab, d_speq, stream);
    mark("end", -1, stream);
    return rc;
}

std::string slab_dma_timeline(SlabPlan &sp)
{
    std::string out;
    char buf[64];
    for (size_t i = 0; i < sp.tl_events.size(); ++i) {
        snprintf(buf, sizeof(buf), "%s=%.3f ", sp.tl_names[i].c_str(), i ? be_event_elapsed_ms(sp.tl_events[0], sp.tl_events[i]) : 0.0f);
        out += buf;
    }
    return out;
}

int slab_set_peers(SlabPlan &sp, void *const *peer_recv, int count)
{
    if (!peer_recv) { sp.fused = false; return NRB_OK; }
    if (count != sp.nranks || count > 8) { set_error("slab: need one receive buffer per rank (at most 8)"); return NRB_ERR_INVALID_DIMS; }
    for (int i = 0; i < 8; ++i) sp.peers[i] = i < count ? (double2 *)peer_recv[i] : nullptr;
    for (int i = 0; i < count; ++i) if (!sp.peers[i]) { set_error("slab: null peer buffer"); return NRB_ERR_INVALID_DIMS; }
    sp.fused = true;
    return NRB_OK;
}

int slab_set_send_peers(SlabPlan &sp, void *const *peer_send, int count)
{
    for (int i = 0; i < 8; ++i) sp.sends[i] = nullptr;
    if (!peer_send) return NRB_OK;
    if (count != sp.nranks || count > 8) { set_error("slab: need one send buffer per rank (at most 8)"); return NRB_ERR_INVALID_DIMS; }
    for (int i = 0; i < count; ++i) if (!peer_send[i]) { set_error("slab: null peer buffer"); return NRB_ERR_INVALID_DIMS; }
    for (int i = 0; i < count; ++i) sp.sends[i] = (double2 *)peer_send[i];
    return NRB_OK;
}

int slab_barrier(SlabPlan &sp, int phase, unsigned long long epoch, void *stream)
{
    return slab_barrier_chunk(sp, phase, 0, epoch, stream);
}

int slab_barrier_chunk(SlabPlan &sp, int phase, int chunk, unsigned long long epoch, void *stream)
{
    if (chunk < 0 || chunk >= kSlabMaxChunks) { set_error("slab: bad chunk"); return NRB_ERR_INVALID_DIMS; }
    if (!sp.fused) { set_error("slab: the flag barrier needs the fused exchange (nrb_slab_set_peers)"); return NRB_ERR_INVALID_DIMS; }
    const u64 G = (u64)sp.nranks;
    const size_t xchg = sp.xchg_elems();   // complex elements
    AuxParams ap;
    memset(&ap, 0, sizeof(ap));
    ap.kind = phase == 1 ? AUX_WAIT : AUX_SIGNAL;      // phase 2: signal, then wait, in ONE launch
    ap.op = phase == 2 ? 1 : 0;
    ap.m = epoch;
    ap.n = (u64)sp.rank + (u64)chunk * G;       // one flag slot per (chunk, rank)
    ap.count = G;
    for (int i = 0; i < sp.nranks; ++i) ap.peer_flags[i] = (unsigned long long *)(sp.peers[i] + xchg);
    ap.out = (double2 *)((unsigned long long *)(sp.peers[sp.rank] + xchg) + (u64)chunk * G);
    if (phase == 1) ap.peer_flags[0] = (unsigned long long *)(sp.peers[sp.rank] + xchg) + (u64)chunk * G;
    if (be_launch_aux(ap, stream) != 0) { set_error(std::string("kernel launch failed: ") + be_last_error()); return NRB_ERR_CUDA; }
    return NRB_OK;
}

int exec_slab_stage(SlabPlan &sp, int stage, int isign, double *d_slab, double *d_speq, double *d_send,
                    double *d_recv, void *stream)
{
    if (isign != 1 && isign != -1) { set_error("isign must be 1 or -1"); return NRB_ERR_INVALID_ISIGN; }
    if (stage != 0 && stage != 1) { set_error("stage must be 0 or 1"); return NRB_ERR_INVALID_DIMS; }
    if (sp.fused) {
        // stage 0 stores go straight to the peers; stage 1 reads the local receive buffer
        // Block r -> p of the exchange (BLK elements): its low-z lines are PUSHED by rank r's stage 0 into block r of p's
        // receive buffer; its high-z lines (pull_eighths / 8 of the z range, when send buffers are set) are written by r
        // into block p of its own send buffer and PULLED from there by p's stage 1.  The NVLink traffic is thereby spread
        // over both passes of the step instead of being serialised in front of the second one.
        const i64 BLK = (i64)sp.blk();
        const u64 N3 = sp.n3c();
        double2 *const base[4] = {(double2 *)d_slab, (double2 *)d_speq, sp.peers[sp.rank], (double2 *)sp.ws};
        const bool pull = sp.sends[sp.rank] != nullptr && tunables().pull_eighths > 0 && N3 >= 64;
        PeerExchange px;
        px.zmask = (unsigned)(N3 - 1);
        px.zthr = pull ? (unsigned)(N3 - N3 / 8 * (u64)tunables().pull_eighths) : (unsigned)N3;
        for (int i = 0; i < sp.nranks; ++i) {
            if (stage == 0) {
                px.out[i] = sp.peers[i] + (i64)sp.rank * BLK;
                px.out[8 + i] = pull ? sp.sends[sp.rank] + (i64)i * BLK : px.out[i];
            } else {
                px.in[i] = sp.peers[sp.rank] + (i64)i * BLK;
                px.in[8 + i] = pull ? sp.sends[i] + (i64)sp.rank * BLK : px.in[i];
            }
        }
        px.out_on = stage == 0;
        px.in_on = stage == 1;
        return run_program(sp.prog[isign == 1 ? 0 : 1][stage], base, 0, stream, nullptr, &px, nullptr, &sp.side);
    }
    double2 *const base[4] = {(double2 *)d_slab, (double2 *)d_speq, (double2 *)(stage == 0 ? d_send : d_recv),
                              (double2 *)sp.ws};
    return run_program(sp.prog[isign == 1 ? 0 : 1][stage], base, 0, stream, nullptr, nullptr, nullptr, &sp.side);
}

int exec_slab_fused(SlabPlan &sp, int isign, double *d_slab, double *d_speq, unsigned long long epoch, void *stream)
{
    if (!sp.fused) { set_error("slab: nrb_slab_set_peers first"); return NRB_ERR_INVALID_DIMS; }
    int rc = exec_slab_stage(sp, 0, isign, d_slab, d_speq, nullptr, nullptr, stream);
    // one launch: publish my epoch to every peer (my stage-0 stores have landed, my send buffer is complete), then wait
    // until every peer's epoch has arrived here
    if (rc == NRB_OK && sp.nranks > 1) rc = slab_barrier(sp, 2, epoch, stream);
    if (rc == NRB_OK) rc = exec_slab_stage(sp, 1, isign, d_slab, d_speq, nullptr, nullptr, stream);
    return rc;
}

} // namespace nrb
