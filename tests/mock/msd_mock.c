/* msd_mock.c -- TEST DOUBLE of libmosdepth_b200.so for CPU tests of the host CLI stand-in (tests/test_cli_on_mock.py).
 *
 * NOT a product path and NOT a CPU fallback: it is built only by that test, into a temporary directory, and linked only into a
 * temporary copy of the CLI.  It answers the C-ABI calls mosdepth_b200/host/main.cpp makes (include/mosdepth_b200.h) from the CPU
 * oracle's own functions (oracle/mosdepth_oracle.c is included below as source; tests may link the oracle), so that everything the
 * HOST does -- flags, header and index parsing, spans, BED tables, the per-contig loop, row formatting, the BGZF/CSI writer --
 * can be compared byte for byte with the oracle's output files without a GPU.  The kernels are tested on the GPU through the
 * real library (tests/test_gpu_*.py); nothing here says anything about them.
 *
 * Contigs are computed lazily and in order (the CLI asks msd_contig_records(ti) first, then the queries of ti), one depth array
 * at a time, exactly like the oracle's own main loop. */
#define main oracle_main_unused
#include "../../oracle/mosdepth_oracle.c"
#undef main
#include "../../include/mosdepth_b200.h"

struct msd_ctx { int dummy; };
static struct msd_ctx g_ctx;
static char g_err[512] = "";
static char g_path[4096];
static bam_t g_bam; static int g_open = 0, g_cur = -1, g_cur_tid = -2, g_nt = 0;
static cov_opts g_co; static seen_t g_seen; static int g_seen_ready = 0;
static uint8_t *g_want = NULL;
static int32_t *g_rs = NULL, *g_re = NULL, *g_rv = NULL; static int64_t g_rcap = 0;

static int fail(int code, const char *msg) { snprintf(g_err, sizeof g_err, "%s", msg); return code; }

int msd_abi_version(void) { return MSD_ABI_VERSION; }
int msd_create(int device, msd_ctx **out) { (void)device; *out = &g_ctx; return MSD_OK; }
void msd_destroy(msd_ctx *ctx) { (void)ctx; }
const char *msd_last_error(const msd_ctx *ctx) { (void)ctx; return g_err; }
int msd_open(msd_ctx *ctx, const char *bam_path, int n_targets, const uint32_t *target_len, uint64_t first_record_voff) {
    (void)ctx; (void)first_record_voff;
    snprintf(g_path, sizeof g_path, "%s", bam_path);
    bam_open(&g_bam, g_path, 0); g_open = 1; g_nt = g_bam.n_targets;
    if (n_targets != g_bam.n_targets) return fail(MSD_EINVAL, "mock: the host's header has another number of targets");
    for (int i = 0; i < n_targets; i++) if (target_len[i] != g_bam.targets[i].length) return fail(MSD_EINVAL, "mock: the host's header has other target lengths");
    free(g_want); g_want = (uint8_t *)xmalloc((size_t)n_targets); memset(g_want, 1, (size_t)n_targets);
    memset(&g_co, 0, sizeof g_co); g_co.max_len = INT64_MAX; g_co.eflag = 1796;
    return MSD_OK;
}
int msd_open_mem(msd_ctx *ctx, const void *b, uint64_t n, int nt, const uint32_t *tl, uint64_t v) { (void)ctx; (void)b; (void)n; (void)nt; (void)tl; (void)v; return fail(MSD_EINVAL, "mock: msd_open_mem is not provided"); }
int msd_stage(msd_ctx *ctx) { (void)ctx; return MSD_OK; }
int msd_set_spans(msd_ctx *ctx, int n, const uint64_t *b, const uint64_t *e) {
    (void)ctx;
    for (int i = 0; i < n; i++) if (e[i] && e[i] < b[i]) return fail(MSD_EINVAL, "mock: span ends before it starts");
    return MSD_OK;
}
int msd_set_targets(msd_ctx *ctx, const uint8_t *want) { (void)ctx; if (want) memcpy(g_want, want, (size_t)g_nt); else memset(g_want, 1, (size_t)g_nt); return MSD_OK; }
int msd_set_filters(msd_ctx *ctx, int mapq, int64_t min_len, int64_t max_len, uint16_t eflag, uint16_t iflag, const char *const *rgs, int n_rg, int mode) {
    (void)ctx;
    memset(&g_co, 0, sizeof g_co);
    g_co.mapq = mapq; g_co.min_len = min_len; g_co.max_len = max_len; g_co.eflag = eflag; g_co.iflag = iflag;
    g_co.fast_mode = mode == MSD_MODE_FAST; g_co.fragment_mode = mode == MSD_MODE_FRAGMENT;
    for (int i = 0; i < n_rg; i++) { g_co.read_groups = (char **)xrealloc(g_co.read_groups, (size_t)(g_co.n_rg + 1) * sizeof(char *)); g_co.read_groups[g_co.n_rg++] = strdup(rgs[i]); }
    return MSD_OK;
}
int msd_prefetch_windows(msd_ctx *ctx, uint32_t window) { (void)ctx; (void)window; return MSD_OK; }
int msd_run(msd_ctx *ctx) {
    (void)ctx;
    if (!g_open) return fail(MSD_EINVAL, "mock: msd_run before msd_open");
    bam_open(&g_bam, g_path, 0);                    /* a fresh sequential reader */
    if (!g_seen_ready) { seen_init(&g_seen); g_seen_ready = 1; }
    g_cur = -1; g_cur_tid = -2;
    return MSD_OK;
}
int msd_last_run_info(const msd_ctx *ctx, msd_run_info *out) { (void)ctx; memset(out, 0, sizeof *out); return MSD_OK; }

static int need(int tid, const char *who) {
    if (tid != g_cur || g_cur_tid == -2) { snprintf(g_err, sizeof g_err, "mock: %s(%d) but the current contig is %d (rc %d)", who, tid, g_cur, g_cur_tid); return MSD_EINVAL; }
    return 0;
}
int msd_contig_records(msd_ctx *ctx, int tid, int64_t *n) {
    (void)ctx;
    if (tid < 0 || tid >= g_nt || !g_want[tid]) return -1;
    if (tid < g_cur) return fail(MSD_EINVAL, "mock: contigs must be asked for in header order");
    if (tid != g_cur) { g_cur_tid = coverage(&g_bam, tid, &g_co, &g_seen); if (g_cur_tid != -2) to_coverage(); g_cur = tid; }
    if (n) *n = g_cur_tid == -2 ? 0 : 1;
    return g_cur_tid;
}
static int64_t hist_len(const int64_t *d, int64_t n) { while (n > 1 && d[n - 1] == 0) n--; return n; }
int msd_contig_stats(msd_ctx *ctx, int tid, msd_depth_stat *st, int64_t *dist, int64_t cap, int64_t *dist_len) {
    (void)ctx; int rc = need(tid, "msd_contig_stats"); if (rc) return rc;
    dist_t d; dist_new(&d, 512); dist_inc(&d, 0, (uint32_t)(g_arr_len - 1));
    depth_stat s = ds_new(g_arr, g_arr_len - 1);
    if (st) { st->cum_length = s.cum_length; st->cum_depth = s.cum_depth; st->min_depth = s.min_depth; st->max_depth = s.max_depth; }
    const int64_t n = hist_len(d.d, d.len);
    if (dist_len) *dist_len = n;
    if (dist) { if (cap < n) { free(d.d); return fail(MSD_EINVAL, "mock: dist_cap too small"); } memcpy(dist, d.d, (size_t)n * 8); }
    free(d.d);
    return MSD_OK;
}
int64_t msd_n_windows(uint32_t tlen, uint32_t window) { if (!window) return -1; return tlen ? ((int64_t)tlen + window - 1) / window : 0; }
static void stat_out(msd_depth_stat *o, depth_stat s) { if (o) { o->cum_length = s.cum_length; o->cum_depth = s.cum_depth; o->min_depth = s.min_depth; o->max_depth = s.max_depth; } }
int64_t msd_windows(msd_ctx *ctx, int tid, uint32_t window, int use_median, double *values, int64_t cap, msd_depth_stat *rst, int64_t *d512) {
    (void)ctx; int rc = need(tid, "msd_windows"); if (rc) return rc;
    const uint32_t tlen = g_bam.targets[tid].length, L = (uint32_t)g_arr_len;
    const int64_t nw = msd_n_windows(tlen, window);
    if (cap < nw) return fail(MSD_EINVAL, "mock: value_cap too small");
    count_stat cs; cs.len = use_median ? 65536 : 0; cs.n = 0; cs.counts = (uint32_t *)xcalloc((size_t)(cs.len ? cs.len : 1), 4);
    depth_stat acc; ds_clear(&acc);
    if (d512) memset(d512, 0, 512 * 8);
    for (int64_t k = 0; k < nw; k++) {
        const uint32_t s = (uint32_t)((uint64_t)k * window); uint64_t e64 = ((uint64_t)k + 1) * window; const uint32_t e = e64 < tlen ? (uint32_t)e64 : tlen;
        const double me = imean(s, e, &cs); values[k] = me;
        const uint32_t a = s < L ? s : L, b = L < e ? L : e;
        acc = ds_add(acc, ds_new(g_arr + a, b > a ? (int64_t)(b - a) : 0));
        if (d512) { const int64_t bin = nim_toInt(me); d512[bin < 511 ? bin : 511] += 1; }
    }
    free(cs.counts); stat_out(rst, acc);
    return nw;
}
int msd_regions(msd_ctx *ctx, int tid, int64_t n, const uint32_t *start, const uint32_t *stop, int use_median, double *values, msd_depth_stat *rst, int64_t *dist,
                int64_t cap, int64_t *dist_len, const int32_t *thr, int n_thr, int64_t *thr_counts) {
    (void)ctx; int rc = need(tid, "msd_regions"); if (rc) return rc;
    const uint32_t L = (uint32_t)g_arr_len;
    count_stat cs; cs.len = use_median ? 65536 : 0; cs.n = 0; cs.counts = (uint32_t *)xcalloc((size_t)(cs.len ? cs.len : 1), 4);
    depth_stat acc; ds_clear(&acc); dist_t d; dist_new(&d, 512);
    for (int64_t k = 0; k < n; k++) {
        if (values) values[k] = imean(start[k], stop[k], &cs);
        const uint32_t a = start[k] < L ? start[k] : L, b = L < stop[k] ? L : stop[k];
        acc = ds_add(acc, ds_new(g_arr + a, b > a ? (int64_t)(b - a) : 0));
        if (dist) dist_inc(&d, start[k], stop[k]);
        if (n_thr && thr_counts) {
            int64_t *c = thr_counts + k * n_thr; for (int j = 0; j < n_thr; j++) c[j] = 0;
            const int64_t e = stop[k] < (uint64_t)g_arr_len ? stop[k] : g_arr_len;
            for (int64_t p = start[k]; p < e; p++) { const int32_t v = g_arr[p]; for (int j = 0; j < n_thr; j++) { if (v < thr[j]) break; c[j]++; } }
        }
    }
    free(cs.counts); stat_out(rst, acc);
    const int64_t nd = hist_len(d.d, d.len);
    if (dist_len) *dist_len = nd;
    if (dist) { if (cap < nd) { free(d.d); return fail(MSD_EINVAL, "mock: dist_cap too small"); } memcpy(dist, d.d, (size_t)nd * 8); }
    free(d.d);
    return MSD_OK;
}
static void push_run(int64_t *n, int32_t s, int32_t e, int32_t v) {
    if (*n == g_rcap) { g_rcap = g_rcap ? 2 * g_rcap : 1024; g_rs = (int32_t *)xrealloc(g_rs, (size_t)g_rcap * 4); g_re = (int32_t *)xrealloc(g_re, (size_t)g_rcap * 4); g_rv = (int32_t *)xrealloc(g_rv, (size_t)g_rcap * 4); }
    g_rs[*n] = s; g_re[*n] = e; g_rv[*n] = v; (*n)++;
}
int msd_per_base_runs(msd_ctx *ctx, int tid, const int32_t **start, const int32_t **stop, const int32_t **depth, int64_t *n_runs) {
    (void)ctx; int rc = need(tid, "msd_per_base_runs"); if (rc) return rc;
    const int64_t tlen = g_arr_len - 1; int64_t n = 0, a = 0;
    for (int64_t i = 1; i <= tlen; i++) if (i == tlen || g_arr[i] != g_arr[a]) { push_run(&n, (int32_t)a, (int32_t)i, g_arr[a]); a = i; }
    *start = g_rs; *stop = g_re; *depth = g_rv; *n_runs = n;
    return MSD_OK;
}
/* msd_mock_bgzf.cpp: the host replay of K10's arithmetic */
int msd_mock_per_base_bgzf(msd_ctx *ctx, int tid, const char *chrom, const uint8_t **z, int64_t *nz, const msd_bgzf_member **m, int64_t *nm, const int32_t **s, const int32_t **e,
                           const int32_t **d, int64_t *n);
int msd_per_base_bgzf(msd_ctx *ctx, int tid, const char *chrom, const uint8_t **z, int64_t *nz, const msd_bgzf_member **m, int64_t *nm, const int32_t **s, const int32_t **e,
                      const int32_t **d, int64_t *n) {
    if (strlen(chrom) > 200) return fail(MSD_EINVAL, "mock: contig name too long for msd_per_base_bgzf");
    return msd_mock_per_base_bgzf(ctx, tid, chrom, z, nz, m, nm, s, e, d, n);
}
int msd_quantized_runs(msd_ctx *ctx, int tid, const int64_t *quants, int nq, const int32_t **start, const int32_t **stop, const int32_t **bin, int64_t *n_runs) {
    (void)ctx; int rc = need(tid, "msd_quantized_runs"); if (rc) return rc;
    int64_t n = 0;
    if (g_arr_len > 0) {      /* write_quantized above, yielding instead of printing */
        const int nl = nq - 1; int64_t last_q, q; linear_search(g_arr[0], quants, nq, &last_q);
        int64_t last_pos = 0; const int64_t high = g_arr_len - 1;
        for (int64_t pos = 0; pos < high - 1; pos++) {
            linear_search(g_arr[pos], quants, nq, &q);
            if (q == last_q) continue;
            if (last_q != -1 && last_q < nl) push_run(&n, (int32_t)last_pos, (int32_t)pos, (int32_t)last_q);
            last_q = q; last_pos = pos;
        }
        if (last_q != -1 && last_pos < high && last_q < nl) push_run(&n, (int32_t)last_pos, (int32_t)(g_arr_len - 1), (int32_t)last_q);
    }
    *start = g_rs; *stop = g_re; *bin = g_rv; *n_runs = n;
    return MSD_OK;
}

/* the rest of the ABI: present so that the ctypes binding (mosdepth_b200/__init__.py) loads against the double; not provided */
/* multi-GPU calls: the test double is one context on no GPU; the CLI only makes these calls with --gpus > 1 */
int msd_set_contig_cover(msd_ctx *ctx, int tid, uint32_t covered_from) { (void)ctx; (void)tid; (void)covered_from; return MSD_OK; }
int msd_contig_span_need(msd_ctx *ctx, int tid, uint32_t *a, uint32_t *b) { (void)ctx; (void)tid; if (a) *a = 0xffffffffu; if (b) *b = 0; return MSD_OK; }
int msd_comm_unique_id(void *id) { (void)id; return fail(MSD_ECOMM, "mock: no NCCL in the test double"); }
int msd_comm_init(msd_ctx *ctx, int rank, int world, const void *id) { (void)ctx; (void)rank; (void)world; (void)id; return fail(MSD_ECOMM, "mock: no NCCL in the test double"); }
int msd_comm_destroy(msd_ctx *ctx) { (void)ctx; return MSD_OK; }
int msd_exchange_spills(msd_ctx *ctx) { (void)ctx; return fail(MSD_ECOMM, "mock: no NCCL in the test double"); }
int msd_totals_allreduce(msd_ctx *ctx) { (void)ctx; return fail(MSD_ECOMM, "mock: no NCCL in the test double"); }
int msd_totals(msd_ctx *ctx, msd_depth_stat *g, msd_depth_stat *r, int64_t *gd, int64_t gc, int64_t *gl, int64_t *rd, int64_t rc, int64_t *rl) {
    (void)ctx; (void)g; (void)r; (void)gd; (void)gc; (void)gl; (void)rd; (void)rc; (void)rl; return fail(MSD_EINVAL, "mock: msd_totals is not provided");
}
int msd_set_contig_range(msd_ctx *c, int t, uint32_t lo, uint32_t hi) { (void)c; (void)t; (void)lo; (void)hi; return fail(MSD_EINVAL, "mock: not provided"); }
int msd_contig_spill(msd_ctx *c, int t, uint32_t *f, uint32_t *n) { (void)c; (void)t; (void)f; (void)n; return fail(MSD_EINVAL, "mock: not provided"); }
int msd_depth_copy(msd_ctx *c, int t, uint32_t p, uint32_t n, int32_t *d, int dev) { (void)c; (void)t; (void)p; (void)n; (void)d; (void)dev; return fail(MSD_EINVAL, "mock: not provided"); }
int msd_depth_add(msd_ctx *c, int t, uint32_t p, uint32_t n, const int32_t *s, int dev) { (void)c; (void)t; (void)p; (void)n; (void)s; (void)dev; return fail(MSD_EINVAL, "mock: not provided"); }
int msd_contig_finalize(msd_ctx *c, int t) { (void)c; (void)t; return fail(MSD_EINVAL, "mock: not provided"); }
int msd_totals_device(msd_ctx *c, void **g, void **r, int64_t *n, void **s) { (void)c; (void)g; (void)r; (void)n; (void)s; return fail(MSD_EINVAL, "mock: not provided"); }
int msd_totals_reset(msd_ctx *c) { (void)c; return MSD_OK; }
int msd_mark(msd_ctx *c, int slot) { (void)c; (void)slot; return MSD_OK; }
int msd_elapsed_ms(msd_ctx *c, int a, int b, double *ms) { (void)c; (void)a; (void)b; if (ms) *ms = 0.0; return MSD_OK; }
uint64_t msd_launch_count(const msd_ctx *c) { (void)c; return 0; }
int64_t msd_debug_inflate(msd_ctx *c, const void *z, uint64_t n, void *o, uint64_t cap) { (void)c; (void)z; (void)n; (void)o; (void)cap; return fail(MSD_EINVAL, "mock: not provided"); }
int msd_debug_depth(msd_ctx *c, int t, int32_t *o, int64_t cap) { (void)c; (void)t; (void)o; (void)cap; return fail(MSD_EINVAL, "mock: not provided"); }
int msd_debug_cumsum(msd_ctx *c, int32_t *a, int64_t n) { (void)c; (void)a; (void)n; return fail(MSD_EINVAL, "mock: not provided"); }
int64_t msd_debug_crc32(msd_ctx *c, const void *b, uint64_t n, uint32_t bb, uint32_t fo, int reps, uint32_t *out, double *ms) { (void)c; (void)b; (void)n; (void)bb; (void)fo; (void)reps; (void)out; (void)ms; return fail(MSD_EINVAL, "mock: not provided"); }
