// extern "C" entry points of libbatman_b200.so (declared in include/batman_b200.h).
#include <math.h>
#include <stdlib.h>
#include <stdarg.h>

#include <new>
#include <vector>

#include "common.cuh"
#include "model.h"

namespace bk {

thread_local char g_err[512] = "";

int set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}
int cuda_fail(cudaError_t e, const char* what) {
    return set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
}

struct ProfEvent { cudaEvent_t a, b; int cat; };
static bool g_prof_on = false;
static std::vector<ProfEvent> g_prof_events;

ProfScope::ProfScope(int cat, cudaStream_t s) : st(s), idx(-1) {
    if (!g_prof_on) return;
    ProfEvent e; e.cat = cat;
    if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess) return;
    cudaEventRecord(e.a, st);
    g_prof_events.push_back(e);
    idx = (int)g_prof_events.size() - 1;
}
ProfScope::~ProfScope() {
    if (idx >= 0) cudaEventRecord(g_prof_events[idx].b, st);
}

// kernels (other translation units)
int launch_predict(const bk_gp_s* h, int q, const double* pts, long k, double* mean, double* kstar, int8_t* kq,
                   int kq_scheme, double kappa, cudaStream_t st);
int launch_var_ozaki(const bk_gp_s* h, int q, int scheme, const int8_t* kq, double kappa, long k, double* std_out,
                     double* scratch, int* fail_flag, cudaStream_t st);
int oz_gang_size(int n_pad, int sms);
int launch_var(const bk_gp_s* h, int q, const double* kstar, long k, double* std_out, cudaStream_t st);
int launch_pack_w(const double* w, int n, double* out, cudaStream_t st);
int launch_sobol(const bk_gp_s* h, int q, const double* a, const double* b, uint64_t seed, const int* mkind,
                 const double* mpar, long row_lo, long count, int second, double shift, double* partials,
                 double* yout, cudaStream_t st);
int launch_pod_inverse(const double* modes, long k, int m, const double* U, const double* mean, int F, double* out,
                       cudaStream_t st);
int launch_pod_sobol(const double* ymodes, long count, int p, int m, int second, const double* U, const double* mean,
                     int F, const double* shift, double* partials, cudaStream_t st);
int launch_design_rows(uint64_t seed, long row_lo, long count, int p, const int* mkind, const double* mpar,
                       int stream_id, double* out, cudaStream_t st);
int launch_sobol_y(const double* y, long ld_block, long count, int p, int F, int second, const double* shift,
                   double* partials, cudaStream_t st);
int launch_reduce_slots(const double* partials, int F, int nsums, double* totals, cudaStream_t st);
int launch_finalize(const double* totals, long N, int p, int F, int second, double* out, cudaStream_t st);
size_t sobol_result_len(int p, int F);
int launch_sobol_jackknife(const double* partials, const double* totals, long N, int p, int F, int second, double* out,
                           cudaStream_t st);
int launch_debug_fastmath(const double* in, long n, double* o_sqrt, double* o_exp, cudaStream_t st);
size_t lml_workspace_bytes(int n, int R, bool with_wt);
int lml_batched(const double* X, const double* y, int n_targets, const int* target, int n, int p, int R,
                const int* kinds, const double* c0, const double* c1, const double* kdiag, const double* ls,
                double* lml_out, void* ws, size_t ws_bytes, cudaStream_t st);
int gp_factorise(const double* X, const double* y, int n, int p, int kind, double c0, double c1, double kdiag,
                 const double* ls, double* L_out, double* alpha_out, double* wt_tiles_out, int8_t* wq_out,
                 double* wsig_out, int wq_scheme, double* lml_out, int* info_out, void* ws, size_t ws_bytes,
                 cudaStream_t st);
int lml_batched_general(const double* X, const double* y, int n_targets, const int* target, int n, int p, int R,
                        int nf, const int* fkind, int nt, const int* texp, const double* tcoef, const double* fls,
                        const double* kdiag, double* lml_out, void* ws, size_t ws_bytes, cudaStream_t st);
int gp_factorise_general(const double* X, const double* y, int n, int p, int nf, const int* fkind, int nt,
                         const int* texp, const double* tcoef, const double* fls, double kdiag, double* L_out,
                         double* alpha_out, double* wt_tiles_out, int8_t* wq_out, double* wsig_out, int wq_scheme,
                         double* lml_out, int* info_out, void* ws, size_t ws_bytes, cudaStream_t st);
int tri_inverse_pack(const double* L, int n, double* wt_tiles_out, int8_t* wq_out, double* wsig_out, int wq_scheme,
                     void* ws, size_t ws_bytes, cudaStream_t st);
int launch_lhs_rows(uint64_t seed, long count, int p, const int* mkind, const double* mpar, const long* perm,
                    double* out, cudaStream_t st);
int launch_moments(const double* y, long rows, int F, const double* shift, double* out, cudaStream_t st);
int launch_sobol_asym(const double* z, long ld_block, long count, int p, int r, const double* center, double* sums,
                      cudaStream_t st);
int sobol_nsums_host(int p, int second);
int sobol_nslots_host();

constexpr long PRED_CHUNK_ALIGN = 256;   // k_predict covers 2 tiles of 128 points per CTA

}  // namespace bk

using namespace bk;

static int check_marginals(const char* who, int p, const int* kind) {
    for (int c = 0; c < p; ++c)
        if (kind[c] < 0 || kind[c] >= BK_MARGINAL_COUNT) return set_error("%s: marginal %d has unknown kind %d", who, c, kind[c]);
    return 0;
}

extern "C" {

int bk_version(void) { return 100; }
const char* bk_last_error(void) { return g_err; }

int bk_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    BK_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    BK_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return 0;
}

int bk_profile_enable(int on) {
    g_prof_on = on != 0;
    return 0;
}

int bk_profile_read(double* ms_by_cat, long* launches_by_cat) {
    if (!ms_by_cat || !launches_by_cat) return set_error("bk_profile_read: NULL pointer");
    for (int c = 0; c < PROF_NCAT; ++c) { ms_by_cat[c] = 0.0; launches_by_cat[c] = 0; }
    for (auto& e : g_prof_events) {
        BK_CUDA(cudaEventSynchronize(e.b));
        float ms = 0.f;
        BK_CUDA(cudaEventElapsedTime(&ms, e.a, e.b));
        ms_by_cat[e.cat] += ms;
        launches_by_cat[e.cat] += 1;
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    g_prof_events.clear();
    return 0;
}

int bk_debug_fastmath(const double* in_dev, long n, double* sqrt_dev, double* exp_neg_dev, void* stream) {
    if (!in_dev || !sqrt_dev || !exp_neg_dev) return set_error("bk_debug_fastmath: NULL pointer");
    return launch_debug_fastmath(in_dev, n, sqrt_dev, exp_neg_dev, (cudaStream_t)stream);
}

int bk_gp_padded_dim(int p) { return padded_dim(p); }
int bk_gp_padded_n(int n) { return padded_n(n); }
size_t bk_gp_wtiles_len(int n) { return (size_t)padded_n(n) * padded_n(n); }

int bk_gp_create(bk_gp_t* out, int n, int p, int m) {
    if (!out) return set_error("bk_gp_create: out is NULL");
    if (n < 1 || p < 1 || m < 1) return set_error("bk_gp_create: n, p, m must be >= 1 (got %d, %d, %d)", n, p, m);
    const int P = padded_dim(p);
    if (P < 0) return set_error("bk_gp_create: input dimension %d > %d is not instantiated", p, MAX_P);
    bk_gp_s* h = new (std::nothrow) bk_gp_s();
    if (!h) return set_error("bk_gp_create: out of host memory");
    h->n = n; h->p = p; h->P = P; h->m = m; h->n_pad = padded_n(n);
    for (int c = 0; c < MAX_P; ++c) { h->scale[c] = 1.0; h->mn[c] = 0.0; }
    h->out.resize(m);
    *out = h;
    return 0;
}

int bk_gp_destroy(bk_gp_t h) {
    if (h && h->lanes) {
        PredictLanes* L = h->lanes;
        int cur = 0;
        const bool sw = cudaGetDevice(&cur) == cudaSuccess && cur != L->device && cudaSetDevice(L->device) == cudaSuccess;
        if (L->k1) { cudaStreamSynchronize(L->k1); cudaStreamDestroy(L->k1); }
        if (L->k2) { cudaStreamSynchronize(L->k2); cudaStreamDestroy(L->k2); }
        if (L->fork) cudaEventDestroy(L->fork);
        for (int i = 0; i < 2; ++i) {
            if (L->k1done[i]) cudaEventDestroy(L->k1done[i]);
            if (L->k2done[i]) cudaEventDestroy(L->k2done[i]);
        }
        if (sw) cudaSetDevice(cur);
        delete L;
    }
    delete h;
    return 0;
}

int bk_gp_set_output(bk_gp_t h, int q, int kind, double c0, double c1, double kdiag, const double* ls_host,
                     const double* xs_dev, const double* alpha_dev, const double* w_tiles_dev, double y_mean,
                     double y_std) {
    if (!h) return set_error("bk_gp_set_output: NULL handle");
    if (q < 0 || q >= h->m) return set_error("bk_gp_set_output: output %d out of range [0, %d)", q, h->m);
    if (kind < BK_MATERN12 || kind > BK_MATERN_INF) return set_error("bk_gp_set_output: unknown kernel kind %d", kind);
    if (!ls_host || !xs_dev || !alpha_dev) return set_error("bk_gp_set_output: NULL pointer");
    GpOutput& o = h->out[q];
    o.kind = kind; o.c0 = c0; o.c1 = c1; o.kdiag = kdiag; o.y_mean = y_mean; o.y_std = y_std;
    o.kmax = fabs(c0) + fabs(c1);            // stationary factor <= 1
    o.nonneg = c0 >= 0.0 && c1 >= 0.0;
    for (int c = 0; c < MAX_P; ++c) o.ls[c] = c < h->p ? ls_host[c] : 1.0;
    o.xs = xs_dev; o.alpha = alpha_dev; o.wt = w_tiles_dev;
    return 0;
}

// shared validation of a general tree description (set_output_general, lml_batched_general, factorise_general)
static int check_general(const char* who, int nf, const int* fkind, int nt, const int* texp, int P) {
    if (nf < 1 || nf > GEN_MAX_F) return set_error("%s: %d stationary leaves (supported: 1..%d)", who, nf, GEN_MAX_F);
    if (nt < 1 || nt > GEN_MAX_T) return set_error("%s: %d product terms (supported: 1..%d)", who, nt, GEN_MAX_T);
    if (nf * P > GEN_MAX_FP) return set_error("%s: %d leaves x padded dimension %d exceeds %d", who, nf, P, GEN_MAX_FP);
    for (int f = 0; f < nf; ++f)
        if (fkind[f] < BK_MATERN12 || fkind[f] > BK_MATERN_INF) return set_error("%s: leaf %d has unknown kind %d", who, f, fkind[f]);
    for (int i = 0; i < nt * nf; ++i)
        if (texp[i] < 0 || texp[i] > 4) return set_error("%s: exponent %d out of range [0, 4]", who, texp[i]);
    return 0;
}

int bk_gp_set_output_general(bk_gp_t h, int q, int nf, const int* fkind_host, const double* fls_host, int nt,
                             const double* tcoef_host, const int* texp_host, double kdiag, const double* xs_dev,
                             const double* alpha_dev, const double* w_tiles_dev, double y_mean, double y_std) {
    if (!h) return set_error("bk_gp_set_output_general: NULL handle");
    if (q < 0 || q >= h->m) return set_error("bk_gp_set_output_general: output %d out of range [0, %d)", q, h->m);
    if (!fkind_host || !fls_host || !tcoef_host || !texp_host || !xs_dev || !alpha_dev)
        return set_error("bk_gp_set_output_general: NULL pointer");
    if (int rc = check_general("bk_gp_set_output_general", nf, fkind_host, nt, texp_host, h->P)) return rc;
    GpOutput& o = h->out[q];
    o.kind = BK_GENERAL; o.c0 = 0.0; o.c1 = 0.0; o.kdiag = kdiag; o.y_mean = y_mean; o.y_std = y_std;
    o.gen = GenSpec();
    o.gen.nf = nf; o.gen.nt = nt;
    o.kmax = 0.0; o.nonneg = true;
    for (int f = 0; f < nf; ++f) {
        o.gen.fkind[f] = fkind_host[f];
        for (int c = 0; c < MAX_P; ++c) o.gls[f][c] = c < h->p ? fls_host[(size_t)f * h->p + c] : 1.0;
    }
    for (int t = 0; t < nt; ++t) {
        o.gen.tcoef[t] = tcoef_host[t];
        o.kmax += fabs(tcoef_host[t]);
        if (tcoef_host[t] < 0.0) o.nonneg = false;
        for (int f = 0; f < nf; ++f) o.gen.texp[t][f] = (unsigned char)texp_host[t * nf + f];
    }
    o.xs = xs_dev; o.alpha = alpha_dev; o.wt = w_tiles_dev;
    return 0;
}

int bk_gp_set_scaler(bk_gp_t h, const double* scale_host, const double* min_host) {
    if (!h) return set_error("bk_gp_set_scaler: NULL handle");
    for (int c = 0; c < h->p; ++c) {
        h->scale[c] = scale_host ? scale_host[c] : 1.0;
        h->mn[c] = min_host ? min_host[c] : 0.0;
    }
    return 0;
}

int bk_pack_w(const double* w_dev, int n, double* w_tiles_dev, void* stream) {
    if (!w_dev || !w_tiles_dev) return set_error("bk_pack_w: NULL pointer");
    return launch_pack_w(w_dev, n, w_tiles_dev, (cudaStream_t)stream);
}

static double pow2_at_least(double v) {
    int e = 0;
    if (!(v > 0.0)) return 1.0;
    frexp(v, &e);
    return ldexp(1.0, e);
}

// per point: K* tile bytes (FP64 fragments or 8 int8 slices: both n_pad * 8) + 304 bytes of sigma-kernel scratch (up to
// 37 partial sums per point, one per CTA of a gang, and an arrival counter per tile of 128 points)
static size_t predict_bytes_per_point(const bk_gp_s* h) { return (size_t)h->n_pad * sizeof(double) + 304; }

// A sigma prediction is a sequence of units (chunk of points, output): K1 writes the K* tiles of the unit, the sigma
// kernel consumes them.  With two K* buffers K1 of unit u+1 runs beside the sigma kernel of unit u.
size_t bk_gp_predict_workspace(bk_gp_t h, long k, int want_std) {
    if (!h || !want_std || k <= 0) return 0;
    long kk = (k + PRED_CHUNK_ALIGN - 1) / PRED_CHUNK_ALIGN * PRED_CHUNK_ALIGN;
    const long cap = 148L * TILE_T * 8;   // 8 test tiles per SM per chunk: 8 waves of the sigma kernel
    const bool several_units = kk > cap || h->m > 1;
    if (kk > cap) kk = cap;
    return (size_t)kk * predict_bytes_per_point(h) * (several_units ? 2 : 1) + 256;
}

static int lanes_get(bk_gp_s* h) {
    int dev = 0;
    BK_CUDA(cudaGetDevice(&dev));
    if (h->lanes && h->lanes->device == dev) return 0;
    if (h->lanes) return set_error("bk_gp_predict: handle was first used on device %d, now on %d", h->lanes->device, dev);
    PredictLanes* L = new (std::nothrow) PredictLanes();
    if (!L) return set_error("bk_gp_predict: out of host memory");
    h->lanes = L;                    // owned by the handle from here on (bk_gp_destroy releases partial state too)
    L->device = dev;
    int least = 0, greatest = 0;
    BK_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    // the sigma kernel needs a whole SM's shared memory per CTA: its CTAs go first whenever an SM frees up, K1 fills
    // the registers and shared memory left beside them
    BK_CUDA(cudaStreamCreateWithPriority(&L->k1, cudaStreamNonBlocking, least));
    BK_CUDA(cudaStreamCreateWithPriority(&L->k2, cudaStreamNonBlocking, greatest));
    BK_CUDA(cudaEventCreateWithFlags(&L->fork, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
        BK_CUDA(cudaEventCreateWithFlags(&L->k1done[i], cudaEventDisableTiming));
        BK_CUDA(cudaEventCreateWithFlags(&L->k2done[i], cudaEventDisableTiming));
    }
    return 0;
}

// K1 then the sigma kernel of one unit, each on the stream of its lane
static int predict_unit(bk_gp_s* h, int q, const double* pts, long cnt, double* mean, double* std_out, double* kstar,
                        double* scratch, int* fail_flag, cudaStream_t s1, cudaStream_t s2, cudaEvent_t between) {
    if (h->var_mode >= 1) {
        const GpOutput& o = h->out[q];
        const double kappa = pow2_at_least(o.kmax * 1.000001);   // |k| <= (|c0| + |c1| or sum |coef_t|)(1 + few eps) < kappa
        const int scheme = h->var_mode - 1;
        if (int rc = launch_predict(h, q, pts, cnt, mean, nullptr, (int8_t*)kstar, scheme, kappa, s1)) return rc;
        if (between) { BK_CUDA(cudaEventRecord(between, s1)); BK_CUDA(cudaStreamWaitEvent(s2, between, 0)); }
        return launch_var_ozaki(h, q, scheme, (const int8_t*)kstar, kappa, cnt, std_out, scratch, fail_flag, s2);
    }
    if (int rc = launch_predict(h, q, pts, cnt, mean, kstar, nullptr, 0, 1.0, s1)) return rc;
    if (between) { BK_CUDA(cudaEventRecord(between, s1)); BK_CUDA(cudaStreamWaitEvent(s2, between, 0)); }
    return launch_var(h, q, kstar, cnt, std_out, s2);
}

int bk_gp_set_output_int8(bk_gp_t h, int q, const int8_t* wq_dev, const double* wsig_dev) {
    if (!h) return set_error("bk_gp_set_output_int8: NULL handle");
    if (q < 0 || q >= h->m) return set_error("bk_gp_set_output_int8: output %d out of range [0, %d)", q, h->m);
    h->out[q].wq = wq_dev;
    h->out[q].wsig = wsig_dev;
    return 0;
}

int bk_sigma_gang_size(int n_train, int sm_count) {
    if (n_train < 1 || sm_count < 1) return 0;
    return oz_gang_size(padded_n(n_train), sm_count);
}

int bk_gp_set_overlap(bk_gp_t h, int on) {
    if (!h) return set_error("bk_gp_set_overlap: NULL handle");
    h->overlap = on != 0;
    return 0;
}

int bk_gp_set_var_mode(bk_gp_t h, int mode) {
    if (!h) return set_error("bk_gp_set_var_mode: NULL handle");
    if (mode < 0 || mode > 2)
        return set_error("bk_gp_set_var_mode: mode must be 0 (FP64 DMMA), 1 (int8, 8x7-bit slices) or 2 (int8, 7x8-bit slices)");
    if (mode == 1 && h->n_pad > 16384) return set_error("bk_gp_set_var_mode: 8x7-bit int32 accumulation is exact only for n_train <= 16384");
    if (mode == 2 && h->n_pad > 8192) return set_error("bk_gp_set_var_mode: 7x8-bit int32 accumulation is exact only for n_train <= 8192");
    if (mode == 2)
        for (int q = 0; q < h->m; ++q)
            if (!h->out[q].nonneg)
                return set_error("bk_gp_set_var_mode: the 7x8-bit scheme slices K* into unsigned bytes and needs non-negative coefficients");
    h->var_mode = mode;
    return 0;
}

size_t bk_gp_wq_bytes(int n_train) { return (size_t)padded_n(n_train) * padded_n(n_train) * 8; }

int bk_gp_predict(bk_gp_t h, const double* pts_dev, long k, double* mean_dev, double* std_dev, void* workspace_dev,
                  size_t workspace_bytes, void* stream) {
    if (!h) return set_error("bk_gp_predict: NULL handle");
    if (k < 0) return set_error("bk_gp_predict: negative point count");
    if (k == 0) return 0;
    if (!pts_dev || !mean_dev) return set_error("bk_gp_predict: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (!std_dev) {
        for (int q = 0; q < h->m; ++q)
            if (int rc = launch_predict(h, q, pts_dev, k, mean_dev, nullptr, nullptr, 0, 1.0, st)) return rc;
        return 0;
    }
    const size_t per_point = predict_bytes_per_point(h);
    const long fit = workspace_bytes > 256 ? (long)((workspace_bytes - 256) / per_point) : 0;   // points of K* the workspace holds
    if (!workspace_dev || fit < PRED_CHUNK_ALIGN)
        return set_error("bk_gp_predict: sigma needs a workspace of at least %zu bytes (got %zu)",
                         per_point * PRED_CHUNK_ALIGN + 256, workspace_bytes);
    char* ws = (char*)workspace_dev;
    int* fail_flag = (int*)(ws + ((workspace_bytes - 256) & ~(size_t)7));      // tail of the caller's workspace
    if (h->var_mode >= 1) BK_CUDA(cudaMemsetAsync(fail_flag, 0, sizeof(int), st));   // read back by bk_gp_predict_status
    // two K* buffers whenever the call has several units and the workspace holds two chunks
    const long k_al = (k + PRED_CHUNK_ALIGN - 1) / PRED_CHUNK_ALIGN * PRED_CHUNK_ALIGN;
    long chunk = fit / PRED_CHUNK_ALIGN * PRED_CHUNK_ALIGN;
    int nbuf = 1;
    if (h->overlap && (k_al > chunk || h->m > 1)) {
        const long half = fit / 2 / PRED_CHUNK_ALIGN * PRED_CHUNK_ALIGN;
        if (half >= PRED_CHUNK_ALIGN) { chunk = half; nbuf = 2; }
    }
    if (chunk > k_al) chunk = k_al;
    double* kbuf[2] = {(double*)ws, (double*)(ws + (size_t)chunk * per_point)};
    const size_t kdoubles = (size_t)chunk * h->n_pad;          // the sigma-kernel scratch follows the K* tiles of a buffer
    if (nbuf == 1) {
        for (long lo = 0; lo < k; lo += chunk) {
            const long cnt = (k - lo) < chunk ? (k - lo) : chunk;
            for (int q = 0; q < h->m; ++q)
                if (int rc = predict_unit(h, q, pts_dev + lo * h->p, cnt, mean_dev + lo * h->m, std_dev + lo * h->m,
                                          kbuf[0], kbuf[0] + kdoubles, fail_flag, st, st, nullptr)) return rc;
        }
        return 0;
    }
    if (int rc = lanes_get(h)) return rc;
    PredictLanes* L = h->lanes;
    BK_CUDA(cudaEventRecord(L->fork, st));
    BK_CUDA(cudaStreamWaitEvent(L->k1, L->fork, 0));
    BK_CUDA(cudaStreamWaitEvent(L->k2, L->fork, 0));
    long u = 0;
    for (long lo = 0; lo < k; lo += chunk) {
        const long cnt = (k - lo) < chunk ? (k - lo) : chunk;
        for (int q = 0; q < h->m; ++q, ++u) {
            const int b = (int)(u & 1);
            if (u >= 2) BK_CUDA(cudaStreamWaitEvent(L->k1, L->k2done[b], 0));   // the sigma kernel of unit u-2 released K* buffer b
            if (int rc = predict_unit(h, q, pts_dev + lo * h->p, cnt, mean_dev + lo * h->m, std_dev + lo * h->m,
                                      kbuf[b], kbuf[b] + kdoubles, fail_flag, L->k1, L->k2, L->k1done[b])) return rc;
            BK_CUDA(cudaEventRecord(L->k2done[b], L->k2));
        }
    }
    // join: the last sigma kernel follows every K1 (through k1done) and every earlier sigma kernel (stream order)
    BK_CUDA(cudaStreamWaitEvent(st, L->k2done[(int)((u - 1) & 1)], 0));
    return 0;
}

int bk_gp_predict_status(const void* workspace_dev, size_t workspace_bytes, int* fault_host, void* stream) {
    if (!workspace_dev || !fault_host) return set_error("bk_gp_predict_status: NULL pointer");
    if (workspace_bytes < 512) return set_error("bk_gp_predict_status: workspace of %zu bytes has no status tail", workspace_bytes);
    const char* ws = (const char*)workspace_dev;
    const int* fail_flag = (const int*)(ws + ((workspace_bytes - 256) & ~(size_t)7));
    cudaStream_t st = (cudaStream_t)stream;
    BK_CUDA(cudaMemcpyAsync(fault_host, fail_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    BK_CUDA(cudaStreamSynchronize(st));
    return 0;
}

size_t bk_lml_workspace(int n_train, int n_candidates, int with_inverse) {
    return lml_workspace_bytes(n_train, n_candidates, with_inverse != 0);
}

int bk_lml_batched(const double* x_dev, const double* y_dev, int n_targets, const int* target_host, int n_train, int p,
                   int n_candidates, const int* kind_host, const double* c0_host, const double* c1_host, const double* kdiag_host,
                   const double* ls_host, double* lml_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (!x_dev || !y_dev || !kind_host || !c0_host || !c1_host || !kdiag_host || !ls_host || !lml_dev || !workspace_dev)
        return set_error("bk_lml_batched: NULL pointer");
    if (n_train < 1 || p < 1 || n_candidates < 1) return set_error("bk_lml_batched: n, p, R must be >= 1");
    return lml_batched(x_dev, y_dev, n_targets, target_host, n_train, p, n_candidates, kind_host, c0_host, c1_host,
                       kdiag_host, ls_host, lml_dev, workspace_dev, workspace_bytes, (cudaStream_t)stream);
}

int bk_gp_factorise(const double* x_dev, const double* y_dev, int n_train, int p, int kind, double c0, double c1,
                    double kdiag, const double* ls_host, double* l_dev, double* alpha_dev, double* w_tiles_dev,
                    int8_t* wq_dev, double* wsig_dev, int wq_scheme, double* lml_dev, int* info_dev, void* workspace_dev,
                    size_t workspace_bytes, void* stream) {
    if (!x_dev || !y_dev || !ls_host || !l_dev || !alpha_dev || !lml_dev || !info_dev || !workspace_dev)
        return set_error("bk_gp_factorise: NULL pointer");
    return gp_factorise(x_dev, y_dev, n_train, p, kind, c0, c1, kdiag, ls_host, l_dev, alpha_dev, w_tiles_dev, wq_dev,
                        wsig_dev, wq_scheme, lml_dev, info_dev, workspace_dev, workspace_bytes, (cudaStream_t)stream);
}

int bk_lml_batched_general(const double* x_dev, const double* y_dev, int n_targets, const int* target_host, int n_train,
                           int p, int n_candidates, int nf,
                           const int* fkind_host, int nt, const int* texp_host, const double* tcoef_host,
                           const double* fls_host, const double* kdiag_host, double* lml_dev, void* workspace_dev,
                           size_t workspace_bytes, void* stream) {
    if (!x_dev || !y_dev || !fkind_host || !texp_host || !tcoef_host || !fls_host || !kdiag_host || !lml_dev || !workspace_dev)
        return set_error("bk_lml_batched_general: NULL pointer");
    if (n_train < 1 || p < 1 || n_candidates < 1) return set_error("bk_lml_batched_general: n, p, R must be >= 1");
    if (padded_dim(p) < 0) return set_error("bk_lml_batched_general: input dimension %d > %d", p, MAX_P);
    if (int rc = check_general("bk_lml_batched_general", nf, fkind_host, nt, texp_host, padded_dim(p))) return rc;
    return lml_batched_general(x_dev, y_dev, n_targets, target_host, n_train, p, n_candidates, nf, fkind_host, nt,
                               texp_host, tcoef_host, fls_host, kdiag_host, lml_dev, workspace_dev, workspace_bytes,
                               (cudaStream_t)stream);
}

int bk_gp_factorise_general(const double* x_dev, const double* y_dev, int n_train, int p, int nf,
                            const int* fkind_host, int nt, const int* texp_host, const double* tcoef_host,
                            const double* fls_host, double kdiag, double* l_dev, double* alpha_dev, double* w_tiles_dev,
                            int8_t* wq_dev, double* wsig_dev, int wq_scheme, double* lml_dev, int* info_dev,
                            void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (!x_dev || !y_dev || !fkind_host || !texp_host || !tcoef_host || !fls_host || !l_dev || !alpha_dev || !lml_dev ||
        !info_dev || !workspace_dev)
        return set_error("bk_gp_factorise_general: NULL pointer");
    if (padded_dim(p) < 0) return set_error("bk_gp_factorise_general: input dimension %d > %d", p, MAX_P);
    if (int rc = check_general("bk_gp_factorise_general", nf, fkind_host, nt, texp_host, padded_dim(p))) return rc;
    return gp_factorise_general(x_dev, y_dev, n_train, p, nf, fkind_host, nt, texp_host, tcoef_host, fls_host, kdiag,
                                l_dev, alpha_dev, w_tiles_dev, wq_dev, wsig_dev, wq_scheme, lml_dev, info_dev,
                                workspace_dev, workspace_bytes, (cudaStream_t)stream);
}

int bk_tri_inverse_pack(const double* l_dev, int n_train, double* w_tiles_dev, int8_t* wq_dev, double* wsig_dev,
                        int wq_scheme, void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (!l_dev || !w_tiles_dev || !workspace_dev) return set_error("bk_tri_inverse_pack: NULL pointer");
    return tri_inverse_pack(l_dev, n_train, w_tiles_dev, wq_dev, wsig_dev, wq_scheme, workspace_dev, workspace_bytes,
                            (cudaStream_t)stream);
}

int bk_sobol_nsums(int p, int second_order) { return sobol_nsums_host(p, second_order); }
int bk_sobol_nslots(void) { return sobol_nslots_host(); }

int bk_design_rows(uint64_t seed, long row_lo, long count, int p, const int* marg_kind_host,
                   const double* marg_par_host, int stream_id, double* out_dev, void* stream) {
    if (!marg_kind_host || !marg_par_host || !out_dev) return set_error("bk_design_rows: NULL pointer");
    if (int rc = check_marginals("bk_design_rows", p, marg_kind_host)) return rc;
    return launch_design_rows(seed, row_lo, count, p, marg_kind_host, marg_par_host, stream_id, out_dev,
                              (cudaStream_t)stream);
}

int bk_sobol_accumulate(bk_gp_t h, const double* a_dev, const double* b_dev, uint64_t seed, const int* marg_kind_host,
                        const double* marg_par_host, long row_lo, long count, int second_order,
                        const double* shift_host, double* partials_dev, void* stream) {
    if (!h || !partials_dev) return set_error("bk_sobol_accumulate: NULL pointer");
    if ((a_dev == nullptr) != (b_dev == nullptr)) return set_error("bk_sobol_accumulate: a_dev and b_dev must both be given or both NULL");
    if (!a_dev && (!marg_kind_host || !marg_par_host))
        return set_error("bk_sobol_accumulate: no base samples and no marginals to generate them from");
    if (!a_dev)
        if (int rc = check_marginals("bk_sobol_accumulate", h->p, marg_kind_host)) return rc;
    if (count <= 0) return 0;
    for (int q = 0; q < h->m; ++q)
        if (int rc = launch_sobol(h, q, a_dev, b_dev, seed, marg_kind_host, marg_par_host, row_lo, count,
                                  second_order ? 1 : 0, shift_host ? shift_host[q] : 0.0, partials_dev, nullptr,
                                  (cudaStream_t)stream))
            return rc;
    return 0;
}

int bk_sobol_accumulate_keep(bk_gp_t h, const double* a_dev, const double* b_dev, uint64_t seed,
                             const int* marg_kind_host, const double* marg_par_host, long row_lo, long count,
                             int second_order, const double* shift_host, double* partials_dev, double* ykeep_dev,
                             void* stream) {
    if (!h || !partials_dev || !ykeep_dev) return set_error("bk_sobol_accumulate_keep: NULL pointer");
    if ((a_dev == nullptr) != (b_dev == nullptr)) return set_error("bk_sobol_accumulate_keep: a_dev and b_dev must both be given or both NULL");
    if (!a_dev && (!marg_kind_host || !marg_par_host))
        return set_error("bk_sobol_accumulate_keep: no base samples and no marginals to generate them from");
    if (!a_dev)
        if (int rc = check_marginals("bk_sobol_accumulate_keep", h->p, marg_kind_host)) return rc;
    if (count <= 0) return 0;
    for (int q = 0; q < h->m; ++q)
        if (int rc = launch_sobol(h, q, a_dev, b_dev, seed, marg_kind_host, marg_par_host, row_lo, count,
                                  second_order ? 1 : 0, shift_host ? shift_host[q] : 0.0, partials_dev, ykeep_dev,
                                  (cudaStream_t)stream))
            return rc;
    return 0;
}

int bk_sobol_asymptotic(const double* z_dev, long ld_block, long count, int p, int r, const double* center_dev,
                        double* sums_dev, void* stream) {
    if (!z_dev || !center_dev || !sums_dev) return set_error("bk_sobol_asymptotic: NULL pointer");
    if (p < 1 || r < 1 || ld_block < count) return set_error("bk_sobol_asymptotic: need p >= 1, r >= 1, ld_block >= count");
    return launch_sobol_asym(z_dev, ld_block, count, p, r, center_dev, sums_dev, (cudaStream_t)stream);
}

int bk_lhs_rows(uint64_t seed, long count, int p, const int* marg_kind_host, const double* marg_par_host,
                const long* perm_dev, double* out_dev, void* stream) {
    if (!marg_kind_host || !marg_par_host || !perm_dev || !out_dev) return set_error("bk_lhs_rows: NULL pointer");
    if (p < 1) return set_error("bk_lhs_rows: p must be >= 1");
    if (int rc = check_marginals("bk_lhs_rows", p, marg_kind_host)) return rc;
    return launch_lhs_rows(seed, count, p, marg_kind_host, marg_par_host, perm_dev, out_dev, (cudaStream_t)stream);
}

int bk_moments(const double* y_dev, long rows, int F, const double* shift_dev, double* out_dev, void* stream) {
    if (!y_dev || !out_dev) return set_error("bk_moments: NULL pointer");
    if (F < 1) return set_error("bk_moments: F must be >= 1");
    return launch_moments(y_dev, rows, F, shift_dev, out_dev, (cudaStream_t)stream);
}

int bk_sobol_modes(bk_gp_t h, const double* a_dev, const double* b_dev, uint64_t seed, const int* marg_kind_host,
                   const double* marg_par_host, long row_lo, long count, int second_order, double* ymodes_dev,
                   void* stream) {
    if (!h || !ymodes_dev) return set_error("bk_sobol_modes: NULL pointer");
    if ((a_dev == nullptr) != (b_dev == nullptr)) return set_error("bk_sobol_modes: a_dev and b_dev must both be given or both NULL");
    if (!a_dev && (!marg_kind_host || !marg_par_host))
        return set_error("bk_sobol_modes: no base samples and no marginals to generate them from");
    if (!a_dev)
        if (int rc = check_marginals("bk_sobol_modes", h->p, marg_kind_host)) return rc;
    if (count <= 0) return 0;
    for (int q = 0; q < h->m; ++q)
        if (int rc = launch_sobol(h, q, a_dev, b_dev, seed, marg_kind_host, marg_par_host, row_lo, count,
                                  second_order ? 1 : 0, 0.0, nullptr, ymodes_dev, (cudaStream_t)stream))
            return rc;
    return 0;
}

int bk_pod_inverse(const double* modes_dev, long k, int m, const double* u_dev, const double* mean_dev, int F,
                   double* out_dev, void* stream) {
    if (!modes_dev || !u_dev || !mean_dev || !out_dev) return set_error("bk_pod_inverse: NULL pointer");
    if (m < 1 || F < 1) return set_error("bk_pod_inverse: m and F must be >= 1");
    return launch_pod_inverse(modes_dev, k, m, u_dev, mean_dev, F, out_dev, (cudaStream_t)stream);
}

int bk_pod_sobol_accumulate(const double* ymodes_dev, long count, int p, int m, int second_order,
                            const double* u_dev, const double* mean_dev, int F, const double* shift_dev,
                            double* partials_dev, void* stream) {
    if (!ymodes_dev || !u_dev || !mean_dev || !shift_dev || !partials_dev)
        return set_error("bk_pod_sobol_accumulate: NULL pointer");
    if (p < 1 || m < 1 || F < 1) return set_error("bk_pod_sobol_accumulate: p, m, F must be >= 1");
    return launch_pod_sobol(ymodes_dev, count, p, m, second_order ? 1 : 0, u_dev, mean_dev, F, shift_dev,
                            partials_dev, (cudaStream_t)stream);
}

int bk_sobol_accumulate_y(const double* y_dev, long ld_block, long count, int p, int F, int second_order,
                          const double* shift_dev, double* partials_dev, void* stream) {
    if (!y_dev || !shift_dev || !partials_dev) return set_error("bk_sobol_accumulate_y: NULL pointer");
    if (p < 1 || F < 1) return set_error("bk_sobol_accumulate_y: p and F must be >= 1");
    if (count <= 0) return 0;
    return launch_sobol_y(y_dev, ld_block, count, p, F, second_order ? 1 : 0, shift_dev, partials_dev,
                          (cudaStream_t)stream);
}

int bk_sobol_reduce_slots(const double* partials_dev, int F, int nsums, double* totals_dev, void* stream) {
    if (!partials_dev || !totals_dev) return set_error("bk_sobol_reduce_slots: NULL pointer");
    return launch_reduce_slots(partials_dev, F, nsums, totals_dev, (cudaStream_t)stream);
}

int bk_sobol_jackknife(const double* partials_dev, const double* totals_dev, long N, int p, int F, int second_order,
                       double* out_dev, void* stream) {
    if (!partials_dev || !totals_dev || !out_dev) return set_error("bk_sobol_jackknife: NULL pointer");
    if (N < 3 || p < 1 || F < 1) return set_error("bk_sobol_jackknife: need N >= 3, p >= 1, F >= 1");
    return launch_sobol_jackknife(partials_dev, totals_dev, N, p, F, second_order ? 1 : 0, out_dev, (cudaStream_t)stream);
}

size_t bk_sobol_result_len(int p, int F) { return sobol_result_len(p, F); }

int bk_sobol_finalize(const double* totals_dev, long N, int p, int F, int second_order, double* out_dev,
                      void* stream) {
    if (!totals_dev || !out_dev) return set_error("bk_sobol_finalize: NULL pointer");
    if (N < 2) return set_error("bk_sobol_finalize: N must be >= 2");
    return launch_finalize(totals_dev, N, p, F, second_order ? 1 : 0, out_dev, (cudaStream_t)stream);
}

}  // extern "C"
