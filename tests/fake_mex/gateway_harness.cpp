// gateway_harness.cpp -- runs pcreg_b200/csrc/pcreg_mex.cpp WITHOUT MATLAB and WITHOUT a GPU: a minimal in-memory
// implementation of the mx*/mex* calls the gateway uses (tests/fake_mex/mex.h) plus STUBS of the C ABI of
// include/pcreg.h that record what they were handed and return canned 0-based / column-major results.  What is tested is
// the marshaling only: argument checks, 1-based <-> 0-based indices, column-major layouts, struct field parsing,
// "the reference returns []" -> 0x0 double, error text -> mexErrMsgIdAndTxt.  TEST INFRASTRUCTURE (tests/test_mex_gateway.py).
//
// Build + run:  g++ -std=c++17 -I tests/fake_mex tests/fake_mex/gateway_harness.cpp pcreg_b200/csrc/pcreg_mex.cpp -o harness && ./harness
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "mex.h"
#include "../../include/pcreg.h"

// ---------------------------------------------------------------------------------------------------------
// fake MATLAB runtime
// ---------------------------------------------------------------------------------------------------------
struct mxArray_tag {
    mxClassID cls = mxDOUBLE_CLASS;
    size_t m = 0, n = 0;
    std::vector<unsigned char> data;
    std::map<std::string, mxArray*> fields;     // struct
    std::string str;                            // char
};
static size_t elem_size(mxClassID c) {
    switch (c) {
        case mxDOUBLE_CLASS: case mxINT64_CLASS: case mxUINT64_CLASS: return 8;
        case mxSINGLE_CLASS: case mxINT32_CLASS: case mxUINT32_CLASS: return 4;
        default: return 1;
    }
}
static int g_live = 0;                          // arrays created by the gateway and not destroyed / returned
static mxArray* new_array(mxClassID c, size_t m, size_t n) {
    mxArray* a = new mxArray_tag;
    a->cls = c; a->m = m; a->n = n;
    a->data.assign(m * n * elem_size(c), 0);    // MATLAB zero-fills
    ++g_live;
    return a;
}
struct FakeMexError : std::runtime_error { using std::runtime_error::runtime_error; };
static int g_locks = 0;
static void (*g_atexit)(void) = nullptr;

extern "C" {
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity) { return new_array(mxDOUBLE_CLASS, m, n); }
mxArray* mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity) { return new_array(cls, m, n); }
mxArray* mxCreateDoubleScalar(double v) { mxArray* a = new_array(mxDOUBLE_CLASS, 1, 1); *(double*)a->data.data() = v; return a; }
void* mxGetData(const mxArray* a) { return (void*)a->data.data(); }
double* mxGetPr(const mxArray* a) { return (double*)a->data.data(); }
double mxGetScalar(const mxArray* a) {
    switch (a->cls) {
        case mxDOUBLE_CLASS: return *(const double*)a->data.data();
        case mxSINGLE_CLASS: return *(const float*)a->data.data();
        case mxINT32_CLASS: return *(const int32_t*)a->data.data();
        case mxUINT64_CLASS: return (double)*(const uint64_t*)a->data.data();
        default: return 0.0;
    }
}
mwSize mxGetM(const mxArray* a) { return a->m; }
mwSize mxGetN(const mxArray* a) { return a->n; }
size_t mxGetNumberOfElements(const mxArray* a) { return a->cls == mxCHAR_CLASS ? a->str.size() : a->m * a->n; }
mxClassID mxGetClassID(const mxArray* a) { return a->cls; }
int mxIsDouble(const mxArray* a) { return a->cls == mxDOUBLE_CLASS; }
int mxIsSingle(const mxArray* a) { return a->cls == mxSINGLE_CLASS; }
int mxIsChar(const mxArray* a) { return a->cls == mxCHAR_CLASS; }
int mxIsStruct(const mxArray* a) { return a->cls == mxSTRUCT_CLASS; }
int mxIsEmpty(const mxArray* a) { return a->cls == mxCHAR_CLASS ? a->str.empty() : a->m * a->n == 0; }
int mxIsComplex(const mxArray*) { return 0; }
mxArray* mxGetField(const mxArray* a, mwSize, const char* name) {
    auto it = a->fields.find(name);
    return it == a->fields.end() ? nullptr : it->second;
}
int mxGetString(const mxArray* a, char* buf, mwSize len) {
    if (a->cls != mxCHAR_CLASS || a->str.size() + 1 > len) return 1;
    memcpy(buf, a->str.c_str(), a->str.size() + 1);
    return 0;
}
void mxDestroyArray(mxArray* a) { if (a) { --g_live; delete a; } }
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    throw FakeMexError(std::string(id) + ": " + buf);
}
void mexLock(void) { ++g_locks; }
int mexAtExit(void (*fn)(void)) { g_atexit = fn; return 0; }
}

// ---------------------------------------------------------------------------------------------------------
// stubs of the C ABI: record the arguments, return canned results
// ---------------------------------------------------------------------------------------------------------
struct Rec {
    int tf_mode = -1, tf_is_double = -1; int64_t tf_n = 0; double tf_t41 = 0.0;
    int init_dev = -1, init_ndev = 0, init_last = -1, inits = 0, shutdowns = 0;
    int model_is_double = -1; int64_t model_n = 0, model_ld = 0; int model_grid = -1;
    int nn_kind = -1; int64_t nn_nq = 0; int nn_is_double = -1;
    int fail_next = 0;                          // next compute call returns PCREG_ERR_CUDA
    int kabsch_status = 0; std::vector<double> kabsch_p1;
    std::vector<int32_t> tri; int64_t ransac_P = 0, ransac_H = 0; pcreg_ransac_opts ransac_o{}; int ransac_rc = PCREG_OK; uint64_t ransac_seed = 0;
    std::vector<int64_t> batch_off; std::vector<uint64_t> batch_seeds; int64_t batch_iter = 0; int64_t batch_ld = 0;
    pcreg_match_opts match_o{}; int64_t match_n1 = 0, match_n2 = 0, match_dim = 0;
    int align_kind = -1; pcreg_align_opts align_o{}; int align_status = 0; int align_is_double = -1;
    pcreg_icp_opts icp_o{}; int64_t icp_ns = 0, icp_H = 0; bool icp_has_w = false; bool icp_has_idx = false;
    double lp_R = 0; int64_t lp_min = 0, lp_max = 0;
    pcreg_desc_opts desc_o{}; int desc_nr = 0, desc_nt = 0, desc_np = 0;
    void* destroyed = nullptr;
};
static Rec R;
static pcreg_model* const HANDLE = (pcreg_model*)(uintptr_t)0xABCD1234u;
#define MAYBE_FAIL() do { if (R.fail_next) { R.fail_next = 0; return PCREG_ERR_CUDA; } } while (0)

extern "C" {
int pcreg_init(const int* devices, int ndev) { R.init_dev = ndev > 0 ? devices[0] : -1; R.init_ndev = ndev; R.init_last = ndev > 0 ? devices[ndev - 1] : -1; ++R.inits; return PCREG_OK; }
int pcreg_shutdown(void) { ++R.shutdowns; return PCREG_OK; }
const char* pcreg_last_error(void) { return "stub: device fell over"; }
int pcreg_model_create(const void*, int is_double, int64_t n, int64_t ld, const pcreg_model_opts* o, pcreg_model** out) {
    MAYBE_FAIL();
    R.model_is_double = is_double; R.model_n = n; R.model_ld = ld; R.model_grid = o->build_grid; *out = HANDLE;
    return PCREG_OK;
}
int pcreg_model_destroy(pcreg_model* m) { R.destroyed = m; return PCREG_OK; }
int pcreg_quick_tf(const void* pts, int is_double, int64_t n, int64_t ld, const double* T16, int mode, void* out, int64_t ld_out) {
    MAYBE_FAIL();
    R.tf_mode = mode; R.tf_n = n; R.tf_is_double = is_double; R.tf_t41 = T16[3];          // element (4,1) of the column-major record
    if (is_double) for (int64_t i = 0; i < n; ++i) ((double*)out)[ld_out + i] = ((const double*)pts)[ld + i] + 100.0;   // y column + 100
    return PCREG_OK;
}
int pcreg_nn_search(const pcreg_model* m, const void*, int is_double, int64_t nq, int64_t, int kind, int32_t* idx, double* d2) {
    MAYBE_FAIL();
    if (m != HANDLE) return PCREG_ERR_STATE;
    R.nn_kind = kind; R.nn_nq = nq; R.nn_is_double = is_double;
    for (int64_t i = 0; i < nq; ++i) { idx[i] = (int32_t)i; if (d2) d2[i] = 10.0 + (double)i; }      // 0-based
    return PCREG_OK;
}
int pcreg_local_points_count(const pcreg_model*, const double*, int64_t nc, int64_t, double Rr, int64_t mn, int64_t mx, int64_t* counts, int32_t* status) {
    MAYBE_FAIL();
    R.lp_R = Rr; R.lp_min = mn; R.lp_max = mx;
    for (int64_t k = 0; k < nc; ++k) { counts[k] = (k % 2 == 0) ? 3 - k / 2 : 7; status[k] = (k % 2 == 0) ? 0 : 1; }   // 3, [], 2, [] ...
    return PCREG_OK;
}
int pcreg_local_points_fill(const pcreg_model*, const double*, int64_t nc, int64_t, double, const int64_t* offsets, const int32_t* status,
                            double* pts, int64_t ld_out, double* dists, int32_t*) {
    MAYBE_FAIL();
    for (int64_t k = 0; k < nc; ++k)
        if (!status[k])
            for (int64_t i = offsets[k]; i < offsets[k + 1]; ++i) {
                for (int a = 0; a < 3; ++a) pts[a * ld_out + i] = 100.0 * (double)k + 10.0 * (double)a + (double)(i - offsets[k]);
                dists[i] = 0.5 * (double)i;
            }
    return PCREG_OK;
}
void pcreg_desc_opts_default(pcreg_desc_opts* o) { o->min_pts = 500; o->max_pts = 6000; o->R = 3.5; o->thVar[0] = o->thVar[1] = 1.0; o->k_frac = 0.0; o->align_points = 1; }
int pcreg_spatial_histogram(const pcreg_model*, const double*, int64_t nkey, int64_t, const pcreg_desc_opts* o, const double*, int nr,
                            const double*, int nt, const double*, int np, double* desc, int32_t* status, int64_t*) {
    MAYBE_FAIL();
    R.desc_o = *o; R.desc_nr = nr; R.desc_nt = nt; R.desc_np = np;
    const int64_t nb = (int64_t)nr * nt * np;
    for (int64_t k = 0; k < nkey; ++k) {
        status[k] = (k == 1) ? 2 : 0;                                              // the second keypoint is rejected
        for (int64_t j = 0; j < nb; ++j) desc[k * nb + j] = status[k] ? NAN : (double)(1000 * k + j);
    }
    return PCREG_OK;
}
void pcreg_match_opts_default(pcreg_match_opts* o) {
    o->unnormalize = 1; o->norm_factor = 2.0; o->change_metric = 1; o->metric_factor = 0.6; o->match_threshold = 10.0; o->max_ratio = 0.99;
    o->metric = PCREG_METRIC_SAD; o->unique = 1;
}
int pcreg_get_matches(const double*, int64_t n1, int64_t, const double*, int64_t n2, int64_t, int64_t dim, const pcreg_match_opts* o,
                      int32_t* pairs, double* metric, int64_t* n_matches) {
    MAYBE_FAIL();
    R.match_o = *o; R.match_n1 = n1; R.match_n2 = n2; R.match_dim = dim;
    pairs[0] = 0; pairs[1] = 1; pairs[2] = 2; pairs[3] = 0;                         // (surface 0, model 1), (surface 2, model 0)
    metric[0] = 0.25; metric[1] = 0.75;
    *n_matches = 2;
    return PCREG_OK;
}
void pcreg_align_opts_default(pcreg_align_opts* o) { o->k_frac = 0.85; o->k_abs = 0; o->R_w = 3.5; o->r_local = 2.0; o->min_local = 25; o->C1 = 0; o->C2 = 0; }
int pcreg_align_points(int kind, const void* pts, int is_double, int64_t ld, const int64_t* offsets, int64_t nbatch, const pcreg_align_opts* o,
                       void* out, double* coeff9, double* c3, int32_t* status) {
    MAYBE_FAIL();
    R.align_kind = kind; R.align_o = *o; R.align_is_double = is_double;
    const int64_t n = offsets[nbatch];
    for (int64_t i = 0; i < 3 * ld && i < 3 * n; ++i) {
        if (is_double) ((double*)out)[i] = ((const double*)pts)[i] + 1.0; else ((float*)out)[i] = ((const float*)pts)[i] + 1.0f;
    }
    for (int k = 0; k < 9; ++k) coeff9[k] = (double)k;
    c3[0] = 7.0; c3[1] = 8.0; c3[2] = 9.0;
    status[0] = R.align_status;
    return PCREG_OK;
}
int pcreg_kabsch_batch(const double* p1, const double*, const double* w, int64_t ld, const int64_t* offsets, int64_t nbatch, int, double* T16, int32_t* status) {
    MAYBE_FAIL();
    R.kabsch_p1.assign(p1, p1 + 3 * ld);
    (void)w; (void)offsets; (void)nbatch;
    for (int k = 0; k < 16; ++k) T16[k] = (double)k;
    status[0] = R.kabsch_status;
    return PCREG_OK;
}
static void ransac_outputs(int64_t P, double* T16, int32_t* inl, int64_t* n_inl, int64_t* n_succ, int64_t* max_inl, int64_t* best) {
    for (int k = 0; k < 16; ++k) T16[k] = 0.5 * (double)k;
    inl[0] = 0; inl[1] = 2; inl[2] = (int32_t)P - 1;                                // 0-based, ascending
    *n_inl = 3; *n_succ = 7; *max_inl = 3; *best = 4;
}
int pcreg_ransac_score(const double*, const double*, int64_t P, int64_t, const int32_t* tri, int64_t nhyp, const pcreg_ransac_opts* o, double* T16,
                       int32_t* inl, int64_t* n_inl, int64_t* n_succ, int64_t* max_inl, int64_t* best, int32_t*, int32_t*, double*) {
    MAYBE_FAIL();
    R.tri.assign(tri, tri + 3 * nhyp); R.ransac_P = P; R.ransac_H = nhyp; R.ransac_o = *o;
    if (R.ransac_rc == PCREG_DEGENERATE) return PCREG_DEGENERATE;
    ransac_outputs(P, T16, inl, n_inl, n_succ, max_inl, best);
    return PCREG_OK;
}
int pcreg_ransac_run(const double*, const double*, int64_t P, int64_t, int64_t iter_num, uint64_t seed, const pcreg_ransac_opts* o, double* T16,
                     int32_t* inl, int64_t* n_inl, int64_t* n_succ, int64_t* max_inl, int64_t* best, int32_t*) {
    MAYBE_FAIL();
    R.ransac_P = P; R.ransac_H = iter_num; R.ransac_o = *o; R.ransac_seed = seed;
    ransac_outputs(P, T16, inl, n_inl, n_succ, max_inl, best);
    return PCREG_OK;
}
int pcreg_ransac_batch(const double*, const double*, int64_t ld, const int64_t* offsets, int64_t nwin, int64_t iter_num, const int32_t* tri,
                       const uint64_t* seeds, const pcreg_ransac_opts* o, double* T16, int32_t* inl, int64_t* n_inl, int64_t* n_succ,
                       int64_t* max_inl, int64_t* best, int32_t* status) {
    MAYBE_FAIL();
    (void)tri;
    R.batch_off.assign(offsets, offsets + nwin + 1); R.batch_seeds.assign(seeds, seeds + nwin); R.batch_iter = iter_num; R.batch_ld = ld; R.ransac_o = *o;
    for (int64_t w = 0; w < nwin; ++w) {
        const int64_t P = offsets[w + 1] - offsets[w];
        status[w] = P < 3 ? 1 : 0;
        for (int k = 0; k < 16; ++k) T16[w * 16 + k] = status[w] ? NAN : 100.0 * (double)w + (double)k;
        n_inl[w] = 0; n_succ[w] = 0; max_inl[w] = 0; best[w] = -1;
        if (!status[w]) {                                                           // first and last pair of the window, relative and 0-based
            inl[offsets[w]] = 0; inl[offsets[w] + 1] = (int32_t)P - 1;
            n_inl[w] = 2; n_succ[w] = 10 + w; max_inl[w] = 2; best[w] = w;
        }
    }
    return PCREG_OK;
}
void pcreg_icp_opts_default(pcreg_icp_opts* o) { o->mode = PCREG_ICP_PLAIN; o->iters = 50; o->k_frac = 0.85; o->R_w = 3.5; o->thDist2 = 0.0; o->nn = PCREG_NN_BRUTE; o->reflection_fix = 0; }
int pcreg_icp_batch(const pcreg_model* m, const void*, int, int64_t ns, int64_t, const double* w, const double* T0, int64_t nhyp, const pcreg_icp_opts* o,
                    double* T16, double* rmse, int32_t* n_used, int32_t* status, int32_t* idx, double*, int64_t* best) {
    MAYBE_FAIL();
    if (m != HANDLE) return PCREG_ERR_STATE;
    R.icp_o = *o; R.icp_ns = ns; R.icp_H = nhyp; R.icp_has_w = w != nullptr; R.icp_has_idx = idx != nullptr;
    for (int64_t h = 0; h < nhyp; ++h) {
        for (int k = 0; k < 16; ++k) T16[h * 16 + k] = T0[h * 16 + k] + 1.0;
        rmse[h] = 0.1 * (double)(h + 1); n_used[h] = (int32_t)(ns - h); status[h] = (int32_t)(h % 2);
        if (idx) for (int64_t i = 0; i < ns; ++i) idx[h * ns + i] = (int32_t)i;       // 0-based
    }
    *best = 1;
    return PCREG_OK;
}
}

// ---------------------------------------------------------------------------------------------------------
// the test driver
// ---------------------------------------------------------------------------------------------------------
static int g_checks = 0, g_failed = 0;
#define CHECK(cond) do { ++g_checks; if (!(cond)) { ++g_failed; fprintf(stderr, "CHECK failed at line %d: %s\n", __LINE__, #cond); } } while (0)

static std::vector<mxArray*> g_inputs;           // owned by the driver ("MATLAB workspace")
static mxArray* dbl(size_t m, size_t n, std::vector<double> v = {}) {
    mxArray* a = new_array(mxDOUBLE_CLASS, m, n); --g_live;
    for (size_t i = 0; i < v.size() && i < m * n; ++i) mxGetPr(a)[i] = v[i];
    g_inputs.push_back(a); return a;
}
static mxArray* sgl(size_t m, size_t n) { mxArray* a = new_array(mxSINGLE_CLASS, m, n); --g_live; for (size_t i = 0; i < m * n; ++i) ((float*)a->data.data())[i] = (float)i; g_inputs.push_back(a); return a; }
static mxArray* str(const char* s) { mxArray* a = new_array(mxCHAR_CLASS, 1, strlen(s)); --g_live; a->str = s; g_inputs.push_back(a); return a; }
static mxArray* u64(uint64_t v) { mxArray* a = new_array(mxUINT64_CLASS, 1, 1); --g_live; *(uint64_t*)a->data.data() = v; g_inputs.push_back(a); return a; }
static mxArray* strct(std::map<std::string, mxArray*> f) { mxArray* a = new_array(mxSTRUCT_CLASS, 1, 1); --g_live; a->fields = std::move(f); g_inputs.push_back(a); return a; }

struct Out { mxArray* a[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; std::string err; };
static Out call(int nlhs, std::vector<const mxArray*> rhs) {
    Out o;
    try {
        mexFunction(nlhs, o.a, (int)rhs.size(), rhs.data());
    } catch (const FakeMexError& e) {
        o.err = e.what();
    }
    return o;
}
static void release(Out& o) { for (auto& p : o.a) if (p) { mxDestroyArray(p); p = nullptr; } }
static bool is_empty00(const mxArray* a) { return a && a->cls == mxDOUBLE_CLASS && a->m == 0 && a->n == 0; }

int main() {
    // ---- init: explicit device, mexLock once, atexit registered ----
    { Out o = call(0, {str("init"), dbl(1, 3, {2, 3, 5})}); CHECK(o.err.empty()); CHECK(R.init_dev == 2 && R.init_ndev == 3 && R.init_last == 5 && R.inits == 1 && g_locks == 1 && g_atexit != nullptr); release(o); }   // a device vector: one process drives several GPUs
    { Out o = call(0, {str("init")}); CHECK(R.inits == 1); release(o); }                              // second init is a no-op

    // ---- model_create: class single stays single, ld = n, grid by default; handle travels as uint64 ----
    mxArray* handle = nullptr;
    { Out o = call(1, {str("model_create"), sgl(11, 3)});
      CHECK(o.err.empty() && R.model_is_double == 0 && R.model_n == 11 && R.model_ld == 11 && R.model_grid == 1);
      CHECK(o.a[0] && o.a[0]->cls == mxUINT64_CLASS && *(uint64_t*)mxGetData(o.a[0]) == (uint64_t)(uintptr_t)HANDLE);
      handle = u64(*(uint64_t*)mxGetData(o.a[0])); release(o); }
    { Out o = call(1, {str("model_create"), dbl(4, 3), dbl(1, 1, {0})}); CHECK(R.model_is_double == 1 && R.model_grid == 0); release(o); }
    { Out o = call(1, {str("model_create"), dbl(4, 2)}); CHECK(!o.err.empty() && o.err.find("N x 3") != std::string::npos); release(o); }

    // ---- nn_search: 0-based indices come back 1-based, as doubles ----
    { Out o = call(2, {str("nn_search"), handle, dbl(5, 3)});
      CHECK(o.err.empty() && R.nn_kind == PCREG_NN_BRUTE && R.nn_nq == 5 && R.nn_is_double == 1);
      CHECK(o.a[0]->m == 5 && o.a[0]->n == 1 && mxGetPr(o.a[0])[0] == 1.0 && mxGetPr(o.a[0])[4] == 5.0 && mxGetPr(o.a[1])[3] == 13.0); release(o); }
    { Out o = call(1, {str("nn_search"), handle, dbl(2, 3), str("grid")}); CHECK(R.nn_kind == PCREG_NN_GRID); release(o); }
    { Out o = call(1, {str("nn_search"), dbl(1, 1, {5}), dbl(2, 3)}); CHECK(!o.err.empty()); release(o); }          // not a uint64 handle

    // ---- quick_tf: class kept, the 4x4 handed over as MATLAB stores it, mode optional ----
    { Out o = call(1, {str("quick_tf"), dbl(3, 3, {1, 2, 3, 4, 5, 6, 7, 8, 9}), dbl(4, 4, {1, 0, 0, 13, 0, 1, 0, 25, 0, 0, 1, -17, 0, 0, 0, 1}), dbl(1, 1, {2})});
      CHECK(o.err.empty() && R.tf_mode == PCREG_TF_MRDIVIDE && R.tf_n == 3 && R.tf_is_double == 1 && R.tf_t41 == 13.0);
      CHECK(o.a[0]->m == 3 && o.a[0]->n == 3 && mxGetPr(o.a[0])[3] == 104.0); release(o); }
    { Out o = call(1, {str("quick_tf"), dbl(3, 3), dbl(3, 3)}); CHECK(!o.err.empty()); release(o); }                    // TF must be 4 x 4

    // ---- estimate_transform: 4x4 column-major as the ABI wrote it; status 1 -> [] ----
    { R.kabsch_status = 0; Out o = call(1, {str("estimate_transform"), dbl(4, 3, {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12}), dbl(4, 3)});
      CHECK(o.err.empty() && o.a[0]->m == 4 && o.a[0]->n == 4 && mxGetPr(o.a[0])[7] == 7.0 && R.kabsch_p1.size() == 12 && R.kabsch_p1[5] == 6.0); release(o); }
    { R.kabsch_status = 1; Out o = call(1, {str("estimate_transform"), dbl(3, 3), dbl(3, 3)}); CHECK(is_empty00(o.a[0])); release(o); R.kabsch_status = 0; }
    { Out o = call(1, {str("estimate_transform"), dbl(3, 3), dbl(4, 3)}); CHECK(!o.err.empty()); release(o); }

    // ---- ransac: Hx3 1-based column-major triplets -> hypothesis-major 0-based; outputs back to 1-based ----
    mxArray* coef = strct({{"thDist", dbl(1, 1, {0.2})}, {"thInlrRatio", dbl(1, 1, {0.15})}, {"REFINE", dbl(1, 1, {0})}, {"iterNum", dbl(1, 1, {321})}});
    { Out o = call(5, {str("ransac"), dbl(9, 3), dbl(9, 3), coef, dbl(2, 3, {1, 4, 2, 5, 3, 6})});    // rows (1,2,3) and (4,5,6)
      CHECK(o.err.empty() && R.ransac_P == 9 && R.ransac_H == 2);
      CHECK(R.tri == std::vector<int32_t>({0, 1, 2, 3, 4, 5}));
      CHECK(R.ransac_o.thDist == 0.2 && R.ransac_o.thInlrRatio == 0.15 && R.ransac_o.refine == 0 && R.ransac_o.reflection_fix == 0);
      CHECK(o.a[0]->m == 4 && mxGetPr(o.a[0])[6] == 3.0);
      CHECK(o.a[1]->m == 3 && o.a[1]->n == 1 && mxGetPr(o.a[1])[0] == 1.0 && mxGetPr(o.a[1])[1] == 3.0 && mxGetPr(o.a[1])[2] == 9.0);
      CHECK(mxGetScalar(o.a[2]) == 7.0 && mxGetScalar(o.a[3]) == 3.0 && fabs(mxGetScalar(o.a[4]) - 100.0 * 3.0 / 9.0) < 1e-12); release(o); }
    { R.ransac_rc = PCREG_DEGENERATE; Out o = call(5, {str("ransac"), dbl(9, 3), dbl(9, 3), coef, dbl(1, 3, {1, 2, 3})});   // ransac.m:75-89
      CHECK(o.err.empty() && is_empty00(o.a[0]) && is_empty00(o.a[1]) && mxGetScalar(o.a[2]) == 0.0 && mxGetScalar(o.a[4]) == 0.0); release(o); R.ransac_rc = PCREG_OK; }
    { Out o = call(1, {str("ransac_seeded"), dbl(9, 3), dbl(9, 3), coef, dbl(1, 1, {77})}); CHECK(o.err.empty() && R.ransac_H == 321 && R.ransac_seed == 77); release(o); }

    // ---- ransac_batch: offsets / seeds through, T 16 x W with NaN columns, inlier MASK over all rows ----
    { Out o = call(5, {str("ransac_batch"), dbl(10, 3), dbl(10, 3), dbl(4, 1, {0, 4, 6, 10}), coef, dbl(3, 1, {5, 6, 7})});
      CHECK(o.err.empty() && R.batch_off == std::vector<int64_t>({0, 4, 6, 10}) && R.batch_seeds == std::vector<uint64_t>({5, 6, 7}) && R.batch_iter == 321 && R.batch_ld == 10);
      CHECK(o.a[0]->m == 16 && o.a[0]->n == 3 && mxGetPr(o.a[0])[3] == 3.0 && std::isnan(mxGetPr(o.a[0])[16]) && mxGetPr(o.a[0])[32 + 5] == 205.0);
      const double* mk = mxGetPr(o.a[1]);
      CHECK(o.a[1]->m == 10 && mk[0] == 1.0 && mk[3] == 1.0 && mk[1] == 0.0 && mk[4] == 0.0 && mk[5] == 0.0 && mk[6] == 1.0 && mk[9] == 1.0 && mk[7] == 0.0);
      CHECK(mxGetPr(o.a[2])[0] == 10.0 && mxGetPr(o.a[2])[1] == 0.0 && mxGetPr(o.a[2])[2] == 12.0);
      CHECK(mxGetPr(o.a[3])[0] == 2.0 && fabs(mxGetPr(o.a[4])[0] - 50.0) < 1e-12 && mxGetPr(o.a[4])[1] == 0.0 && fabs(mxGetPr(o.a[4])[2] - 50.0) < 1e-12); release(o); }
    { Out o = call(1, {str("ransac_batch"), dbl(10, 3), dbl(10, 3), dbl(3, 1, {0, 4, 9}), coef, dbl(2, 1, {1, 2})}); CHECK(!o.err.empty() && o.err.find("offsets") != std::string::npos); release(o); }

    // ---- get_matches: struct parsing incl. the char field, uint32 P x 2 1-based pairs ----
    mxArray* par = strct({{"UNNORMALIZE", dbl(1, 1, {1})}, {"norm_factor", dbl(1, 1, {2})}, {"CHANGE_METRIC", dbl(1, 1, {1})}, {"metric_factor", dbl(1, 1, {0.6})},
                          {"MatchThreshold", dbl(1, 1, {10})}, {"MaxRatio", dbl(1, 1, {0.99})}, {"Metric", str("SAD")}, {"Unique", dbl(1, 1, {1})}, {"Method", str("Approximate")}});
    { Out o = call(2, {str("get_matches"), dbl(3, 7), dbl(5, 7), par});
      CHECK(o.err.empty() && R.match_n1 == 3 && R.match_n2 == 5 && R.match_dim == 7 && R.match_o.metric == PCREG_METRIC_SAD && R.match_o.unique == 1 &&
            R.match_o.unnormalize == 1 && R.match_o.max_ratio == 0.99 && R.match_o.match_threshold == 10.0 && R.match_o.metric_factor == 0.6);
      const uint32_t* pr = (const uint32_t*)mxGetData(o.a[0]);
      CHECK(o.a[0]->cls == mxUINT32_CLASS && o.a[0]->m == 2 && o.a[0]->n == 2 && pr[0] == 1 && pr[1] == 3 && pr[2] == 2 && pr[3] == 1);
      CHECK(mxGetPr(o.a[1])[1] == 0.75); release(o); }
    { mxArray* bad = strct({{"Metric", str("L7")}}); Out o = call(1, {str("get_matches"), dbl(3, 7), dbl(5, 7), bad}); CHECK(!o.err.empty()); release(o); }
    { Out o = call(1, {str("get_matches"), dbl(3, 7), dbl(5, 6), par}); CHECK(!o.err.empty()); release(o); }

    // ---- align: options per kind, class preserved, status 1 -> [] [] but the centroid stays ----
    { Out o = call(3, {str("align"), dbl(1, 1, {1}), sgl(6, 3), dbl(1, 1, {1}), dbl(1, 1, {0})});
      CHECK(o.err.empty() && R.align_kind == PCREG_ALIGN_KNN_FRAC && R.align_o.C1 == 1 && R.align_o.C2 == 0 && R.align_is_double == 0);
      CHECK(o.a[0]->cls == mxSINGLE_CLASS && o.a[0]->m == 6 && ((float*)mxGetData(o.a[0]))[4] == 5.0f && mxGetPr(o.a[1])[8] == 8.0 && mxGetPr(o.a[2])[1] == 8.0); release(o); }
    { Out o = call(1, {str("align"), dbl(1, 1, {2}), dbl(6, 3), dbl(1, 1, {1500})}); CHECK(R.align_kind == PCREG_ALIGN_KNN_ABS && R.align_o.k_abs == 1500); release(o); }
    { R.align_status = 1; Out o = call(3, {str("align"), dbl(1, 1, {4}), dbl(6, 3)}); CHECK(is_empty00(o.a[0]) && is_empty00(o.a[1]) && o.a[2]->n == 3); release(o); R.align_status = 0; }

    // ---- local_points: concatenated neighbourhoods, per-centre counts with 0 where the reference returns [] ----
    { Out o = call(3, {str("local_points"), handle, dbl(3, 3), dbl(1, 1, {3.5}), dbl(1, 1, {30}), dbl(1, 1, {INFINITY})});
      CHECK(o.err.empty() && R.lp_R == 3.5 && R.lp_min == 30 && R.lp_max == -1);
      CHECK(o.a[0]->m == 5 && o.a[0]->n == 3 && mxGetPr(o.a[0])[0] == 0.0 && mxGetPr(o.a[0])[3] == 200.0 && mxGetPr(o.a[0])[5 + 4] == 211.0);
      CHECK(o.a[1]->m == 5 && mxGetPr(o.a[2])[0] == 3.0 && mxGetPr(o.a[2])[1] == 0.0 && mxGetPr(o.a[2])[2] == 2.0); release(o); }

    // ---- spatial_histogram: options struct ('all' as char), only the surviving keypoints in keypoint order ----
    { mxArray* opt = strct({{"min_pts", dbl(1, 1, {150})}, {"max_pts", dbl(1, 1, {INFINITY})}, {"R", dbl(1, 1, {3.0})}, {"thVar", dbl(1, 2, {1.2, 1.5})},
                            {"k", str("all")}, {"ALIGN_POINTS", dbl(1, 1, {0})}});
      Out o = call(2, {str("spatial_histogram"), handle, dbl(3, 3, {1, 2, 3, 4, 5, 6, 7, 8, 9}), opt, dbl(1, 3), dbl(1, 4), dbl(1, 3)});
      CHECK(o.err.empty() && R.desc_o.min_pts == 150 && R.desc_o.max_pts == -1 && R.desc_o.R == 3.0 && R.desc_o.thVar[1] == 1.5 && R.desc_o.k_frac == 0.0 &&
            R.desc_o.align_points == 0 && R.desc_nr == 2 && R.desc_nt == 3 && R.desc_np == 2);
      CHECK(o.a[0]->m == 2 && o.a[0]->n == 3 && mxGetPr(o.a[0])[0] == 1.0 && mxGetPr(o.a[0])[1] == 3.0 && mxGetPr(o.a[0])[2 + 1] == 6.0);
      CHECK(o.a[1]->m == 2 && o.a[1]->n == 12 && mxGetPr(o.a[1])[0] == 0.0 && mxGetPr(o.a[1])[1] == 2000.0 && mxGetPr(o.a[1])[2 * 5 + 1] == 2005.0); release(o);
      mxArray* opt2 = strct({{"k", dbl(1, 1, {0.85})}});
      Out o2 = call(1, {str("spatial_histogram"), handle, dbl(3, 3), opt2, dbl(1, 3), dbl(1, 4), dbl(1, 3)}); CHECK(R.desc_o.k_frac == 0.85 && R.desc_o.min_pts == 500); release(o2); }

    // ---- icp: T0 as 16 x H, opts struct over the defaults, 1-based best / idx ----
    { mxArray* opts = strct({{"mode", dbl(1, 1, {1})}, {"iters", dbl(1, 1, {30})}, {"nn", dbl(1, 1, {1})}, {"thDist2", dbl(1, 1, {4})}});
      std::vector<double> t0(32); for (int i = 0; i < 32; ++i) t0[i] = i;
      Out o = call(6, {str("icp"), handle, dbl(4, 3), dbl(16, 2, t0), opts, dbl(4, 1, {1, 1, 1, 1})});
      CHECK(o.err.empty() && R.icp_o.mode == PCREG_ICP_KNN && R.icp_o.iters == 30 && R.icp_o.nn == PCREG_NN_GRID && R.icp_o.thDist2 == 4.0 && R.icp_o.k_frac == 0.85);
      CHECK(R.icp_ns == 4 && R.icp_H == 2 && R.icp_has_w && R.icp_has_idx);
      CHECK(o.a[0]->m == 16 && o.a[0]->n == 2 && mxGetPr(o.a[0])[17] == 18.0 && fabs(mxGetPr(o.a[1])[1] - 0.2) < 1e-15);
      CHECK(o.a[2]->cls == mxINT32_CLASS && ((int32_t*)mxGetData(o.a[2]))[1] == 3 && ((int32_t*)mxGetData(o.a[3]))[1] == 1 && mxGetScalar(o.a[4]) == 2.0);
      CHECK(o.a[5]->m == 4 && o.a[5]->n == 2 && ((int32_t*)mxGetData(o.a[5]))[0] == 1 && ((int32_t*)mxGetData(o.a[5]))[7] == 4); release(o);
      Out o2 = call(1, {str("icp"), handle, dbl(4, 3), dbl(16, 1), opts}); CHECK(o2.err.empty() && !R.icp_has_w && !R.icp_has_idx); release(o2); }

    // ---- errors: unknown command, a failing ABI call reports pcreg_last_error() and leaks nothing ----
    { Out o = call(1, {str("frobnicate")}); CHECK(o.err.find("unknown command 'frobnicate'") != std::string::npos); release(o); }
    { Out o = call(1, {dbl(1, 1, {3})}); CHECK(o.err.find("command string") != std::string::npos); release(o); }
    { const int live0 = g_live; R.fail_next = 1; Out o = call(2, {str("nn_search"), handle, dbl(5, 3)});
      CHECK(o.err.find("stub: device fell over") != std::string::npos && g_live == live0); release(o); }
    { const int live0 = g_live; R.fail_next = 1; Out o = call(6, {str("icp"), handle, dbl(4, 3), dbl(16, 1), strct({})});
      CHECK(!o.err.empty() && g_live == live0); release(o); }
    { const int live0 = g_live; R.fail_next = 1; Out o = call(3, {str("align"), dbl(1, 1, {0}), dbl(6, 3)}); CHECK(!o.err.empty() && g_live == live0); release(o); }

    // ---- model_destroy + the registered exit handler ----
    { Out o = call(0, {str("model_destroy"), handle}); CHECK(o.err.empty() && R.destroyed == (void*)HANDLE); release(o); }
    if (g_atexit) g_atexit();
    CHECK(R.shutdowns == 1);
    CHECK(g_live == 0);                           // every array the gateway created was returned to us or destroyed

    for (mxArray* a : g_inputs) delete a;
    printf("%s: %d checks, %d failed\n", g_failed ? "FAILED" : "OK", g_checks, g_failed);
    return g_failed ? 1 : 0;
}
