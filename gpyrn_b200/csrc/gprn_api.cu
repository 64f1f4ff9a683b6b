// gprn_api.cu -- C ABI (include/gprn_b200.h) and host orchestration of the batched ELBO / prediction
// pipelines.  Everything that touches numbers runs in the CUDA kernels of the *.cuh files; this file
// only sizes workspaces, builds the per-iteration launch lists and moves the (tiny) inputs/outputs.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/gprn_b200.h"
#include "common.cuh"
#include "assemble.cuh"
#include "factor.cuh"
#include "small.cuh"
#include "mid.cuh"
#include "elbo.cuh"
#include "predict.cuh"

using namespace gprn;

static thread_local std::string g_err;
static int fail(const std::string& msg) { g_err = msg; return 1; }

#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" +    \
                        std::to_string(__LINE__) + ")");                                                \
    } while (0)
static const bool g_debug_sync = getenv("GPRN_DEBUG_SYNC") != nullptr;   // serialise launches to localise a fault
// GPRN_PROFILE=1: synchronise after every launch and attribute the host wall time since the previous launch check
// to the source line of the launch (development aid; dumped to stderr by gprn_destroy).
static const bool g_profile = getenv("GPRN_PROFILE") != nullptr;
#include <chrono>
#include <map>
#include <mutex>
#include <set>
static std::mutex g_mutex;                       // guards the process-wide tables below (handles may live on different threads)
static std::map<int, std::pair<double, long>> g_prof;
static std::chrono::steady_clock::time_point g_prof_last;
static void prof_tick(int line) {
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_mutex);
    auto now = std::chrono::steady_clock::now();
    double us = std::chrono::duration<double, std::micro>(now - g_prof_last).count();
    auto& e = g_prof[line];
    e.first += us;
    e.second += 1;
    g_prof_last = std::chrono::steady_clock::now();
}
#define LAUNCH_CHECK(h)                                                                                 \
    do {                                                                                                \
        (h)->launches++;                                                                                \
        cudaError_t e_ = cudaGetLastError();                                                            \
        if (e_ == cudaSuccess && g_debug_sync && !(h)->capturing) e_ = cudaDeviceSynchronize();         \
        if (g_profile && !(h)->capturing) prof_tick(__LINE__);                                          \
        if (e_ != cudaSuccess)                                                                          \
            return fail(std::string("kernel launch: ") + cudaGetErrorString(e_) + " (" + __FILE__ +     \
                        ":" + std::to_string(__LINE__) + ")");                                          \
    } while (0)

// Launch with a per-launch priority (cudaLaunchAttributePriority): the latency-bound panel / in-block kernels of one
// stream group are dispatched ahead of the pending CTAs of another group's wide GEMM launch.
static int g_prio_hi = 0, g_prio_mode = 1;
static std::once_flag g_prio_once;
template <typename... KArgs, typename... Args>
static void launch_hi(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    std::call_once(g_prio_once, [] {
        int lo = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &g_prio_hi);
        g_prio_mode = getenv("GPRN_PANEL_PRIO") ? atoi(getenv("GPRN_PANEL_PRIO")) : 1;
    });
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributePriority;
    at[0].val.priority = g_prio_hi;
    cfg.attrs = at;
    cfg.numAttrs = g_prio_mode ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

// Device-resident variational state of `n` chains (optimiser starts, MCMC walkers, sets of a sweep): the batched form
// of the reference's self._mu / self._var cache (meanfield.py:112-113, 598-607, 643-646).
struct ChainStore {
    DevBuf mu, var, valid;
    int64_t n = 0;
};

struct gprn_handle {
    int device = 0, N = 0, Np = 0, nt = 0, p = 0, q = 0, M = 0, H = 0, d = 0;
    bool model_set = false;
    bool capturing = false;                     // inside a stream capture (CUDA graph of one iteration)
    bool small_mode = false;                    // this call runs the fused small-N pipeline (decide_small_path)
    bool mid_mode = false;                      // this call runs the multi-CTA dataflow pipeline (decide_mid_path)
    int* mid_ticket = nullptr;                  // two start-order counters behind the tile flags (mid.cuh)
    bool mid_latency = false;                   // mid_mode with every CTA resident at once (a few evaluations in flight)
    cudaStream_t own_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // stream groups: set 0 for the fixed-point iteration, set 1 for the set-up of newly admitted sets, which runs
    // next to it on `side`
    static const int NAUX = 8;
    cudaStream_t aux[2][NAUX] = {};
    cudaEvent_t ev_fork[2] = {}, ev_join[2][NAUX] = {};
    cudaStream_t side = nullptr;
    cudaEvent_t ev_side_fork = nullptr, ev_side_join = nullptr;
    cudaStream_t cross = nullptr;               // cross-node trace terms, beside the start of the weight phase
    cudaEvent_t ev_cross_fork = nullptr, ev_cross_join = nullptr;
    int64_t launches = 0;
    int64_t graph_launches = 0;
    double last_ms = 0.0;
    int64_t last_total_iters = 0;
    int64_t last_lockstep_rounds = 0;
    uint64_t ws_limit = 0;
    int max_slots = 0;
    // data
    double *d_time = nullptr, *d_y = nullptr, *d_yerr2 = nullptr, *d_ysub_shared = nullptr;
    // model
    int32_t *d_tok = nullptr, *d_len = nullptr, *d_par_off = nullptr;
    std::vector<int32_t> h_tok, h_len, h_par_off, h_npar;
    // workspace (grow only)
    std::vector<DevBuf*> all;
    DevBuf K, W, X, XK, vecs, state, small, lists, rlist, hyper, ysub, ks, pred, scratch, scratch2, gpart, res, kscratch, mid_state, mid_zp, mid_ld;
    ChainStore chain;         // user-visible store (gprn_chain_*)
    ChainStore tmp_chain;     // backing of mu_inout / var_inout of gprn_elbo_batched
    int num_sms = 148;
    // pinned staging
    int* h_lists = nullptr;
    size_t h_lists_n = 0;
    int* h_active = nullptr;
    size_t h_active_n = 0;
    int* h_ret = nullptr;
    size_t h_ret_n = 0;
    // CUDA graphs of one lock-step iteration, keyed by the number of active slots (see run_pool)
    std::map<int, cudaGraphExec_t> iter_graphs;
    std::map<int, int64_t> graph_kernels;       // kernel + memset nodes per cached graph (launch accounting)
    std::map<int, cudaGraphExec_t> loop_graphs; // device-resident iteration loop (WHILE node) per active count
    std::map<int, int64_t> loop_kernels;        // nodes of the loop body
    bool loop_unavailable = false;              // the runtime refused the conditional node: per-iteration graphs instead
    std::vector<unsigned char> graph_sig;       // bytes of the Engine (workspace addresses + context) they were captured for
};

// Fused single-kernel pipeline (small.cuh): q == 1 (no cross-node terms, which need the factors in HBM) and
// N <= 256 always; N <= 512 when the call keeps at least eight matrices per SM in flight -- one persistent CTA walks a
// whole matrix, so fewer mid-size matrices are better served by the dataflow kernels of mid.cuh (nt CTAs per matrix).
// Decided per call (decide_small_path) because the workspace layout differs.  GPRN_NO_SMALL=1 disables the path,
// GPRN_SMALL_MAX_NT=4 restricts it to N <= 256, GPRN_SMALL_MIN_FILL sets the matrices per SM it wants above N = 256.
static bool use_small_path(const gprn_handle* h) { return h->small_mode; }
static bool decide_small_path(const gprn_handle* h, int64_t sets_in_flight) {
    static const int max_nt = getenv("GPRN_SMALL_MAX_NT") ? atoi(getenv("GPRN_SMALL_MAX_NT")) : SMALL_MAX_NT;
    if (h->q != 1 || h->nt > std::min(max_nt, SMALL_MAX_NT) || getenv("GPRN_NO_SMALL") != nullptr) return false;
    // nt > 4: the dataflow kernels (mid.cuh, throughput mode) keep every SM busy with any number of matrices, the
    // persistent single-CTA kernel needs several matrices per resident CTA to amortise its tail.  Measured at N = 500,
    // p = 4 (evaluations/s, fused : dataflow): 128 sets 1.97 k : ~2.5 k, 512 sets 2.88 k : 2.56 k.
    // GPRN_SMALL_MIN_FILL overrides the matrices-per-SM threshold.
    const char* fe = getenv("GPRN_SMALL_MIN_FILL");
    const int fill = fe ? std::max(1, atoi(fe)) : 8;
    return h->nt <= 4 || sets_in_flight * h->M >= fill * (int64_t)h->num_sms;
}

// Threads of the O(N) per-matrix kernels (prep_*, post): one element per thread up to 1024.  A function of N only, never
// of the batch -- post_kernel's block sums depend on it.
static int vec_threads(const gprn_handle* h) { return std::min(1024, std::max(256, h->Np)); }

// Latency path (mid.cuh): N <= 1024 and so few matrices in flight that every CTA of a launch (nt per matrix) is
// resident at once -- a single ELBOcalc, a handful of walkers.  One matrix is then worked on by nt CTAs coupled by
// tile flags instead of by one CTA (small.cuh) or ~20 dependent launches (factor.cuh).  Any q: for q > 1 the set-up
// also inverts chol(K) (transposed tiles, read by cross_frob_kernel<true> and mid_trmv_lower_kernel).
// GPRN_NO_MID=1 disables it.
static bool use_mid_path(const gprn_handle* h) { return h->mid_mode; }
static bool mid_colocated(const gprn_handle* h, int64_t sets_in_flight) {
    return sets_in_flight * h->M * h->nt <= 2 * (int64_t)h->num_sms;
}
static bool decide_mid_path(const gprn_handle* h, int64_t sets_in_flight) {
    if (h->nt > MID_MAX_NT || getenv("GPRN_NO_MID") != nullptr) return false;
    if (mid_colocated(h, sets_in_flight)) return true;                 // latency mode (N <= 1024)
    if (h->nt > 8) return false;       // 512 < N <= 1024 in batches: the GEMM-based kernels of factor.cuh
    // throughput mode (CTAs numbered by start order): wherever the fused single-CTA kernel does not apply -- q > 1, or
    // 256 < N <= 512 with fewer than eight matrices per SM -- instead of the ~20 dependent launches per matrix
    // of factor.cuh.  GPRN_MID_COLOCATED_ONLY=1 restores the multi-kernel path there.
    if (getenv("GPRN_MID_COLOCATED_ONLY") != nullptr) return false;
    return !decide_small_path(h, sets_in_flight);
}

static void drop_graphs(gprn_handle* h) {
    for (auto& kv : h->iter_graphs) cudaGraphExecDestroy(kv.second);
    for (auto& kv : h->loop_graphs) cudaGraphExecDestroy(kv.second);
    h->iter_graphs.clear();
    h->graph_kernels.clear();
    h->loop_graphs.clear();
    h->loop_kernels.clear();
    h->graph_sig.clear();
}

static int ensure(DevBuf& b, size_t bytes) {
    if (b.bytes >= bytes) return 0;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (e != cudaSuccess) return fail(std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
    b.bytes = bytes;
    return 0;
}
static int ensure_zeroed(DevBuf& b, size_t bytes) {
    if (b.bytes >= bytes) return 0;
    if (ensure(b, bytes)) return 1;
    cudaError_t e = cudaMemset(b.p, 0, b.bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return fail(std::string("cudaMemset: ") + cudaGetErrorString(e));
    return 0;
}
static int ensure_pinned(int*& p, size_t& n, size_t want) {
    if (n >= want) return 0;
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = 0;
    cudaError_t e = cudaMallocHost(&p, want * sizeof(int));
    if (e != cudaSuccess) return fail(std::string("cudaMallocHost: ") + cudaGetErrorString(e));
    n = want;
    return 0;
}

// Padded matrix order: multiples of 64; beyond 1024 multiples of 256 so that the two-level (large-N) path applies.
static int padded_size(int n) {
    const int unit = n > 1024 ? OUTER_KB : NB;
    return ((n + unit - 1) / unit) * unit;
}
static bool use_two_level(int Np) { return Np >= 512 && Np % OUTER_KB == 0 && getenv("GPRN_NO_TWO_LEVEL") == nullptr; }

// Opt-in to > 48 KB of dynamic shared memory.  The attribute belongs to the (function, device) pair, so it is set
// once per device ordinal: handles on different GPUs of one process each need it.
static std::set<int> g_attr_devices;
static int set_kernel_attrs(int device) {
    std::lock_guard<std::mutex> lk(g_mutex);
    if (g_attr_devices.count(device)) return 0;
    CU(cudaFuncSetAttribute(panel_col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PANEL_SMEM));
    CU(cudaFuncSetAttribute(potrf_col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_COL_SMEM));
    CU(cudaFuncSetAttribute(trsm_col_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRSM_COL_SMEM(1)));
    CU(cudaFuncSetAttribute(trsm_col_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRSM_COL_SMEM(2)));
    CU(cudaFuncSetAttribute(syrk_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * TILE_SMEM)));
    CU(cudaFuncSetAttribute(trtri_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRTRI_SMEM));
    CU(cudaFuncSetAttribute(trtri_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRTRI_DIAG_SMEM));
    CU(cudaFuncSetAttribute(syrk_outer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM128_SMEM));
    CU(cudaFuncSetAttribute(trtri_outer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM128_SMEM));
    CU(cudaFuncSetAttribute(trtri_inblock_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRTRI_SMEM));
    CU(cudaFuncSetAttribute(small_pipeline_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_SMEM + 8192));
    CU(cudaFuncSetAttribute(small_pipeline_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_SMEM + 8192));
    CU(cudaFuncSetAttribute(mid_pipeline_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MID_SMEM));
    CU(cudaFuncSetAttribute(mid_pipeline_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MID_SMEM));
    CU(cudaFuncSetAttribute(cross_frob_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * TILE_SMEM)));
    CU(cudaFuncSetAttribute(cross_frob_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * TILE_SMEM)));
    CU(cudaFuncSetAttribute(predict_norm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * TILE_SMEM)));
    CU(cudaFuncSetAttribute(predict_norm128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM128_SMEM));
    g_attr_devices.insert(device);
    return 0;
}

static int check_device(int device) {
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail("no such CUDA device: " + std::to_string(device));
    int major = 0, minor = 0;
    CU(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    CU(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
    if (major != 10)
        return fail(std::string("this library is built for sm_100a only, device is sm_") + std::to_string(major) +
                    std::to_string(minor));
    return 0;
}

extern "C" const char* gprn_last_error(void) { return g_err.c_str(); }
extern "C" int gprn_version(void) { return 200; }
extern "C" int gprn_built_for_sm(void) { return 100; }

extern "C" int gprn_create(int device, int N, int p, int q, const double* time, const double* y, const double* yerr,
                           gprn_handle** out) {
    if (!out || !time || !y || !yerr) return fail("gprn_create: null argument");
    if (N < 1 || p < 1 || q < 1) return fail("gprn_create: N, p, q must be positive");
    if (check_device(device)) return 1;
    CU(cudaSetDevice(device));
    if (set_kernel_attrs(device)) return 1;
    gprn_handle* h = new gprn_handle();
    CU(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
    h->device = device;
    h->N = N;
    h->Np = padded_size(N);
    h->nt = h->Np / NB;
    h->p = p;
    h->q = q;
    h->M = q * (p + 1);
    h->d = N * q * (p + 1);
    CU(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&h->cross, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&h->ev_cross_fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_cross_join, cudaEventDisableTiming));
    CU(cudaEventCreate(&h->ev0));
    CU(cudaEventCreate(&h->ev1));
    CU(cudaEventCreateWithFlags(&h->ev_side_fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_side_join, cudaEventDisableTiming));
    for (int a = 0; a < 2; a++) {
        CU(cudaEventCreateWithFlags(&h->ev_fork[a], cudaEventDisableTiming));
        for (int g = 0; g < gprn_handle::NAUX; g++) {
            CU(cudaStreamCreateWithFlags(&h->aux[a][g], cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&h->ev_join[a][g], cudaEventDisableTiming));
        }
    }
    CU(cudaMalloc(&h->d_time, sizeof(double) * h->Np));
    CU(cudaMalloc(&h->d_y, sizeof(double) * p * N));
    CU(cudaMalloc(&h->d_yerr2, sizeof(double) * p * N));
    CU(cudaMalloc(&h->d_ysub_shared, sizeof(double) * p * N));
    std::vector<double> tp(h->Np, 0.0), e2((size_t)p * N);
    std::copy(time, time + N, tp.begin());
    for (size_t i = 0; i < (size_t)p * N; i++) e2[i] = yerr[i] * yerr[i];      // meanfield.py:127
    CU(cudaMemcpy(h->d_time, tp.data(), sizeof(double) * h->Np, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->d_y, y, sizeof(double) * p * N, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->d_yerr2, e2.data(), sizeof(double) * p * N, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->d_ysub_shared, y, sizeof(double) * p * N, cudaMemcpyHostToDevice));
    CU(cudaDeviceSynchronize());   // pageable-memory copies above must have landed before any non-blocking stream runs
    h->all = {&h->K, &h->W, &h->X, &h->XK, &h->vecs, &h->state, &h->small, &h->lists, &h->rlist, &h->hyper, &h->ysub,
              &h->ks, &h->pred, &h->scratch, &h->scratch2, &h->gpart, &h->res, &h->kscratch, &h->mid_state, &h->mid_zp, &h->mid_ld,
              &h->chain.mu, &h->chain.var, &h->chain.valid, &h->tmp_chain.mu, &h->tmp_chain.var, &h->tmp_chain.valid};
    *out = h;
    return 0;
}

extern "C" int gprn_destroy(gprn_handle* h) {
    if (!h) return 0;
    if (g_profile) {
        std::lock_guard<std::mutex> lk(g_mutex);
        double tot = 0;
        for (auto& kv : g_prof) tot += kv.second.first;
        for (auto& kv : g_prof)
            fprintf(stderr, "[gprn profile] line %4d: %9.1f ms  %7ld launches  %5.1f %%\n", kv.first, kv.second.first * 1e-3,
                    kv.second.second, 100.0 * kv.second.first / tot);
        g_prof.clear();
    }
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    drop_graphs(h);
    for (DevBuf* b : h->all)
        if (b->p) cudaFree(b->p);
    cudaFree(h->d_time); cudaFree(h->d_y); cudaFree(h->d_yerr2); cudaFree(h->d_ysub_shared);
    if (h->d_tok) cudaFree(h->d_tok);
    if (h->d_len) cudaFree(h->d_len);
    if (h->d_par_off) cudaFree(h->d_par_off);
    if (h->h_lists) cudaFreeHost(h->h_lists);
    if (h->h_active) cudaFreeHost(h->h_active);
    if (h->h_ret) cudaFreeHost(h->h_ret);
    cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1);
    cudaEventDestroy(h->ev_side_fork); cudaEventDestroy(h->ev_side_join);
    for (int a = 0; a < 2; a++) {
        cudaEventDestroy(h->ev_fork[a]);
        for (int g = 0; g < gprn_handle::NAUX; g++) { cudaStreamDestroy(h->aux[a][g]); cudaEventDestroy(h->ev_join[a][g]); }
    }
    cudaStreamDestroy(h->side);
    cudaStreamDestroy(h->cross);
    cudaEventDestroy(h->ev_cross_fork); cudaEventDestroy(h->ev_cross_join);
    cudaStreamDestroy(h->own_stream);
    delete h;
    return 0;
}

extern "C" int gprn_set_workspace_limit(gprn_handle* h, uint64_t bytes) {
    if (!h) return fail("null handle");
    h->ws_limit = bytes;
    return 0;
}

extern "C" int gprn_set_max_slots(gprn_handle* h, int slots) {
    if (!h) return fail("null handle");
    if (slots < 0) return fail("gprn_set_max_slots: slots must be >= 0 (0 = as many as fit)");
    h->max_slots = slots;
    return 0;
}

static int op_npar(int op) {
    switch (op) {
        case GPRN_OP_SE: return 2;
        case GPRN_OP_PER: return 3;
        case GPRN_OP_QP: return 4;
        case GPRN_OP_RQ: return 3;
        case GPRN_OP_M32: return 2;
        case GPRN_OP_M52: return 2;
        case GPRN_OP_WN: return 1;
        case GPRN_OP_CONST: return 1;
        case GPRN_OP_RQP: return 5;
        case GPRN_OP_COS: return 2;
        case GPRN_OP_EXP: return 2;
        case GPRN_OP_DSE: return 2;
        case GPRN_OP_DPER: return 3;
        case GPRN_OP_DQP: return 4;
        case GPRN_OP_GEXP: return 3;
        case GPRN_OP_PIECE: return 1;
        case GPRN_OP_PAC: return 3;
        case GPRN_OP_NPER: return 4;
        case GPRN_OP_QNPER: return 5;
        case GPRN_OP_COSP: return 3;
        case GPRN_OP_QCOSP: return 4;
        case GPRN_OP_ADD: case GPRN_OP_MUL: return 0;
        default: return -1;
    }
}
// validates a postfix program; returns number of parameters or -1
static int check_prog(const int32_t* tok, int n) {
    if (n < 1 || n > GPRN_MAX_PROG) return -1;
    int depth = 0, npar = 0;
    for (int i = 0; i < n; i++) {
        int k = op_npar(tok[i]);
        if (k < 0) return -1;
        if (tok[i] == GPRN_OP_ADD || tok[i] == GPRN_OP_MUL) {
            if (depth < 2) return -1;
            depth--;
        } else {
            depth++;
            if (depth > 6) return -1;
            npar += k;
        }
    }
    if (depth != 1 || npar > GPRN_MAX_PROG * 4) return -1;
    return npar;
}

extern "C" int gprn_set_model(gprn_handle* h, const int32_t* node_prog, const int32_t* node_prog_off,
                              const int32_t* weight_prog, const int32_t* weight_prog_off, int n_hyper) {
    if (!h || !node_prog || !node_prog_off || !weight_prog || !weight_prog_off) return fail("gprn_set_model: null argument");
    CU(cudaSetDevice(h->device));
    const int M = h->M, q = h->q, qp = h->q * h->p;
    h->h_tok.assign((size_t)M * GPRN_MAX_PROG, 0);
    h->h_len.assign(M, 0);
    h->h_par_off.assign(M, 0);
    h->h_npar.assign(M, 0);
    int off = 0;
    for (int m = 0; m < M; m++) {
        const int32_t* src;
        int n;
        if (m < q) { src = node_prog + node_prog_off[m]; n = node_prog_off[m + 1] - node_prog_off[m]; }
        else { int k = m - q; src = weight_prog + weight_prog_off[k]; n = weight_prog_off[k + 1] - weight_prog_off[k]; }
        int npar = check_prog(src, n);
        if (npar < 0) return fail("gprn_set_model: malformed kernel program for component " + std::to_string(m));
        for (int t = 0; t < n; t++) h->h_tok[(size_t)m * GPRN_MAX_PROG + t] = src[t];
        h->h_len[m] = n;
        h->h_par_off[m] = off;
        h->h_npar[m] = npar;
        off += npar;
    }
    (void)qp;
    if (off + h->p != n_hyper)
        return fail("gprn_set_model: n_hyper (" + std::to_string(n_hyper) + ") != kernel parameters (" +
                    std::to_string(off) + ") + p jitters");
    h->H = n_hyper;
    if (!h->d_tok) {
        CU(cudaMalloc(&h->d_tok, sizeof(int32_t) * M * GPRN_MAX_PROG));
        CU(cudaMalloc(&h->d_len, sizeof(int32_t) * M));
        CU(cudaMalloc(&h->d_par_off, sizeof(int32_t) * M));
    }
    CU(cudaMemcpy(h->d_tok, h->h_tok.data(), sizeof(int32_t) * M * GPRN_MAX_PROG, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->d_len, h->h_len.data(), sizeof(int32_t) * M, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->d_par_off, h->h_par_off.data(), sizeof(int32_t) * M, cudaMemcpyHostToDevice));
    CU(cudaDeviceSynchronize());
    h->model_set = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// batched factorisation driver
// ------------------------------------------------------------------------------------------------
// Np: padded order of the listed matrices (the handle's, except for the gprn_debug_* hooks).
// force_fused: -1 choose the panel step by the work in flight, 0 two launches (potrf + trsm), 1 one fused launch.
static int factor_batch(gprn_handle* h, int Np, double* W, const int* d_ids, int nmat, double* logdet, int* mstatus,
                        int* ctr /* per-matrix tickets, zero between launches */, double* X /* null: no inverse */,
                        cudaStream_t st, int nmat_concurrent = 0 /* matrices in flight on all streams */,
                        int force_fused = -1) {
    if (nmat_concurrent < nmat) nmat_concurrent = nmat;
    double* Gp = (double*)h->gpart.p;     // split-K partials of the inverse (two-level path), indexed by matrix id
    const int nt = Np / NB;
    const bool two = use_two_level(Np);
    // Cholesky: panels of 4 tile columns, left-looking inside the panel, right-looking trailing update per panel
    for (int k0 = 0; k0 < nt; k0 += 4) {
        const int ke = std::min(k0 + 4, nt);
        for (int k = k0; k < ke; k++) {
            const int n = nt - k - 1;
            // One launch per panel step (fused: shortest dependency chain) while the GPU is latency bound; two
            // launches (potrf + 4-tile trsm: a third of the SM time) once enough matrices are in flight for the
            // SMs to be the bottleneck.  Measured crossover (B200): nmat * nt^2 ~ 65536.  GPRN_FUSED_PANEL=0/1 forces.
            static const int fused_env = getenv("GPRN_FUSED_PANEL") ? atoi(getenv("GPRN_FUSED_PANEL")) : -1;
            const int force = force_fused >= 0 ? force_fused : fused_env;
            const bool fused_panel = force >= 0 ? force != 0 : (size_t)nmat_concurrent * nt * nt < 65536;
            if (fused_panel) {
                launch_hi(panel_col_kernel, dim3(std::max(1, (n + 1) / 2), nmat), dim3(256), PANEL_SMEM, st, W, d_ids, Np, k, k0, logdet, mstatus, ctr);
                LAUNCH_CHECK(h);
            } else {
                launch_hi(potrf_col_kernel, dim3(nmat), dim3(256), POTRF_COL_SMEM, st, W, d_ids, Np, k, k0, logdet, mstatus);
                LAUNCH_CHECK(h);
                if (n > 0) {
                    static const int trsm_groups = getenv("GPRN_TRSM_GROUPS") ? atoi(getenv("GPRN_TRSM_GROUPS")) : 1;
                    if (trsm_groups == 2)
                        launch_hi(trsm_col_kernel<2>, dim3((n + 3) / 4, nmat), dim3(256), TRSM_COL_SMEM(2), st, W, d_ids, Np, k, k0);
                    else
                        launch_hi(trsm_col_kernel<1>, dim3((n + 1) / 2, nmat), dim3(128), TRSM_COL_SMEM(1), st, W, d_ids, Np, k, k0);
                    LAUNCH_CHECK(h);
                }
            }
        }
        if (ke < nt) {
            if (two) {
                // lazy trailing update (see syrk_outer_kernel): panel index pidx = k0 / 4
                static const bool eager = getenv("GPRN_EAGER_SYRK") != nullptr;
                const int pidx = k0 / 4, t0 = ke * NB, n128 = (Np - t0) / 128;
                if (eager) {
                    syrk_outer_kernel<<<dim3(G_SPLIT * n128 * (n128 + 1) / 2, nmat), G_THREADS, GEMM128_SMEM, st>>>(W, d_ids, Np, k0 * NB, OUTER_KB, t0, 0);
                } else if (pidx % 2 == 0) {
                    syrk_outer_kernel<<<dim3(G_SPLIT * (2 * n128 - 1), nmat), G_THREADS, GEMM128_SMEM, st>>>(W, d_ids, Np, k0 * NB, OUTER_KB, t0, 2);
                } else {
                    syrk_outer_kernel<<<dim3(G_SPLIT * n128 * (n128 + 1) / 2, nmat), G_THREADS, GEMM128_SMEM, st>>>(W, d_ids, Np, (k0 - 4) * NB, 2 * OUTER_KB, t0, 0);
                }
            } else {
                const int n = nt - ke;
                syrk_update_kernel<<<dim3(n * (n + 1) / 2, nmat), 128, 2 * TILE_SMEM, st>>>(W, d_ids, Np, k0, ke);
            }
            LAUNCH_CHECK(h);
        }
    }
    if (X) {
        launch_hi(trtri_diag_kernel, dim3(nt, nmat), dim3(64), TRTRI_DIAG_SMEM, st, X, W, d_ids, Np);
        LAUNCH_CHECK(h);
        if (two || nt <= 4) {
            for (int i0 = 0; i0 < nt; i0 += 4) {
                int kc = TRTRI_KC;
                if (i0 > 0 && two) {
                    // The K range of a tile is always cut into chunks of TRTRI_KC: the chunking fixes the summation
                    // order of the inverse, so it must not depend on how many matrices happen to be in flight -- a
                    // set's result is bit-identical whatever batch it is evaluated in (round 1 skipped the split on
                    // a full GPU for ~1 %: not worth a batch-dependent rounding).  GPRN_NO_SPLITK=1: never split.
                    static const bool no_split = getenv("GPRN_NO_SPLITK") != nullptr;
                    kc = no_split ? std::max(i0 * NB, TRTRI_KC) : TRTRI_KC;
                    const int units = trtri_outer_units(i0 * NB, kc);
                    trtri_outer_kernel<<<dim3((OUTER_KB / G_BM) * units, nmat), G_THREADS, GEMM128_SMEM, st>>>(X, W, Gp, d_ids, Np, i0 * NB, units, kc);
                    LAUNCH_CHECK(h);
                }
                const int ncol = std::min(i0 + 4, nt) - 1;
                if (ncol > 0) {
                    launch_hi(trtri_inblock_kernel, dim3(ncol, nmat), dim3(128), TRTRI_SMEM, st, X, W, two ? Gp : nullptr, d_ids, Np, i0, kc);
                    LAUNCH_CHECK(h);
                }
            }
        } else {
            for (int i = 1; i < nt; i++) {
                trtri_row_kernel<<<dim3(i, nmat), 128, TRTRI_SMEM, st>>>(X, W, d_ids, Np, i);
                LAUNCH_CHECK(h);
            }
        }
    }
    return 0;
}

// Large matrices: the factorisation of one matrix alternates between wide GEMM launches and narrow, latency-
// bound panel launches.  Independent matrices are therefore split into groups that run the whole pipeline on
// concurrent streams, so that one group's panel steps overlap another group's trailing updates.
// aux_set: which of the handle's two stream-group sets to fork onto (0: iteration, 1: set-up running beside it).
static int factor_batch_multi(gprn_handle* h, double* W, const int* d_ids, int nmat, double* logdet, int* mstatus,
                              int* ctr, double* X, cudaStream_t st, int aux_set = 0, int nmat_other = 0) {
    static const int groups_env = getenv("GPRN_FACTOR_GROUPS") ? atoi(getenv("GPRN_FACTOR_GROUPS")) : 0;
    int G = groups_env > 0 ? groups_env : 8;     // measured on the C4 bench: 4 -> 8 groups +0.9 %
    if (!use_two_level(h->Np) || nmat < 2 || G < 2)
        return factor_batch(h, h->Np, W, d_ids, nmat, logdet, mstatus, ctr, X, st, nmat + nmat_other);
    G = std::min(std::min(G, (int)gprn_handle::NAUX), nmat);
    CU(cudaEventRecord(h->ev_fork[aux_set], st));
    int start = 0;
    for (int g = 0; g < G; g++) {
        const int len = nmat / G + (g < nmat % G ? 1 : 0);
        CU(cudaStreamWaitEvent(h->aux[aux_set][g], h->ev_fork[aux_set], 0));
        if (factor_batch(h, h->Np, W, d_ids + start, len, logdet, mstatus, ctr, X, h->aux[aux_set][g], nmat + nmat_other)) return 1;
        CU(cudaEventRecord(h->ev_join[aux_set][g], h->aux[aux_set][g]));
        CU(cudaStreamWaitEvent(st, h->ev_join[aux_set][g], 0));
        start += len;
    }
    return 0;
}

// z = X v ; u = X^T z ; g = colnorm2(X)   for the listed matrices
static int solve_batch(gprn_handle* h, const double* X, const int* d_ids, int nmat, double* vv, double* zv,
                       double* uv, double* gv, cudaStream_t st) {
    const int Np = h->Np, nt = h->nt;
    trmv_lower_kernel<<<dim3(Np / 8, nmat), 256, 0, st>>>(zv, X, vv, d_ids, nullptr, Np);
    LAUNCH_CHECK(h);
    trmv_upper_norm_kernel<<<dim3(nt, nmat), 256, 0, st>>>(uv, gv, X, zv, d_ids, Np);
    LAUNCH_CHECK(h);
    return 0;
}

static int small_batch(gprn_handle* h, const double* K, const int* d_ids, int nmat, const double* dvec, const double* vv,
                       double* uv, double* gv, double* logdet, int* mstatus, int do_inverse, double* scratch,
                       cudaStream_t st) {
    SmallArgs a;
    a.K = K; a.ids = d_ids; a.nmat = nmat; a.Np = h->Np; a.dvec = dvec; a.vv = vv;
    a.scratch = scratch; a.uv = uv; a.gv = gv; a.logdet = logdet; a.mstatus = mstatus;
    a.do_inverse = do_inverse;
    // GPRN_SMALL_CTAS=1 (experiment): one persistent CTA per SM instead of two -- what co-residency buys
    static const bool one_per_sm = getenv("GPRN_SMALL_CTAS") && atoi(getenv("GPRN_SMALL_CTAS")) == 1;
    // GPRN_SMALL_NO_TMA=1: tile loads as per-thread cp.async copies instead of TMA bulk copies (A/B switch)
    static const bool no_tma = getenv("GPRN_SMALL_NO_TMA") != nullptr;
    const int grid = one_per_sm ? std::min(nmat, h->num_sms) : std::min(nmat, SMALL_CTAS_PER_SM * h->num_sms);
    const size_t smem = SMALL_SMEM + (one_per_sm ? 8192 : 0);
    if (no_tma) small_pipeline_kernel<false><<<grid, 256, smem, st>>>(a);
    else small_pipeline_kernel<true><<<grid, 256, smem, st>>>(a);
    LAUNCH_CHECK(h);
    return 0;
}

static int mid_batch(gprn_handle* h, const double* K, double* W, double* X, const int* d_ids, int nmat, const double* dvec,
                     const double* vv, double* uv, double* gv, double* logdet, int* mstatus, int do_inverse,
                     cudaStream_t st, int aux_set = 0 /* 1: the set-up stream running beside an iteration */) {
    MidArgs a;
    a.K = K; a.W = W; a.X = X; a.ids = d_ids; a.Np = h->Np; a.dvec = dvec; a.vv = vv; a.uv = uv; a.gv = gv;
    a.logdet = logdet; a.mstatus = mstatus; a.do_inverse = do_inverse;
    a.tstate = (int*)h->mid_state.p;
    a.zp = (double*)h->mid_zp.p;
    a.ldpart = (double*)h->mid_ld.p;
    // one CTA per SM while they all fit (256 threads: potrf64 on eight warps), else two 128-thread CTAs per SM
    static const int force_nw = getenv("GPRN_MID_WARPS") ? atoi(getenv("GPRN_MID_WARPS")) : 0;
    const bool wide = force_nw ? force_nw == 8 : h->nt * nmat <= h->num_sms;
    // CTAs are numbered by start order, not by block index: needed when a launch has more CTAs than are resident at
    // once (throughput mode), and used always -- a set-up launch may share the GPU with an iteration launch, and a
    // CTA must only ever wait for CTAs that are running (one atomic per CTA, < 1 us per launch)
    a.ticket = h->mid_ticket + aux_set;
    if (wide) mid_pipeline_kernel<8><<<dim3(h->nt, nmat), 256, MID_SMEM, st>>>(a);
    else mid_pipeline_kernel<4><<<dim3(h->nt, nmat), 128, MID_SMEM, st>>>(a);
    LAUNCH_CHECK(h);
    mid_finish_kernel<<<dim3(h->nt, nmat), 256, 0, st>>>(a);
    LAUNCH_CHECK(h);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// the engine: `nslot` workspace slots, each holding one evaluation in flight
// ------------------------------------------------------------------------------------------------
struct Engine {          // plain data: compared bytewise to key the cached iteration graphs
    int nslot;
    ElboCtx c;
    double *K, *W, *X, *XK;
    // device lists (ints): active slots and their matrix ids, freshly admitted slots and their matrix ids,
    // slot -> pool index; retired slots live in their own buffer (uploaded after the poll)
    int *d_sets, *d_idn, *d_idw, *d_ida, *d_fsets, *d_fida, *d_slot_set, *d_rsets, *d_ctr;
    size_t nlist;          // ints in the main list block
};

static size_t per_set_bytes(const gprn_handle* h, bool need_factors) {
    const size_t Np = h->Np, M = h->M;
    const bool small = use_small_path(h) && !need_factors;
    size_t mats = (small ? 1 : (h->q > 1 ? 4 : 3)) * M * Np * Np * sizeof(double);
    size_t vecs = 8 * M * Np * sizeof(double);
    size_t state = 4 * (size_t)h->d * sizeof(double);
    size_t misc = 4096;
    if (use_mid_path(h))      // tile flags, z partials [column][row tile][64], log-det partials per matrix (setup_engine)
        misc += M * (MID_TILES * sizeof(int) + ((size_t)MID_MAX_NT * MID_MAX_NT * NB + MID_MAX_NT * 32) * sizeof(double));
    if (!small && use_two_level(h->Np))
        misc += M * (size_t)(TRTRI_MAXCH(h->Np) - 1) * OUTER_KB * Np * sizeof(double);      // split-K partials of the inverse
    return mats + vecs + state + misc;
}

// Sizes the workspace for `nslot` slots and lays the context out over it.  need_factors: the caller reads L / X
// from HBM (prediction, prior draws), so the fused small path does not apply.
static int setup_engine(gprn_handle* h, int nslot, Engine& E, bool need_factors = false) {
    const size_t Np = h->Np, M = h->M;
    const size_t matbytes = (size_t)nslot * M * Np * Np * sizeof(double);
    if (ensure(h->K, matbytes)) return 1;
    if (use_small_path(h) && !need_factors) {
        // fused path: factors never reach HBM; a 320 KB scratch per persistent CTA instead (one set of scratch
        // tiles for the iteration kernel, one for the set-up kernel that may run beside it)
        if (ensure(h->scratch, (size_t)SMALL_CTAS_PER_SM * h->num_sms * SMALL_SCRATCH_DOUBLES * sizeof(double))) return 1;
        if (ensure(h->scratch2, (size_t)SMALL_CTAS_PER_SM * h->num_sms * SMALL_SCRATCH_DOUBLES * sizeof(double))) return 1;
    } else {
        if (ensure(h->W, matbytes)) return 1;
        // the inverse factors must hold zeros in their (never written) upper tiles: trtri_outer_kernel reads them
        if (ensure_zeroed(h->X, matbytes)) return 1;
        if (h->q > 1 && ensure_zeroed(h->XK, matbytes)) return 1;
        if (use_two_level(h->Np) && TRTRI_MAXCH(h->Np) > 1 &&
            ensure(h->gpart, (size_t)nslot * M * (TRTRI_MAXCH(h->Np) - 1) * OUTER_KB * Np * sizeof(double))) return 1;
    }
    if (use_mid_path(h)) {
        // per matrix: tile flags, the z partials [column][row tile][64], the log-det partials [row][lane]
        const size_t nm_ = (size_t)nslot * M;
        // flags (zero between launches: mid_finish_kernel) + the two start-order tickets (iteration / set-up stream)
        if (ensure_zeroed(h->mid_state, (nm_ * MID_TILES + 2) * sizeof(int))) return 1;
        h->mid_ticket = (int*)h->mid_state.p + nm_ * MID_TILES;
        if (ensure(h->mid_zp, nm_ * (size_t)MID_MAX_NT * MID_MAX_NT * NB * sizeof(double))) return 1;
        if (ensure(h->mid_ld, nm_ * MID_MAX_NT * 32 * sizeof(double))) return 1;
    }
    const size_t ve = (size_t)nslot * M * Np;
    if (ensure(h->vecs, 7 * ve * sizeof(double))) return 1;
    if (ensure(h->state, 4 * (size_t)nslot * h->d * sizeof(double))) return 1;
    // small: logdetK, logdetA, ment, mlp, mquad [nslot*M]; cross_lin [nslot]; crossbuf [nslot*npairs*nt*nt];
    //        hist [nslot*3]; elbo [nslot] doubles; then ints
    const size_t ncross = (size_t)(h->q * (h->q - 1) / 2) * h->nt * h->nt;
    const size_t nd = 5 * (size_t)nslot * M + (size_t)nslot * (1 + ncross + 3 + 1);
    const size_t ni = 3 * (size_t)nslot + 2 * (size_t)nslot * M;
    if (ensure_zeroed(h->small, nd * sizeof(double) + ni * sizeof(int))) return 1;
    const size_t nl = (size_t)nslot * (3 + 3 * M);
    if (ensure(h->lists, nl * sizeof(int))) return 1;
    if (ensure(h->rlist, (size_t)nslot * sizeof(int))) return 1;
    if (ensure_pinned(h->h_lists, h->h_lists_n, nl)) return 1;
    if (ensure_pinned(h->h_active, h->h_active_n, 3 * (size_t)nslot)) return 1;
    if (ensure_pinned(h->h_ret, h->h_ret_n, nslot)) return 1;

    E.nslot = nslot;
    E.nlist = nl;
    E.K = (double*)h->K.p; E.W = (double*)h->W.p; E.X = (double*)h->X.p; E.XK = (double*)h->XK.p;
    ElboCtx& c = E.c;
    c.N = h->N; c.Np = h->Np; c.p = h->p; c.q = h->q; c.M = h->M; c.H = h->H; c.d = h->d;
    c.yraw = h->d_y; c.yerr2 = h->d_yerr2;
    c.ysub = h->d_ysub_shared;
    c.ysub_shared = 1;
    c.hyper = nullptr;
    c.par_off = h->d_par_off;
    double* v = (double*)h->vecs.p;
    c.Dv = v; c.bv = v + ve; c.vv = v + 2 * ve; c.zv = v + 3 * ve; c.uv = v + 4 * ve; c.gv = v + 5 * ve; c.gK = v + 6 * ve;
    double* s = (double*)h->state.p;
    const size_t sd = (size_t)nslot * h->d;
    c.mu = s; c.var = s + sd; c.mu_new = s + 2 * sd; c.var_new = s + 3 * sd;
    double* sm = (double*)h->small.p;
    const size_t nm = (size_t)nslot * M;
    c.logdetK = sm; c.logdetA = sm + nm; c.ment = sm + 2 * nm; c.mlp = sm + 3 * nm; c.mquad = sm + 4 * nm;
    c.cross_lin = sm + 5 * nm;
    c.crossbuf = c.cross_lin + nslot;
    c.hist = c.crossbuf + (size_t)nslot * ncross;
    c.elbo = c.hist + (size_t)nslot * 3;
    int* si = (int*)(c.elbo + nslot);
    c.iters = si; c.status = si + nslot; c.active = si + 2 * nslot; c.mstatus = si + 3 * nslot;
    E.d_ctr = si + 3 * nslot + nm;      // tickets of panel_col_kernel: zero at allocation, self-resetting
    int* l = (int*)h->lists.p;
    E.d_sets = l;
    E.d_idn = l + nslot;
    E.d_idw = E.d_idn + (size_t)nslot * h->q;
    E.d_ida = l + nslot + (size_t)nslot * M;
    E.d_fsets = E.d_ida + (size_t)nslot * M;
    E.d_fida = E.d_fsets + nslot;
    E.d_slot_set = E.d_fida + (size_t)nslot * M;
    E.d_rsets = (int*)h->rlist.p;
    c.slot_set = E.d_slot_set;
    c.max_iter = 0;
    return 0;
}

// Builds the launch lists (active slots, fresh slots, slot -> pool index) in pinned memory and uploads them.
static int upload_lists(gprn_handle* h, Engine& E, const std::vector<int>& act, const std::vector<int>& fresh,
                        const std::vector<int>& slot_set, cudaStream_t st) {
    const int q = h->q, p = h->p, M = h->M, S = E.nslot;
    int* L = h->h_lists;
    int* sets = L;
    int* idn = L + S;
    int* idw = idn + (size_t)S * q;
    int* ida = L + S + (size_t)S * M;
    int* fsets = ida + (size_t)S * M;
    int* fida = fsets + S;
    int* sset = fida + (size_t)S * M;
    for (int a = 0; a < (int)act.size(); a++) {
        const int s = act[a];
        sets[a] = s;
        for (int j = 0; j < q; j++) idn[(size_t)a * q + j] = s * M + j;
        for (int k = 0; k < q * p; k++) idw[(size_t)a * q * p + k] = s * M + q + k;
        for (int m = 0; m < M; m++) ida[(size_t)a * M + m] = s * M + m;
    }
    for (int a = 0; a < (int)fresh.size(); a++) {
        const int s = fresh[a];
        fsets[a] = s;
        for (int m = 0; m < M; m++) fida[(size_t)a * M + m] = s * M + m;
    }
    for (int s = 0; s < S; s++) sset[s] = slot_set[s];
    CU(cudaMemcpyAsync(h->lists.p, L, sizeof(int) * E.nlist, cudaMemcpyHostToDevice, st));
    return 0;
}

// Set-up of the `nf` freshly admitted slots (lists in E.d_fsets / E.d_fida): initial state, K_m, chol(K_m) with its
// log-det (and inverse factor + diag(K^-1) when q > 1: the cross-node terms need them).
static int launch_setup(gprn_handle* h, Engine& E, int nf, ChainView cs, cudaStream_t st, int aux_set, int nmat_other) {
    ElboCtx& c = E.c;
    const int q = h->q, M = h->M, Np = h->Np, nt = h->nt, ntri = nt * (nt + 1) / 2;
    init_state_kernel<<<nf, 256, 0, st>>>(c, E.d_fsets, cs);
    LAUNCH_CHECK(h);
    ProgTable pt{h->d_tok, h->d_len, h->d_par_off};
    kassemble_sym_kernel<<<dim3(ntri, M, nf), 256, 0, st>>>(E.K, h->d_time, c.hyper, h->H, pt, M, h->N, Np, 1e-6, E.d_fsets, E.d_slot_set);
    LAUNCH_CHECK(h);
    if (use_small_path(h)) {
        double* scr = (double*)(aux_set ? h->scratch2.p : h->scratch.p);
        if (small_batch(h, E.K, E.d_fida, nf * M, nullptr, nullptr, nullptr, nullptr, c.logdetK, c.mstatus, 0, scr, st)) return 1;
    } else if (use_mid_path(h)) {
        // q > 1: the cross-node terms and the prior's quadratic forms need L_K^-1 (transposed tiles) and diag(K^-1)
        if (mid_batch(h, E.K, E.W, q > 1 ? E.XK : E.X, E.d_fida, nf * M, nullptr, nullptr, nullptr, q > 1 ? c.gK : nullptr,
                      c.logdetK, c.mstatus, q > 1 ? 1 : 0, st, aux_set)) return 1;
    } else {
        form_a_kernel<<<dim3(ntri, nf * M), 256, 0, st>>>(E.W, E.K, nullptr, E.d_fida, Np);
        LAUNCH_CHECK(h);
        if (factor_batch_multi(h, E.W, E.d_fida, nf * M, c.logdetK, c.mstatus, E.d_ctr, q > 1 ? E.XK : nullptr, st, aux_set, nmat_other)) return 1;
        if (q > 1) {
            trmv_upper_norm_kernel<<<dim3(nt, nf * M), 256, 0, st>>>(nullptr, c.gK, E.XK, nullptr, E.d_fida, Np);
            LAUNCH_CHECK(h);
        }
    }
    return 0;
}

// One lock-step fixed-point iteration (meanfield.py:634-646, ELBOaux :651-710) of the `na` active slots listed in
// E.d_sets: node phase, cross-node terms, weight phase, ELBO + stopping rule.  nmat_other: matrices a concurrent
// set-up has in flight (only steers the panel-step heuristics).
static int launch_iteration(gprn_handle* h, Engine& E, int na, cudaStream_t st, int nmat_other,
                            cudaGraphConditionalHandle loop_cond = 0) {
    ElboCtx& c = E.c;
    const int q = h->q, p = h->p, M = h->M, Np = h->Np, nt = h->nt, ntri = nt * (nt + 1) / 2;
    const bool small = use_small_path(h);
    double* scr = (double*)h->scratch.p;
    CU(cudaMemsetAsync(c.logdetA, 0, sizeof(double) * (size_t)E.nslot * M, st));
    // node phase
    prep_nodes_kernel<<<dim3(q, na), vec_threads(h), 0, st>>>(c, E.d_sets);
    LAUNCH_CHECK(h);
    if (small) {
        if (small_batch(h, E.K, E.d_idn, na * q, c.Dv, c.vv, c.uv, c.gv, c.logdetA, c.mstatus, 1, scr, st)) return 1;
    } else if (use_mid_path(h)) {
        if (mid_batch(h, E.K, E.W, E.X, E.d_idn, na * q, c.Dv, c.vv, c.uv, c.gv, c.logdetA, c.mstatus, 1, st)) return 1;
    } else {
        form_a_kernel<<<dim3(ntri, na * q), 256, 0, st>>>(E.W, E.K, c.Dv, E.d_idn, Np);
        LAUNCH_CHECK(h);
        if (factor_batch_multi(h, E.W, E.d_idn, na * q, c.logdetA, c.mstatus, E.d_ctr, E.X, st, 0, nmat_other)) return 1;
        if (solve_batch(h, E.X, E.d_idn, na * q, c.vv, c.zv, c.uv, c.gv, st)) return 1;
    }
    post_kernel<<<dim3(q, na), vec_threads(h), 0, st>>>(c, E.d_sets, 0, q == 1);
    LAUNCH_CHECK(h);
    if (q > 1 && use_mid_path(h)) {
        // latency path: small matrices, factors stored as transposed tiles; no extra stream (the whole iteration is
        // the body of a conditional graph node)
        cross_linear_kernel<<<na, 256, 0, st>>>(c, E.d_sets);
        LAUNCH_CHECK(h);
        cross_frob_kernel<true><<<dim3(nt * nt, q * (q - 1) / 2, na), 128, 2 * TILE_SMEM, st>>>(c, E.XK, E.X, E.d_sets);
        LAUNCH_CHECK(h);
    } else if (q > 1) {
        // Cross-node trace terms (quirk Q3): GEMM-class work that only the ELBO needs.  It runs on its own stream
        // beside the latency-bound start of the weight phase (it reads the node matrices' D and X, which the weight
        // phase does not touch) and is joined before elbo_finish_kernel.
        CU(cudaEventRecord(h->ev_cross_fork, st));
        CU(cudaStreamWaitEvent(h->cross, h->ev_cross_fork, 0));
        cross_linear_kernel<<<na, 256, 0, h->cross>>>(c, E.d_sets);
        LAUNCH_CHECK(h);
        cross_frob_kernel<false><<<dim3(nt * nt, q * (q - 1) / 2, na), 128, 2 * TILE_SMEM, h->cross>>>(c, E.XK, E.X, E.d_sets);
        LAUNCH_CHECK(h);
        CU(cudaEventRecord(h->ev_cross_join, h->cross));
    }
    // weight phase
    prep_weights_kernel<<<dim3(q * p, na), vec_threads(h), 0, st>>>(c, E.d_sets);
    LAUNCH_CHECK(h);
    if (small) {
        if (small_batch(h, E.K, E.d_idw, na * q * p, c.Dv, c.vv, c.uv, c.gv, c.logdetA, c.mstatus, 1, scr, st)) return 1;
    } else if (use_mid_path(h)) {
        if (mid_batch(h, E.K, E.W, E.X, E.d_idw, na * q * p, c.Dv, c.vv, c.uv, c.gv, c.logdetA, c.mstatus, 1, st)) return 1;
    } else {
        form_a_kernel<<<dim3(ntri, na * q * p), 256, 0, st>>>(E.W, E.K, c.Dv, E.d_idw, Np);
        LAUNCH_CHECK(h);
        if (factor_batch_multi(h, E.W, E.d_idw, na * q * p, c.logdetA, c.mstatus, E.d_ctr, E.X, st, 0, nmat_other)) return 1;
        if (solve_batch(h, E.X, E.d_idw, na * q * p, c.vv, c.zv, c.uv, c.gv, st)) return 1;
    }
    post_kernel<<<dim3(q * p, na), vec_threads(h), 0, st>>>(c, E.d_sets, q, q == 1);
    LAUNCH_CHECK(h);
    if (q > 1) {   // quadratic forms with the reference's vector pairing (quirk Q4)
        gather_quad_vec_kernel<<<dim3(M, na), 256, 0, st>>>(c, E.d_sets, 0);
        LAUNCH_CHECK(h);
        if (use_mid_path(h)) mid_trmv_lower_kernel<<<dim3(nt, na * M), 256, 0, st>>>(c.zv, E.XK, c.vv, E.d_ida, Np);
        else trmv_lower_kernel<<<dim3(Np / 8, na * M), 256, 0, st>>>(c.zv, E.XK, c.vv, E.d_ida, nullptr, Np);
        LAUNCH_CHECK(h);
        quad_kernel<<<dim3(M, na), 256, 0, st>>>(c, E.d_sets, 0);
        LAUNCH_CHECK(h);
    }
    if (q > 1 && !use_mid_path(h)) CU(cudaStreamWaitEvent(st, h->ev_cross_join, 0));
    elbo_finish_kernel<<<na, 1024, 0, st>>>(c, E.d_sets, loop_cond);
    LAUNCH_CHECK(h);
    return 0;
}

// The iteration LOOP as one CUDA graph (latency path, mid.cuh): a WHILE conditional node whose body is the captured
// iteration.  elbo_finish_kernel, the last kernel of the body, clears the node's condition as soon as ANY active set
// has finished (converged, max_iter, not positive definite), so the device runs lock-step iterations back to back
// with no host round trip until there is something to retire -- for a single evaluation: the whole fixed-point loop
// of ELBOcalc (meanfield.py:634-646) in one launch.  The host then polls, retires and refills exactly as after a
// single iteration.  GPRN_NO_LOOP=1 falls back to one graph launch + poll per iteration.
// Cached graphs hold raw workspace pointers, the context by value and the kernel choice of the call that captured them:
// they are valid for exactly this Engine and this path (the path can change under the same Engine when an experiment
// switch such as GPRN_NO_MID is toggled between calls).  Drops the cache when either differs.
static void check_graph_signature(gprn_handle* h, const Engine& E) {
    const unsigned char* eb = reinterpret_cast<const unsigned char*>(&E);
    const unsigned char mode[3] = {(unsigned char)h->mid_mode, (unsigned char)h->small_mode, (unsigned char)h->mid_latency};
    const size_t n = sizeof(Engine) + sizeof(mode);
    if (h->graph_sig.size() != n || memcmp(h->graph_sig.data(), eb, sizeof(Engine)) != 0 ||
        memcmp(h->graph_sig.data() + sizeof(Engine), mode, sizeof(mode)) != 0) {
        drop_graphs(h);
        h->graph_sig.assign(eb, eb + sizeof(Engine));
        h->graph_sig.insert(h->graph_sig.end(), mode, mode + sizeof(mode));
    }
}

static bool use_device_loop(const gprn_handle* h) {
    static const bool off = getenv("GPRN_NO_LOOP") != nullptr || getenv("GPRN_NO_GRAPH") != nullptr;
    return !off && !g_debug_sync && !g_profile && !h->loop_unavailable && use_mid_path(h) && h->mid_latency;
}
static int iteration_loop_graph(gprn_handle* h, Engine& E, int na, cudaStream_t st) {
    check_graph_signature(h, E);
    auto it = h->loop_graphs.find(na);
    if (it == h->loop_graphs.end()) {
        if (h->loop_graphs.size() >= 64) return 2;       // many distinct counts: per-iteration launches for the rest
        cudaGraph_t g = nullptr;
        CU(cudaGraphCreate(&g, 0));
        cudaGraphConditionalHandle cond;
        cudaError_t e = cudaGraphConditionalHandleCreate(&cond, g, 1, cudaGraphCondAssignDefault);   // 1 at every launch
        cudaGraphNodeParams np = {};
        np.type = cudaGraphNodeTypeConditional;
        np.conditional.handle = cond;
        np.conditional.type = cudaGraphCondTypeWhile;
        np.conditional.size = 1;
        cudaGraphNode_t node;
        if (e == cudaSuccess) e = cudaGraphAddNode(&node, g, nullptr, 0, &np);
        cudaGraph_t body = e == cudaSuccess ? np.conditional.phGraph_out[0] : nullptr;
        if (e == cudaSuccess) e = cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) {                    // not fatal: one graph launch + poll per iteration instead
            cudaGraphDestroy(g);
            cudaGetLastError();
            h->loop_unavailable = true;
            return 2;
        }
        h->capturing = true;
        const int64_t before = h->launches;
        int rc = launch_iteration(h, E, na, st, 0, cond);
        h->capturing = false;
        cudaGraph_t captured = nullptr;
        e = cudaStreamEndCapture(st, &captured);
        h->launches = before;                      // counted per executed round after the poll
        if (rc) { cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess) { cudaGraphDestroy(g); return fail(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e)); }
        size_t nn = 0;
        CU(cudaGraphGetNodes(body, nullptr, &nn));
        cudaGraphExec_t ge = nullptr;
        e = cudaGraphInstantiate(&ge, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) {
            cudaGetLastError();
            h->loop_unavailable = true;
            return 2;
        }
        it = h->loop_graphs.emplace(na, ge).first;
        h->loop_kernels[na] = (int64_t)nn;
    }
    CU(cudaGraphLaunch(it->second, st));
    h->graph_launches++;
    return 0;
}

// The iteration as a CUDA graph.  A lock-step iteration is the same launch sequence every time -- all arguments
// are workspace addresses, the slot lists are device arrays whose CONTENT changes -- and only its grid sizes depend
// on the number of active slots, so one graph per `na` is captured lazily (multi-stream fork / join included) and
// replayed: one host launch per iteration instead of hundreds to thousands (C4: ~2000).  GPRN_NO_GRAPH=1 disables.
static int iteration_graph(gprn_handle* h, Engine& E, int na, cudaStream_t st) {
    static const bool no_graph = getenv("GPRN_NO_GRAPH") != nullptr;
    // the fused small path is a handful of launches per iteration and its active count changes every round
    // (and so is the dataflow path in throughput mode)
    if (no_graph || g_debug_sync || g_profile || use_small_path(h) || (use_mid_path(h) && !h->mid_latency))
        return launch_iteration(h, E, na, st, 0);
    // cached graphs hold raw workspace pointers and the context by value: they are valid for exactly this Engine
    check_graph_signature(h, E);
    auto it = h->iter_graphs.find(na);
    if (it == h->iter_graphs.end()) {
        if (h->iter_graphs.size() >= 64) {          // many distinct counts (large pools of small sets): launch directly
            return launch_iteration(h, E, na, st, 0);
        }
        cudaGraph_t g = nullptr;
        CU(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        h->capturing = true;
        const int64_t before = h->launches;
        int rc = launch_iteration(h, E, na, st, 0);
        h->capturing = false;
        cudaError_t e = cudaStreamEndCapture(st, &g);
        h->launches = before;                      // counted per replay below
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess) return fail(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
        size_t nn = 0;
        CU(cudaGraphGetNodes(g, nullptr, &nn));
        cudaGraphExec_t ge = nullptr;
        e = cudaGraphInstantiate(&ge, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return fail(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
        it = h->iter_graphs.emplace(na, ge).first;
        h->graph_kernels[na] = (int64_t)nn;
    }
    CU(cudaGraphLaunch(it->second, st));
    h->graph_launches++;
    h->launches += h->graph_kernels[na];
    return 0;
}

struct PoolJob {
    int64_t B = 0;                    // size of the pool (index space of hyper / outputs / chain store)
    gprn_next_set_fn next = nullptr;  // work source (null: 0 .. B-1 in order)
    void* user = nullptr;
    int64_t cursor = 0;
    ChainView cs = {nullptr, nullptr, nullptr};
    int commit = 0;
    int max_iter = 0;
    double* d_elbo = nullptr;         // device, [B]
    int* d_iters = nullptr;
    int* d_status = nullptr;
    int* d_taken = nullptr;
};
static int64_t job_next(PoolJob& j) {
    if (j.next) return j.next(j.user);
    return j.cursor < j.B ? j.cursor++ : -1;
}

// Continuous batching over the pool.  Every pass of the loop (a "round") admits sets from the work source into the
// free slots, runs the set-up of the newly admitted slots on the side stream NEXT TO one lock-step iteration of the
// slots already active, reads the `active` flags (the one host round trip per round) and retires the slots whose
// evaluation has finished; their slots are refilled in the next round.  All launches of a phase are batched over
// the slots' matrix lists, so the steady state is the same wide launches as a fixed lock-step batch -- without
// its tail: a converged set is replaced instead of leaving its share of the GPU idle.
// The duplicate pre-loop ELBOaux of the reference (meanfield.py:627, quirk Q7) is not executed: its value equals
// iteration 1.
static int run_pool(gprn_handle* h, Engine& E, PoolJob& job, cudaStream_t st) {
    ElboCtx& c = E.c;
    const int S = E.nslot, M = h->M;
    c.max_iter = job.max_iter;
    if ((size_t)S * M > 65535) return fail("internal: too many slots for the launch grids");
    CU(cudaMemsetAsync(E.d_ctr, 0, sizeof(int) * (size_t)S * M, st));
    std::vector<int> slot_set(S, 0), slot_iters(S, 0), act, fresh, freeslots, next_act, retired;
    act.reserve(S); fresh.reserve(S); next_act.reserve(S); retired.reserve(S);
    for (int s = S - 1; s >= 0; s--) freeslots.push_back(s);
    bool empty = false, dirty = true;
    h->last_lockstep_rounds = 0;
    for (;;) {
        fresh.clear();
        while (!freeslots.empty() && !empty) {
            const int64_t idx = job_next(job);
            if (idx < 0) { empty = true; break; }
            if (idx >= job.B) return fail("gprn_elbo_pool: the work source returned an index outside the pool");
            const int s = freeslots.back();
            freeslots.pop_back();
            slot_set[s] = (int)idx;
            slot_iters[s] = 0;
            fresh.push_back(s);
        }
        const int na = (int)act.size(), nf = (int)fresh.size();
        if (na == 0 && nf == 0) break;
        if (dirty || nf) {
            if (upload_lists(h, E, act, fresh, slot_set, st)) return 1;
            dirty = false;
        }
        if (nf && na) {
            CU(cudaEventRecord(h->ev_side_fork, st));
            CU(cudaStreamWaitEvent(h->side, h->ev_side_fork, 0));
            if (launch_setup(h, E, nf, job.cs, h->side, 1, na * M)) return 1;
            CU(cudaEventRecord(h->ev_side_join, h->side));
        } else if (nf) {
            if (launch_setup(h, E, nf, job.cs, st, 0, 0)) return 1;
        }
        if (na) {
            bool looped = !nf && use_device_loop(h);
            if (looped) {
                const int rc = iteration_loop_graph(h, E, na, st);
                if (rc == 1) return 1;
                if (rc == 2) looped = false;
            }
            if (!looped && (nf ? launch_iteration(h, E, na, st, nf * M) : iteration_graph(h, E, na, st))) return 1;
            if (nf) CU(cudaStreamWaitEvent(st, h->ev_side_join, 0));
            // convergence poll: the one host round trip of a round (of a whole run of rounds with the device loop).
            // iters | status | active are contiguous: one copy brings the iteration counts along.
            CU(cudaMemcpyAsync(h->h_active, c.iters, sizeof(int) * 3 * S, cudaMemcpyDeviceToHost, st));
            if (use_mid_path(h) && h->mid_latency) {
                // latency path: the evaluation is a few milliseconds and the caller waits for it -- spin on the stream
                // instead of a blocking wait, whose wake-up after a multi-millisecond sleep costs up to 0.5 ms
                cudaError_t qe;
                while ((qe = cudaStreamQuery(st)) == cudaErrorNotReady) {}
                if (qe != cudaSuccess) return fail(std::string("cudaStreamQuery: ") + cudaGetErrorString(qe));
            } else {
                CU(cudaStreamSynchronize(st));
            }
            {
                const int s0 = act[0];
                const int64_t rounds = looped ? std::max(1, h->h_active[s0] - slot_iters[s0]) : 1;
                h->last_lockstep_rounds += rounds;
                if (looped) h->launches += rounds * h->loop_kernels[na];
                for (int s : act) slot_iters[s] = h->h_active[s];
            }
            const int* h_act = h->h_active + 2 * S;
            next_act.clear();
            retired.clear();
            for (int s : act) (h_act[s] ? next_act : retired).push_back(s);
            if (!retired.empty()) {
                const int nr = (int)retired.size();
                for (int a = 0; a < nr; a++) h->h_ret[a] = retired[a];
                CU(cudaMemcpyAsync(E.d_rsets, h->h_ret, sizeof(int) * nr, cudaMemcpyHostToDevice, st));
                retire_kernel<<<nr, 256, 0, st>>>(c, E.d_rsets, job.d_elbo, job.d_iters, job.d_status, job.d_taken, job.cs, job.commit);
                LAUNCH_CHECK(h);
                for (int s : retired) freeslots.push_back(s);
                dirty = true;
            }
            act.swap(next_act);
        }
        if (nf) {
            act.insert(act.end(), fresh.begin(), fresh.end());
            dirty = true;
        }
    }
    return 0;
}

static int chunk_size(gprn_handle* h, int64_t B, bool need_factors = false) {
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    // buffers already held by the handle count as available
    size_t held = 0;
    for (DevBuf* b : h->all) held += b->bytes;
    size_t budget = (size_t)((free_b + held) * 0.85);
    if (h->ws_limit && h->ws_limit < budget) budget = h->ws_limit;
    size_t per = per_set_bytes(h, need_factors);
    size_t n = budget / per;
    size_t grid_cap = 65535 / (size_t)h->M;          // matrix lists index blockIdx.y / .z
    n = std::min(n, grid_cap);
    n = std::min(n, (size_t)B);
    return (int)n;
}

static int chain_view(gprn_handle* h, ChainStore* cs, int64_t need, ChainView& v) {
    v = ChainView{nullptr, nullptr, nullptr};
    if (!cs) return 0;
    if (cs->n < need) return fail("chain-state store holds " + std::to_string(cs->n) + " chains, the pool has " +
                                  std::to_string(need) + " (call gprn_chain_resize)");
    v.mu = (double*)cs->mu.p; v.var = (double*)cs->var.p; v.valid = (int*)cs->valid.p;
    (void)h;
    return 0;
}
static int chain_resize(gprn_handle* h, ChainStore& cs, int64_t n) {
    if (ensure(cs.mu, sizeof(double) * (size_t)n * h->d)) return 1;
    if (ensure(cs.var, sizeof(double) * (size_t)n * h->d)) return 1;
    if (ensure(cs.valid, sizeof(int) * (size_t)n)) return 1;
    CU(cudaMemset(cs.valid.p, 0, sizeof(int) * (size_t)n));
    CU(cudaDeviceSynchronize());
    cs.n = n;
    return 0;
}

static int elbo_impl(gprn_handle* h, int64_t B, const double* hyper, bool hyper_on_device, const double* ysub_host,
                     int ysub_shared, gprn_next_set_fn next, void* user, int max_slots, ChainStore* cstore, int commit,
                     int max_iter, double* elbo_out, int32_t* iters_out, int32_t* status_out, int32_t* taken_out,
                     bool out_on_device, void* stream) {
    if (!h) return fail("null handle");
    if (!h->model_set) return fail("gprn_elbo_batched: call gprn_set_model first");
    if (B < 1) return fail("gprn_elbo_batched: B must be >= 1");
    if (B > 0x7fffffff) return fail("gprn_elbo_batched: pool too large");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->own_stream;
    if (max_iter < 0) max_iter = 10000;                       // meanfield.py:615-616
    {
        const int cap0 = max_slots > 0 ? max_slots : h->max_slots;
        const int64_t in_flight = cap0 > 0 ? std::min<int64_t>(B, cap0) : B;
        h->mid_mode = decide_mid_path(h, in_flight);
        h->mid_latency = h->mid_mode && mid_colocated(h, in_flight);
        h->small_mode = !h->mid_mode && decide_small_path(h, in_flight);
    }
    int nslot = chunk_size(h, B);
    if (nslot < 1) return fail("gprn_elbo_batched: not enough device memory for one evaluation of this size");
    const int cap = max_slots > 0 ? max_slots : h->max_slots;
    if (cap > 0) nslot = std::min(nslot, cap);
    const bool per_set_y = ysub_host && !ysub_shared;
    Engine E;
    memset(&E, 0, sizeof(E));
    if (setup_engine(h, nslot, E)) return 1;
    CU(cudaEventRecord(h->ev0, st));
    if (ysub_host && ysub_shared)
        CU(cudaMemcpyAsync(h->d_ysub_shared, ysub_host, sizeof(double) * h->p * h->N, cudaMemcpyHostToDevice, st));
    if (per_set_y) {
        if (ensure(h->ysub, (size_t)B * h->p * h->N * sizeof(double))) return 1;
        CU(cudaMemcpyAsync(h->ysub.p, ysub_host, sizeof(double) * (size_t)B * h->p * h->N, cudaMemcpyHostToDevice, st));
        E.c.ysub = (const double*)h->ysub.p;
        E.c.ysub_shared = 0;
    }
    if (hyper_on_device) {
        E.c.hyper = hyper;
    } else {
        if (ensure(h->hyper, (size_t)B * h->H * sizeof(double))) return 1;
        CU(cudaMemcpyAsync(h->hyper.p, hyper, sizeof(double) * (size_t)B * h->H, cudaMemcpyHostToDevice, st));
        E.c.hyper = (const double*)h->hyper.p;
    }
    PoolJob job;
    job.B = B; job.next = next; job.user = user; job.max_iter = max_iter; job.commit = commit;
    if (chain_view(h, cstore, B, job.cs)) return 1;
    // results: straight into the caller's device arrays, or staged in the handle and copied out at the end.
    // Entries of sets this call does not evaluate (another rank took them from a shared work source) stay zero.
    const size_t res_bytes = (size_t)B * (sizeof(double) + 3 * sizeof(int));
    if (ensure(h->res, res_bytes)) return 1;
    double* r_elbo = (double*)h->res.p;
    int* r_iters = (int*)(r_elbo + B);
    int* r_status = r_iters + B;
    int* r_taken = r_status + B;
    CU(cudaMemsetAsync(h->res.p, 0, res_bytes, st));
    if (out_on_device) {
        if (elbo_out) { CU(cudaMemsetAsync(elbo_out, 0, sizeof(double) * B, st)); r_elbo = elbo_out; }
        if (iters_out) { CU(cudaMemsetAsync(iters_out, 0, sizeof(int) * B, st)); r_iters = iters_out; }
        if (status_out) { CU(cudaMemsetAsync(status_out, 0, sizeof(int) * B, st)); r_status = status_out; }
        if (taken_out) { CU(cudaMemsetAsync(taken_out, 0, sizeof(int) * B, st)); r_taken = taken_out; }
    }
    job.d_elbo = r_elbo; job.d_iters = r_iters; job.d_status = r_status; job.d_taken = r_taken;
    if (run_pool(h, E, job, st)) return 1;
    if (!out_on_device) {
        if (elbo_out) CU(cudaMemcpyAsync(elbo_out, r_elbo, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
        if (iters_out) CU(cudaMemcpyAsync(iters_out, r_iters, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
        if (status_out) CU(cudaMemcpyAsync(status_out, r_status, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
        if (taken_out) CU(cudaMemcpyAsync(taken_out, r_taken, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
    }
    // iteration total for flop accounting
    std::vector<int> its(B);
    CU(cudaMemcpyAsync(its.data(), r_iters, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(h->ev1, st));
    CU(cudaEventSynchronize(h->ev1));
    h->last_total_iters = 0;
    for (int64_t b = 0; b < B; b++) h->last_total_iters += its[b];
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_ms = ms;
    return 0;
}

extern "C" int gprn_elbo_batched(gprn_handle* h, int B, const double* hyper, const double* ysub, int ysub_shared,
                                 int init_mode, double* mu_inout, double* var_inout, int max_iter, double* elbo_out,
                                 int32_t* iters_out, int32_t* status_out, void* stream) {
    if (!h) return fail("null handle");
    if (!hyper) return fail("gprn_elbo_batched: hyper is null");
    if (B < 1) return fail("gprn_elbo_batched: B must be >= 1");
    ChainStore* cs = nullptr;
    if (init_mode == 1 && (!mu_inout || !var_inout)) return fail("gprn_elbo_batched: init_mode=1 needs mu_inout and var_inout");
    if (mu_inout || var_inout) {
        // host state in / out: staged through a device store of B chains (valid entries = the given initial state)
        CU(cudaSetDevice(h->device));
        cs = &h->tmp_chain;
        if (chain_resize(h, *cs, B)) return 1;
        if (init_mode == 1) {
            CU(cudaMemcpy(cs->mu.p, mu_inout, sizeof(double) * (size_t)B * h->d, cudaMemcpyHostToDevice));
            CU(cudaMemcpy(cs->var.p, var_inout, sizeof(double) * (size_t)B * h->d, cudaMemcpyHostToDevice));
            std::vector<int> ones(B, 1);
            CU(cudaMemcpy(cs->valid.p, ones.data(), sizeof(int) * (size_t)B, cudaMemcpyHostToDevice));
            CU(cudaDeviceSynchronize());
        }
    }
    if (elbo_impl(h, B, hyper, false, ysub, ysub_shared, nullptr, nullptr, 0, cs, 2, max_iter, elbo_out, iters_out,
                  status_out, nullptr, false, stream)) return 1;
    if (cs) {
        if (mu_inout) CU(cudaMemcpy(mu_inout, cs->mu.p, sizeof(double) * (size_t)B * h->d, cudaMemcpyDeviceToHost));
        if (var_inout) CU(cudaMemcpy(var_inout, cs->var.p, sizeof(double) * (size_t)B * h->d, cudaMemcpyDeviceToHost));
    }
    return 0;
}

extern "C" int gprn_upload_ysub(gprn_handle* h, const double* ysub) {
    if (!h || !ysub) return fail("gprn_upload_ysub: null argument");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpy(h->d_ysub_shared, ysub, sizeof(double) * h->p * h->N, cudaMemcpyHostToDevice));
    CU(cudaDeviceSynchronize());
    return 0;
}

extern "C" int gprn_elbo_batched_dev(gprn_handle* h, int B, const double* d_hyper, int max_iter, double* d_elbo_out,
                                     int32_t* d_iters_out, int32_t* d_status_out, void* stream) {
    if (!d_hyper) return fail("gprn_elbo_batched_dev: d_hyper is null");
    return elbo_impl(h, B, d_hyper, true, nullptr, 1, nullptr, nullptr, 0, nullptr, 0, max_iter, d_elbo_out, d_iters_out,
                     d_status_out, nullptr, true, stream);
}

extern "C" int gprn_elbo_pool(gprn_handle* h, int64_t B, const double* hyper, int hyper_on_device, const double* ysub,
                              int ysub_shared, gprn_next_set_fn next, void* user, int max_slots, int state_mode,
                              int max_iter, double* elbo_out, int32_t* iters_out, int32_t* status_out,
                              int32_t* taken_out, int out_on_device, void* stream) {
    if (!h) return fail("null handle");
    if (!hyper) return fail("gprn_elbo_pool: hyper is null");
    if (state_mode < 0 || state_mode > 2) return fail("gprn_elbo_pool: state_mode must be 0, 1 or 2");
    return elbo_impl(h, B, hyper, hyper_on_device != 0, ysub, ysub ? ysub_shared : 1, next, user, max_slots,
                     state_mode ? &h->chain : nullptr, state_mode, max_iter, elbo_out, iters_out, status_out,
                     taken_out, out_on_device != 0, stream);
}

// ------------------------------------------------------------------------------------------------
// chain-state store
// ------------------------------------------------------------------------------------------------
extern "C" int gprn_chain_resize(gprn_handle* h, int64_t n) {
    if (!h || n < 1) return fail("gprn_chain_resize: bad argument");
    CU(cudaSetDevice(h->device));
    return chain_resize(h, h->chain, n);
}
extern "C" int gprn_chain_set(gprn_handle* h, int64_t first, int64_t n, const double* mu, const double* var) {
    if (!h || !mu || !var) return fail("gprn_chain_set: null argument");
    if (first < 0 || n < 1 || first + n > h->chain.n) return fail("gprn_chain_set: range outside the store");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpy((double*)h->chain.mu.p + (size_t)first * h->d, mu, sizeof(double) * (size_t)n * h->d, cudaMemcpyHostToDevice));
    CU(cudaMemcpy((double*)h->chain.var.p + (size_t)first * h->d, var, sizeof(double) * (size_t)n * h->d, cudaMemcpyHostToDevice));
    std::vector<int> ones(n, 1);
    CU(cudaMemcpy((int*)h->chain.valid.p + first, ones.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice));
    CU(cudaDeviceSynchronize());
    return 0;
}
extern "C" int gprn_chain_get(gprn_handle* h, int64_t first, int64_t n, double* mu, double* var, int32_t* valid) {
    if (!h) return fail("gprn_chain_get: null handle");
    if (first < 0 || n < 1 || first + n > h->chain.n) return fail("gprn_chain_get: range outside the store");
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    if (mu) CU(cudaMemcpy(mu, (double*)h->chain.mu.p + (size_t)first * h->d, sizeof(double) * (size_t)n * h->d, cudaMemcpyDeviceToHost));
    if (var) CU(cudaMemcpy(var, (double*)h->chain.var.p + (size_t)first * h->d, sizeof(double) * (size_t)n * h->d, cudaMemcpyDeviceToHost));
    if (valid) CU(cudaMemcpy(valid, (int*)h->chain.valid.p + first, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost));
    return 0;
}
extern "C" int gprn_chain_invalidate(gprn_handle* h, int64_t first, int64_t n) {
    if (!h) return fail("gprn_chain_invalidate: null handle");
    if (first < 0 || n < 1 || first + n > h->chain.n) return fail("gprn_chain_invalidate: range outside the store");
    CU(cudaSetDevice(h->device));
    CU(cudaMemset((int*)h->chain.valid.p + first, 0, sizeof(int) * (size_t)n));
    CU(cudaDeviceSynchronize());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// covariance matrix assembly (parity entry point for rows a1/a2)
// ------------------------------------------------------------------------------------------------
// Bump allocator over one grow-only device buffer: the scratch of gprn_kmatrix lives on the handle, so a kernel
// object's k(r) / K-matrix call costs no cudaMalloc / cudaFree once the buffer has reached its size.
struct Carve {
    char* base;
    size_t off = 0;
    explicit Carve(void* p) : base((char*)p) {}
    template <typename T>
    T* take(size_t n) {
        T* r = (T*)(base + off);
        off += ((n * sizeof(T) + 255) / 256) * 256;
        return r;
    }
};
static size_t carve_bytes(std::initializer_list<size_t> sizes) {
    size_t t = 0;
    for (size_t s : sizes) t += ((s + 255) / 256) * 256;
    return t;
}

extern "C" int gprn_kmatrix(gprn_handle* h, const int32_t* prog, int prog_len, const double* pars, int n_pars,
                            const double* t_rows, int n_rows, const double* t_cols, int n_cols, double nugget,
                            double* K_out, void* stream) {
    if (!h || !prog || !pars || !t_rows || !K_out) return fail("gprn_kmatrix: null argument");
    int need = check_prog(prog, prog_len);
    if (need < 0) return fail("gprn_kmatrix: malformed kernel program");
    if (need != n_pars) return fail("gprn_kmatrix: program needs " + std::to_string(need) + " parameters, got " + std::to_string(n_pars));
    const bool square = (t_cols == nullptr);
    if (square) { t_cols = t_rows; n_cols = n_rows; }
    if (n_rows < 1 || n_cols < 1) return fail("gprn_kmatrix: empty matrix");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->own_stream;
    const size_t kb = sizeof(double) * (size_t)n_rows * n_cols;
    if (ensure(h->kscratch, carve_bytes({sizeof(double) * (size_t)n_rows, sizeof(double) * (size_t)n_cols,
                                         sizeof(double) * (size_t)std::max(1, n_pars), sizeof(int32_t) * (size_t)prog_len, kb})))
        return 1;
    Carve cv(h->kscratch.p);
    double* d_tr = cv.take<double>(n_rows);
    double* d_tc = cv.take<double>(n_cols);
    double* d_par = cv.take<double>(std::max(1, n_pars));
    int32_t* d_tok = cv.take<int32_t>(prog_len);
    double* d_K = cv.take<double>((size_t)n_rows * n_cols);
    CU(cudaMemcpyAsync(d_tr, t_rows, sizeof(double) * n_rows, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_tc, t_cols, sizeof(double) * n_cols, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_par, pars, sizeof(double) * n_pars, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_tok, prog, sizeof(int32_t) * prog_len, cudaMemcpyHostToDevice, st));
    kassemble_rect_kernel<<<dim3((n_cols + NB - 1) / NB, (n_rows + NB - 1) / NB), 256, 0, st>>>(
        d_K, (size_t)n_cols, d_tr, n_rows, d_tc, n_cols, d_tok, prog_len, d_par, n_pars, square ? 1 : 0, nugget, 0, 0,
        nullptr, nullptr);
    LAUNCH_CHECK(h);
    CU(cudaMemcpyAsync(K_out, d_K, kb, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

// Per-device scratch of gprn_keval (which has no handle): grow-only, guarded by the process mutex.
static std::map<int, DevBuf> g_keval_scratch;

extern "C" int gprn_keval(int device, const int32_t* prog, int prog_len, const double* pars, int n_pars,
                          const double* r, int64_t n_rows, int64_t n_cols, int square, double* out) {
    if (!prog || !pars || !r || !out) return fail("gprn_keval: null argument");
    int need = check_prog(prog, prog_len);
    if (need < 0 || need != n_pars) return fail("gprn_keval: malformed kernel program or wrong parameter count");
    const long long n = (long long)n_rows * n_cols;
    if (n < 1) return fail("gprn_keval: empty array");
    if (device < 0) CU(cudaGetDevice(&device));          // the calling thread's current device
    if (check_device(device)) return 1;
    CU(cudaSetDevice(device));
    std::lock_guard<std::mutex> lk(g_mutex);
    DevBuf& buf = g_keval_scratch[device];
    if (ensure(buf, carve_bytes({sizeof(double) * (size_t)n, sizeof(double) * (size_t)n,
                                 sizeof(double) * (size_t)std::max(1, n_pars), sizeof(int32_t) * (size_t)prog_len})))
        return 1;
    Carve cv(buf.p);
    double* d_r = cv.take<double>(n);
    double* d_o = cv.take<double>(n);
    double* d_par = cv.take<double>(std::max(1, n_pars));
    int32_t* d_tok = cv.take<int32_t>(prog_len);
    CU(cudaMemcpy(d_r, r, sizeof(double) * n, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_par, pars, sizeof(double) * n_pars, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_tok, prog, sizeof(int32_t) * prog_len, cudaMemcpyHostToDevice));
    int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
    keval_kernel<<<blocks, 256>>>(d_o, d_r, n, n_cols, d_tok, prog_len, d_par, square);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(std::string("keval launch: ") + cudaGetErrorString(e));
    CU(cudaMemcpy(out, d_o, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// factorisation test hooks
// ------------------------------------------------------------------------------------------------
// Frees the listed device allocations when it goes out of scope (error paths included).
struct DevGuard {
    std::vector<void*> ptrs;
    ~DevGuard() { for (void* p : ptrs) if (p) cudaFree(p); }
    template <typename T>
    int alloc(T** out, size_t n) {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e != cudaSuccess) return fail(std::string("cudaMalloc: ") + cudaGetErrorString(e));
        ptrs.push_back(p);
        *out = (T*)p;
        return 0;
    }
};

static void pad_identity(std::vector<double>& pad, const double* A, int n, int Np) {
    pad.assign((size_t)Np * Np, 0.0);
    for (int i = 0; i < Np; i++) pad[(size_t)i * Np + i] = 1.0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) pad[(size_t)i * Np + j] = A[(size_t)i * n + j];
}

extern "C" int gprn_debug_factor(gprn_handle* h, int n, const double* A, double* L_out, double* X_out, double* logdet_out) {
    if (!h || !A || n < 1) return fail("gprn_debug_factor: bad argument");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = h->own_stream;
    const int Np = padded_size(n);
    std::vector<double> pad;
    pad_identity(pad, A, n, Np);
    DevGuard g;
    double *dW = nullptr, *dX = nullptr, *dld = nullptr;
    int *dids = nullptr, *dst = nullptr, *dctr = nullptr;
    if (g.alloc(&dW, (size_t)Np * Np) || g.alloc(&dX, (size_t)Np * Np) || g.alloc(&dld, 1) || g.alloc(&dids, 1) ||
        g.alloc(&dst, 1) || g.alloc(&dctr, 1)) return 1;
    // everything on `st`: the handle's stream is non-blocking, so legacy-stream memsets would race with it
    CU(cudaMemcpyAsync(dW, pad.data(), sizeof(double) * Np * Np, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(dX, 0, sizeof(double) * Np * Np, st));
    CU(cudaMemsetAsync(dld, 0, sizeof(double), st));
    CU(cudaMemsetAsync(dids, 0, sizeof(int), st));
    CU(cudaMemsetAsync(dst, 0, sizeof(int), st));
    CU(cudaMemsetAsync(dctr, 0, sizeof(int), st));
    if (use_two_level(Np) && TRTRI_MAXCH(Np) > 1 &&
        ensure(h->gpart, (size_t)(TRTRI_MAXCH(Np) - 1) * OUTER_KB * Np * sizeof(double))) return 1;
    if (factor_batch(h, Np, dW, dids, 1, dld, dst, dctr, dX, st)) return 1;
    CU(cudaStreamSynchronize(st));
    std::vector<double> out((size_t)Np * Np);
    if (L_out) {
        CU(cudaMemcpy(out.data(), dW, sizeof(double) * Np * Np, cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) L_out[(size_t)i * n + j] = j <= i ? out[(size_t)i * Np + j] : 0.0;
    }
    if (X_out) {
        CU(cudaMemcpy(out.data(), dX, sizeof(double) * Np * Np, cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) X_out[(size_t)i * n + j] = j <= i ? out[(size_t)i * Np + j] : 0.0;
    }
    if (logdet_out) CU(cudaMemcpy(logdet_out, dld, sizeof(double), cudaMemcpyDeviceToHost));
    int stt = 0;
    CU(cudaMemcpy(&stt, dst, sizeof(int), cudaMemcpyDeviceToHost));
    if (stt) return fail("gprn_debug_factor: matrix is not positive definite");
    return 0;
}

// Stress self-check of the ticketed panel step (stands in for compute-sanitizer racecheck, which this pool does not
// allow): `nmat` copies of A are factored `reps` times through the ONE-launch panel kernel (last-reader ticket, no
// spinning) and compared bit for bit with the two-launch path (potrf_col + trsm_col), which is documented
// bit-identical.  mismatches_out: number of (repetition, matrix) pairs whose factor or log-det differed.
extern "C" int gprn_debug_panel_stress(gprn_handle* h, int n, const double* A, int nmat, int reps, int64_t* mismatches_out) {
    if (!h || !A || n < 1 || nmat < 1 || reps < 1 || !mismatches_out) return fail("gprn_debug_panel_stress: bad argument");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = h->own_stream;
    const int Np = padded_size(n);
    const size_t mat = (size_t)Np * Np;
    std::vector<double> pad;
    pad_identity(pad, A, n, Np);
    DevGuard g;
    double *dA = nullptr, *dW = nullptr, *dRef = nullptr, *dld = nullptr, *dldref = nullptr;
    int *dids = nullptr, *dst = nullptr, *dctr = nullptr;
    unsigned long long* dmis = nullptr;
    if (g.alloc(&dA, mat) || g.alloc(&dW, mat * nmat) || g.alloc(&dRef, mat) || g.alloc(&dld, (size_t)nmat) ||
        g.alloc(&dldref, 1) || g.alloc(&dids, (size_t)nmat) || g.alloc(&dst, (size_t)nmat) || g.alloc(&dctr, (size_t)nmat) ||
        g.alloc(&dmis, 1)) return 1;
    std::vector<int> ids(nmat);
    for (int i = 0; i < nmat; i++) ids[i] = i;
    CU(cudaMemcpyAsync(dA, pad.data(), sizeof(double) * mat, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(dids, ids.data(), sizeof(int) * nmat, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(dst, 0, sizeof(int) * nmat, st));
    CU(cudaMemsetAsync(dctr, 0, sizeof(int) * nmat, st));
    CU(cudaMemsetAsync(dmis, 0, sizeof(unsigned long long), st));
    // reference: two-launch path, one matrix
    CU(cudaMemcpyAsync(dRef, dA, sizeof(double) * mat, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemsetAsync(dldref, 0, sizeof(double), st));
    if (factor_batch(h, Np, dRef, dids, 1, dldref, dst, dctr, nullptr, st, 0, 0)) return 1;
    for (int rep = 0; rep < reps; rep++) {
        for (int i = 0; i < nmat; i++)
            CU(cudaMemcpyAsync(dW + (size_t)i * mat, dA, sizeof(double) * mat, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemsetAsync(dld, 0, sizeof(double) * nmat, st));
        if (factor_batch(h, Np, dW, dids, nmat, dld, dst, dctr, nullptr, st, 0, 1)) return 1;
        lower_mismatch_kernel<<<nmat, 256, 0, st>>>(dW, dRef, dld, dldref, Np, dmis);
        LAUNCH_CHECK(h);
    }
    unsigned long long mis = 0;
    CU(cudaMemcpyAsync(&mis, dmis, sizeof(mis), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *mismatches_out = (int64_t)mis;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// prediction
// ------------------------------------------------------------------------------------------------
// B hyper-parameter sets with their variational states.  The sets are processed in blocks that fit the workspace:
// assembly, factorisation and the alpha = A^-1 m solves of a block are batched over ALL its GPs (one launch list),
// then each set's T test epochs go through Kstar assembly / mean / variance-norm kernels batched over its M GPs.
static int predict_impl(gprn_handle* h, int B, const double* hyper, const double* mu, const double* var,
                        const double* tstar, int T, const double* mean_at_tstar, int mean_shared, double* pred_mean,
                        double* pred_var, double* node_pred, double* weight_pred, void* stream) {
    if (!h || !hyper || !mu || !var || !tstar || !pred_mean || !pred_var) return fail("gprn_predict: null argument");
    if (!h->model_set) return fail("gprn_predict: call gprn_set_model first");
    if (T < 1 || B < 1) return fail("gprn_predict: T and B must be >= 1");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->own_stream;
    const int N = h->N, Np = h->Np, nt = h->nt, M = h->M, q = h->q, p = h->p;
    const int ntri = nt * (nt + 1) / 2;
    int nslot = chunk_size(h, B, true);
    if (nslot < 1) return fail("gprn_predict: not enough device memory for one model of this size");
    Engine E;
    memset(&E, 0, sizeof(E));
    if (setup_engine(h, nslot, E, true)) return 1;
    ElboCtx& c = E.c;
    if (ensure(h->hyper, (size_t)B * h->H * sizeof(double))) return 1;
    CU(cudaMemcpyAsync(h->hyper.p, hyper, sizeof(double) * (size_t)B * h->H, cudaMemcpyHostToDevice, st));
    c.hyper = (const double*)h->hyper.p;
    // test points in chunks of Tc epochs; Kstar chunk of all M GPs: [M][Tc][Np] (<= ~2 GB)
    const bool big = (Np % G_BN == 0) && Np >= 256;          // DMMA GEMM core for the variance norms
    const int tunit = big ? G_BM : NB;
    int TC = 4096;
    if (const char* e = getenv("GPRN_PREDICT_TC")) TC = std::max(tunit, (atoi(e) / tunit) * tunit);    // tests: force chunking
    while (TC > tunit && (size_t)M * TC * Np * sizeof(double) > ((size_t)2 << 30)) TC /= 2;
    const int Tc = std::min(TC, ((T + tunit - 1) / tunit) * tunit);
    const int nparts = big ? Np / G_BN : 1;
    const size_t kstride = (size_t)Tc * Np;
    if (ensure(h->ks, sizeof(double) * (size_t)M * kstride)) return 1;
    // columns >= N of every Kstar row are never written and must read as zero
    CU(cudaMemsetAsync(h->ks.p, 0, sizeof(double) * (size_t)M * kstride, st));
    // pred: tstar[T], gp_mean[M][T], gp_var[M][T], rownorm[M][nparts][Tc], mean_t[p][T], out mean/var [T*p] x 2
    const size_t rstride = (size_t)nparts * Tc;
    const size_t pd = (size_t)T + 2 * (size_t)M * T + (size_t)M * rstride + (size_t)p * T + 2 * (size_t)T * p;
    if (ensure(h->pred, sizeof(double) * pd)) return 1;
    double* d_ts = (double*)h->pred.p;
    double* d_gm = d_ts + T;
    double* d_gv = d_gm + (size_t)M * T;
    double* d_rn = d_gv + (size_t)M * T;
    double* d_mt = d_rn + (size_t)M * rstride;
    double* d_pm = d_mt + (size_t)p * T;
    double* d_pv = d_pm + (size_t)T * p;
    double* d_ks = (double*)h->ks.p;
    CU(cudaMemcpyAsync(d_ts, tstar, sizeof(double) * T, cudaMemcpyHostToDevice, st));
    if (!mean_at_tstar) CU(cudaMemsetAsync(d_mt, 0, sizeof(double) * (size_t)p * T, st));
    else if (mean_shared) CU(cudaMemcpyAsync(d_mt, mean_at_tstar, sizeof(double) * (size_t)p * T, cudaMemcpyHostToDevice, st));
    const int square = (T == N) ? 1 : 0;        // WhiteNoise quirk Q9 applies to Kstar by shape
    ProgTable pt{h->d_tok, h->d_len, h->d_par_off};
    std::vector<double> mv, vv;
    std::vector<int> act, none, slot_set(nslot, 0), mst;
    for (int b0 = 0; b0 < B; b0 += nslot) {
        const int nb = std::min(nslot, B - b0);
        // variational means -> vv, variances -> Dv (zero padded, one vector per GP in matrix order)
        mv.assign((size_t)nb * M * Np, 0.0);
        vv.assign((size_t)nb * M * Np, 0.0);
        act.resize(nb);
        for (int s = 0; s < nb; s++) {
            act[s] = s;
            slot_set[s] = b0 + s;
            const double* mub = mu + (size_t)(b0 + s) * h->d;
            const double* varb = var + (size_t)(b0 + s) * h->d;
            for (int m = 0; m < M; m++) {
                size_t so;
                if (m < q) so = (size_t)m * N;
                else { int ji = m - q, j = ji / p, i = ji % p; so = (size_t)q * N + (size_t)(i * q + j) * N; }   // muW[p,q,N]
                double* dm = mv.data() + ((size_t)s * M + m) * Np;
                double* dv = vv.data() + ((size_t)s * M + m) * Np;
                for (int n = 0; n < N; n++) { dm[n] = mub[so + n]; dv[n] = varb[so + n]; }
            }
        }
        CU(cudaMemcpyAsync(c.vv, mv.data(), sizeof(double) * mv.size(), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(c.Dv, vv.data(), sizeof(double) * vv.size(), cudaMemcpyHostToDevice, st));
        if (upload_lists(h, E, act, none, slot_set, st)) return 1;
        kassemble_sym_kernel<<<dim3(ntri, M, nb), 256, 0, st>>>(E.K, h->d_time, c.hyper, h->H, pt, M, N, Np, 1.25e-12,
                                                                 nullptr, E.d_slot_set);                    // _gp.py:49
        LAUNCH_CHECK(h);
        CU(cudaMemsetAsync(c.logdetA, 0, sizeof(double) * (size_t)nb * M, st));
        CU(cudaMemsetAsync(c.mstatus, 0, sizeof(int) * (size_t)nb * M, st));
        CU(cudaMemsetAsync(E.d_ctr, 0, sizeof(int) * (size_t)nb * M, st));
        form_a_kernel<<<dim3(ntri, nb * M), 256, 0, st>>>(E.W, E.K, c.Dv, E.d_ida, Np);       // + diag(v), _gp.py:125
        LAUNCH_CHECK(h);
        if (factor_batch_multi(h, E.W, E.d_ida, nb * M, c.logdetA, c.mstatus, E.d_ctr, E.X, st)) return 1;
        if (solve_batch(h, E.X, E.d_ida, nb * M, c.vv, c.zv, c.uv, c.gv, st)) return 1;   // uv = alpha
        for (int s = 0; s < nb; s++) {
            const int b = b0 + s;
            const double* hy = c.hyper + (size_t)b * h->H;
            const double* Xs = E.X + (size_t)s * M * Np * Np;
            const double* alpha = c.uv + (size_t)s * M * Np;
            if (mean_at_tstar && !mean_shared)
                CU(cudaMemcpyAsync(d_mt, mean_at_tstar + (size_t)b * p * T, sizeof(double) * (size_t)p * T, cudaMemcpyHostToDevice, st));
            for (int t0 = 0; t0 < T; t0 += Tc) {
                const int tn = std::min(Tc, T - t0);
                const int tpad = ((tn + tunit - 1) / tunit) * tunit;
                // rows tn..tpad-1 may hold stale finite values of an earlier chunk: every output row depends on its
                // own Kstar row only, and rows >= tn are never read back
                kassemble_rect_kernel<<<dim3((N + NB - 1) / NB, (tn + NB - 1) / NB, M), 256, 0, st>>>(
                    d_ks, (size_t)Np, d_ts + t0, tn, h->d_time, N, h->d_tok, 0, hy, h->H, square, 0.0, t0, kstride,
                    h->d_len, h->d_par_off);
                LAUNCH_CHECK(h);
                rect_gemv_kernel<<<dim3((tn + 7) / 8, M), 256, 0, st>>>(d_gm + t0, (size_t)T, d_ks, kstride, (size_t)Np, alpha, (size_t)Np, tn, N);
                LAUNCH_CHECK(h);
                if (big)
                    predict_norm128_kernel<<<dim3(tpad / G_BM, Np / G_BN, M), G_THREADS, GEMM128_SMEM, st>>>(d_rn, rstride, Tc, d_ks, kstride, Xs, Np);
                else
                    predict_norm_kernel<<<dim3(tpad / NB, M), 128, 2 * TILE_SMEM, st>>>(d_rn, rstride, d_ks, kstride, Xs, Np, N);
                LAUNCH_CHECK(h);
                predict_var_kernel<<<dim3((tn + 255) / 256, M), 256, 0, st>>>(d_gv + t0, (size_t)T, d_rn, rstride, nparts, Tc, tn, h->d_tok, h->d_len, hy, h->d_par_off, 1.25e-12);
                LAUNCH_CHECK(h);
            }
            predict_combine_kernel<<<(T * p + 255) / 256, 256, 0, st>>>(d_pm, d_pv, d_gm, d_gv, d_mt, hy + h->H - p, T, p, q);
            LAUNCH_CHECK(h);
            CU(cudaMemcpyAsync(pred_mean + (size_t)b * T * p, d_pm, sizeof(double) * (size_t)T * p, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(pred_var + (size_t)b * T * p, d_pv, sizeof(double) * (size_t)T * p, cudaMemcpyDeviceToHost, st));
            if (node_pred) CU(cudaMemcpyAsync(node_pred + (size_t)b * q * T, d_gm, sizeof(double) * (size_t)q * T, cudaMemcpyDeviceToHost, st));
            if (weight_pred) CU(cudaMemcpyAsync(weight_pred + (size_t)b * q * p * T, d_gm + (size_t)q * T, sizeof(double) * (size_t)q * p * T, cudaMemcpyDeviceToHost, st));
        }
        mst.resize((size_t)nb * M);
        CU(cudaMemcpyAsync(mst.data(), c.mstatus, sizeof(int) * mst.size(), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        for (size_t i = 0; i < mst.size(); i++)
            if (mst[i])
                return fail("gprn_predict: K + diag(var) is not positive definite for component " + std::to_string(i % M) +
                            " of set " + std::to_string(b0 + (int)(i / M)));
    }
    return 0;
}

extern "C" int gprn_predict(gprn_handle* h, const double* hyper, const double* mu, const double* var,
                            const double* tstar, int T, const double* mean_at_tstar, double* pred_mean,
                            double* pred_var, double* node_pred, double* weight_pred, void* stream) {
    return predict_impl(h, 1, hyper, mu, var, tstar, T, mean_at_tstar, 1, pred_mean, pred_var, node_pred, weight_pred, stream);
}

extern "C" int gprn_predict_batched(gprn_handle* h, int B, const double* hyper, const double* mu, const double* var,
                                    const double* tstar, int T, const double* mean_at_tstar, int mean_shared,
                                    double* pred_mean, double* pred_var, double* node_pred, double* weight_pred,
                                    void* stream) {
    return predict_impl(h, B, hyper, mu, var, tstar, T, mean_at_tstar, mean_shared, pred_mean, pred_var, node_pred,
                        weight_pred, stream);
}

// ------------------------------------------------------------------------------------------------
// prior draws
// ------------------------------------------------------------------------------------------------
extern "C" int gprn_sample(gprn_handle* h, const double* hyper, const double* z, double nugget, double* out,
                           void* stream) {
    if (!h || !hyper || !z || !out) return fail("gprn_sample: null argument");
    if (!h->model_set) return fail("gprn_sample: call gprn_set_model first");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->own_stream;
    const int N = h->N, Np = h->Np, nt = h->nt, M = h->M;
    const int ntri = nt * (nt + 1) / 2;
    Engine E;
    memset(&E, 0, sizeof(E));
    if (setup_engine(h, 1, E, true)) return 1;
    ElboCtx& c = E.c;
    if (ensure(h->hyper, (size_t)h->H * sizeof(double))) return 1;
    CU(cudaMemcpyAsync(h->hyper.p, hyper, sizeof(double) * h->H, cudaMemcpyHostToDevice, st));
    c.hyper = (const double*)h->hyper.p;
    std::vector<double> zp((size_t)M * Np, 0.0);
    for (int m = 0; m < M; m++)
        for (int n = 0; n < N; n++) zp[(size_t)m * Np + n] = z[(size_t)m * N + n];
    CU(cudaMemcpyAsync(c.vv, zp.data(), sizeof(double) * M * Np, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(c.Dv, 0, sizeof(double) * M * Np, st));
    std::vector<int> act(1, 0), none, slot_set(1, 0);
    if (upload_lists(h, E, act, none, slot_set, st)) return 1;
    ProgTable pt{h->d_tok, h->d_len, h->d_par_off};
    kassemble_sym_kernel<<<dim3(ntri, M, 1), 256, 0, st>>>(E.K, h->d_time, c.hyper, h->H, pt, M, N, Np, nugget, nullptr, E.d_slot_set);
    LAUNCH_CHECK(h);
    CU(cudaMemsetAsync(c.logdetA, 0, sizeof(double) * M, st));
    CU(cudaMemsetAsync(c.mstatus, 0, sizeof(int) * M, st));
    CU(cudaMemsetAsync(E.d_ctr, 0, sizeof(int) * M, st));
    form_a_kernel<<<dim3(ntri, M), 256, 0, st>>>(E.W, E.K, c.Dv, E.d_ida, Np);
    LAUNCH_CHECK(h);
    if (factor_batch_multi(h, E.W, E.d_ida, M, c.logdetA, c.mstatus, E.d_ctr, nullptr, st)) return 1;
    trmv_lower_kernel<<<dim3(Np / 8, M), 256, 0, st>>>(c.zv, E.W, c.vv, E.d_ida, nullptr, Np);   // L z
    LAUNCH_CHECK(h);
    for (int m = 0; m < M; m++)
        CU(cudaMemcpyAsync(out + (size_t)m * N, c.zv + (size_t)m * Np, sizeof(double) * N, cudaMemcpyDeviceToHost, st));
    std::vector<int> mst(M);
    CU(cudaMemcpyAsync(mst.data(), c.mstatus, sizeof(int) * M, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int m = 0; m < M; m++)
        if (mst[m])
            return fail("gprn_sample: K + nugget*I is not positive definite for component " + std::to_string(m) +
                        " (raise the nugget or add a WhiteNoise term)");
    return 0;
}

#ifdef GPRN_TRACE
// Development-only entry points of the -DGPRN_TRACE build (tools/trace_run.py); not part of the C ABI.
static TraceRec* g_trace_dev = nullptr;
static unsigned g_trace_cap_host = 0;
extern "C" int gprn_trace_begin(unsigned cap) {
    if (!g_trace_dev || cap > g_trace_cap_host) {
        if (g_trace_dev) cudaFree(g_trace_dev);
        CU(cudaMalloc(&g_trace_dev, sizeof(TraceRec) * (size_t)cap));
        g_trace_cap_host = cap;
    }
    unsigned zero = 0;
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpyToSymbol(g_trace_buf, &g_trace_dev, sizeof(g_trace_dev)));
    CU(cudaMemcpyToSymbol(g_trace_cap, &g_trace_cap_host, sizeof(unsigned)));
    CU(cudaMemcpyToSymbol(g_trace_n, &zero, sizeof(unsigned)));
    CU(cudaDeviceSynchronize());
    return 0;
}
extern "C" int gprn_trace_mid_phases(unsigned long long* out128) {    // [set-up | iteration][row][phase]; resets
#ifdef GPRN_TRACE
    unsigned long long tmp[2 * MID_MAX_NT * 8];
    CU(cudaMemcpyFromSymbol(tmp, g_mid_phase, sizeof(tmp)));
    memcpy(out128, tmp, sizeof(tmp));
    memset(tmp, 0, sizeof(tmp));
    CU(cudaMemcpyToSymbol(g_mid_phase, tmp, sizeof(tmp)));
    return 0;
#else
    (void)out128;
    return fail("library built without -DGPRN_TRACE");
#endif
}
extern "C" int gprn_trace_small_phases(unsigned long long* out12, int reset) {
    CU(cudaDeviceSynchronize());
    unsigned long long tmp[16];
    CU(cudaMemcpyFromSymbol(tmp, g_small_phase, sizeof(tmp)));
    for (int i = 0; i < 12; i++) out12[i] = tmp[i];
    if (reset) {
        memset(tmp, 0, sizeof(tmp));
        CU(cudaMemcpyToSymbol(g_small_phase, tmp, sizeof(tmp)));
    }
    return 0;
}
extern "C" int gprn_trace_dump(const char* path) {
    CU(cudaDeviceSynchronize());
    unsigned n = 0;
    CU(cudaMemcpyFromSymbol(&n, g_trace_n, sizeof(unsigned)));
    if (n > g_trace_cap_host) n = g_trace_cap_host;
    std::vector<TraceRec> recs(n);
    if (n) CU(cudaMemcpy(recs.data(), g_trace_dev, sizeof(TraceRec) * (size_t)n, cudaMemcpyDeviceToHost));
    FILE* f = fopen(path, "wb");
    if (!f) return fail("gprn_trace_dump: cannot open file");
    fwrite(recs.data(), sizeof(TraceRec), n, f);
    fclose(f);
    TraceRec* none = nullptr;
    CU(cudaMemcpyToSymbol(g_trace_buf, &none, sizeof(none)));
    return (int)n;
}
#endif

extern "C" int64_t gprn_launch_count(gprn_handle* h) { return h ? h->launches : 0; }
extern "C" int gprn_reset_launch_count(gprn_handle* h) {
    if (h) { h->launches = 0; h->graph_launches = 0; }
    return 0;
}
extern "C" double gprn_last_elbo_ms(gprn_handle* h) { return h ? h->last_ms : 0.0; }
extern "C" int64_t gprn_last_total_iters(gprn_handle* h) { return h ? h->last_total_iters : 0; }
extern "C" int64_t gprn_graph_launch_count(gprn_handle* h) { return h ? h->graph_launches : 0; }
extern "C" int64_t gprn_last_rounds(gprn_handle* h) { return h ? h->last_lockstep_rounds : 0; }
