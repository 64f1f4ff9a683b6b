/*
 * gprn_b200.h -- C ABI of the B200-native mean-field GPRN hot path.
 *
 * The reference (iastro-pt/gpyrn) is pure Python and has no FFI of its own; the boundary this
 * library sits under is the method surface of gpyrn.meanfield.inference.  Each entry
 * point below names the reference code it replaces (paths relative to the reference checkout).
 * INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a non-zero
 * code on failure (message via gprn_last_error(), thread-local); no exceptions cross the boundary.
 * All arithmetic is FP64.  Host buffers are owned by the caller; device memory is owned by the
 * handle.  One handle per (process, device); calls on one handle must be serialised by the caller.
 * `stream` is a cudaStream_t passed as void* (NULL = the handle's own stream); calls return after
 * their results are in the caller's buffers.
 */
#ifndef GPRN_B200_H
#define GPRN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gprn_handle gprn_handle;

/* kernel-program opcodes (postfix).  Leaf opcodes consume their parameters, in the reference's
 * `.pars` order, from the hyper-parameter vector; GPRN_OP_ADD / GPRN_OP_MUL pop two values.
 * Replaces covFunction.__call__ of gpyrn/covfunc.py:169-170 (SE), :211-213 (Periodic), :251-255
 * (QuasiPeriodic), :286-288 (RationalQuadratic), :370-373 (Matern32), :391-396 (Matern52),
 * :144-148 (WhiteNoise), :67-68 (Sum), :76-77 (Multiplication). */
enum {
    GPRN_OP_SE = 1,  /* theta, ell            */
    GPRN_OP_PER = 2, /* theta, P, ell         */
    GPRN_OP_QP = 3,  /* theta, elle, P, ellp  */
    GPRN_OP_RQ = 4,  /* theta, alpha, ell     */
    GPRN_OP_M32 = 5, /* theta, ell            */
    GPRN_OP_M52 = 6, /* theta, ell            */
    GPRN_OP_WN = 7,  /* w                     */
    /* SURVEY 8(f).3 "next" kernels: covfunc.py:107-125 (Constant), :291-310 (RQP), :313-328 (Cosine), :331-352 (Exponential) */
    GPRN_OP_CONST = 8, /* c                              */
    GPRN_OP_RQP = 9,   /* theta, alpha, elle, P, ellp    */
    GPRN_OP_COS = 10,  /* theta, P                       */
    GPRN_OP_EXP = 11,  /* theta, ell                     */
    /* Derivative(k) = d2k/dxi dxj of the twice-differentiable kernels (covfunc.py:80-104 Derivative;
     * :182-185 SE._dkdxidj, :215-221 Periodic._dkdxidj, :257-266 QuasiPeriodic._dkdxidj); same parameters as k */
    GPRN_OP_DSE = 12,  /* theta, ell                     */
    GPRN_OP_DPER = 13, /* theta, P, ell                  */
    GPRN_OP_DQP = 14,  /* theta, elle, P, ellp           */
    /* the stationary "other" kernels of the reference that work there: covfunc.py:415-432 (GammaExp), :458-474
     * (Piecewise), :477-496 (Paciorek), :499-519 (NewPeriodic), :522-546 (QuasiNewPeriodic), :645-665 (CosPeriodic),
     * :668-688 (QuasiCosPeriodic).  Linear, Polynomial and the Harmonic kernels take (t1, t2) and cannot be called
     * by the reference's own inference; NewRQP raises there (np.sine). */
    GPRN_OP_GEXP = 15,   /* theta, gamma, ell              */
    GPRN_OP_PIECE = 16,  /* eta                            */
    GPRN_OP_PAC = 17,    /* amplitude, ell_1, ell_2        */
    GPRN_OP_NPER = 18,   /* amplitude, alpha2, P, ell      */
    GPRN_OP_QNPER = 19,  /* amplitude, alpha2, ell_e, P, ell_p */
    GPRN_OP_COSP = 20,   /* amplitude, P, ell              */
    GPRN_OP_QCOSP = 21,  /* amplitude, ell_e, P, ell_p     */
    GPRN_OP_ADD = 100,
    GPRN_OP_MUL = 101
};
#define GPRN_MAX_PROG 16 /* tokens per kernel program */

/* per-evaluation status words */
enum {
    GPRN_STATUS_OK = 0,
    GPRN_STATUS_NOT_PD = 1,  /* a Cholesky pivot was <= 0 or NaN: ELBO is NaN (what JAX's cholesky yields) */
    GPRN_STATUS_MAX_ITER = 2 /* meanfield.py:648 'Max iterations reached' */
};

const char* gprn_last_error(void);

/* Library/ABI version and the compute capability it was built for (100 = sm_100a). */
int gprn_version(void);
int gprn_built_for_sm(void);

/* Data holder: replaces inference.__init__ (gpyrn/meanfield.py:106-134).
 * time[N]; y[p*N] raw observations row-major (p,N); yerr[p*N]. */
int gprn_create(int device, int N, int p, int q, const double* time, const double* y, const double* yerr,
                gprn_handle** out);
int gprn_destroy(gprn_handle* h);

/* Optional cap (bytes) on the device workspace used by gprn_elbo_batched (0 = automatic). */
int gprn_set_workspace_limit(gprn_handle* h, uint64_t bytes);
/* Optional cap on the number of evaluations kept in flight by the batched entry points (0 = as many as fit). */
int gprn_set_max_slots(gprn_handle* h, int slots);

/* Model structure: replaces the kernel objects held by inference.set_components
 * (gpyrn/meanfield.py:136-178).  Programs are concatenated; node_prog_off has q+1 entries,
 * weight_prog_off q*p+1 (weight index j*p+i = node j -> output i, meanfield.py:749-750).
 * A hyper-parameter set is laid out as [node0 pars .. node{q-1} pars, weight0 pars .., jitter_0..p-1]
 * (the reference's get_parameters order without the mean parameters, meanfield.py:193-202);
 * n_hyper is its length. */
int gprn_set_model(gprn_handle* h, const int32_t* node_prog, const int32_t* node_prog_off,
                   const int32_t* weight_prog, const int32_t* weight_prog_off, int n_hyper);

/* B independent ELBO evaluations: replaces inference.ELBOcalc + ELBOaux + _updateSigMu +
 * _expectedLogLike + _expectedLogPrior + _entropy + _initMuVar + _KMatrix + _cholNugget
 * (gpyrn/meanfield.py:561-649, 651-710, 713-893, 895-990, 992-1067, 1069-1093, 491-510, 413-434, 71-88).
 *   hyper      [B*n_hyper]   host
 *   ysub       y - mean, host: [B*p*N] or, with ysub_shared != 0, [p*N] used by every set
 *   init_mode  0: mu/var from _initMuVar on the device (mu_inout/var_inout are outputs if non-NULL)
 *              1: mu_inout/var_inout [B*d] supply the initial state ('previous') and receive the result
 *   max_iter   <0 -> 10000 (meanfield.py:615-616)
 *   elbo_out[B], iters_out[B], status_out[B]; trace_out (may be NULL) [B*trace_len] last ELBOs
 * d = N*q*(p+1); state layout as the reference's flat u (meanfield.py:487-488). */
int gprn_elbo_batched(gprn_handle* h, int B, const double* hyper, const double* ysub, int ysub_shared,
                      int init_mode, double* mu_inout, double* var_inout, int max_iter, double* elbo_out,
                      int32_t* iters_out, int32_t* status_out, void* stream);

/* Same, with hyper / elbo_out / iters_out / status_out already resident in device memory
 * (ysub must have been uploaded with gprn_upload_ysub).  Asynchronous on `stream` except for the
 * per-iteration convergence poll; used by bench.py for the HBM-resident `value`. */
int gprn_upload_ysub(gprn_handle* h, const double* ysub /* p*N, shared by all sets */);
int gprn_elbo_batched_dev(gprn_handle* h, int B, const double* d_hyper, int max_iter, double* d_elbo_out,
                          int32_t* d_iters_out, int32_t* d_status_out, void* stream);

/* Pool form of the batched evaluation: the data-parallel entry for sweeps, multi-start optimisation and MCMC
 * walkers (the reference's pattern is a process pool over independent walkers, gpyrn/examples/example_4.py:66-68,
 * each calling inference.nELBO / logposterior, gpyrn/meanfield.py:1095-1111, 1214-1219).
 *   hyper      [B*n_hyper]  the whole pool, host or (hyper_on_device != 0) device memory
 *   ysub       as gprn_elbo_batched (host), or NULL: keep what gprn_upload_ysub / the last call left on the device
 *   next,user  work source: called from the calling thread, returns the next pool index to evaluate or -1 when
 *              the pool is exhausted.  NULL: every set 0..B-1 in order.  Several processes (one per GPU) that
 *              share one counter split a pool dynamically: each takes a new set whenever one of its slots frees up.
 *   max_slots  evaluations kept in flight on this GPU (0: the handle's default, see gprn_set_max_slots).  A set
 *              that converges is retired and its slot refilled from the work source in the next lock-step round
 *              (continuous batching).
 *   state_mode 0: every evaluation starts from _initMuVar ('init').
 *              1: chain-state store ('previous', meanfield.py:598-607): set b starts from chain b when that entry is
 *                 valid and from 'init' otherwise; the final state is written back when the evaluation converged
 *                 (the reference caches self._mu / self._var only then, :643-646).
 *              2: as 1, but the final state is always written back.
 *   outputs    [B] each, host or (out_on_device != 0) device; entries of sets this call did not evaluate are zero
 *              and taken_out (may be NULL) is 1 exactly for the sets it did evaluate. */
typedef int64_t (*gprn_next_set_fn)(void* user);
int gprn_elbo_pool(gprn_handle* h, int64_t B, const double* hyper, int hyper_on_device, const double* ysub,
                   int ysub_shared, gprn_next_set_fn next, void* user, int max_slots, int state_mode, int max_iter,
                   double* elbo_out, int32_t* iters_out, int32_t* status_out, int32_t* taken_out, int out_on_device,
                   void* stream);

/* Device-resident variational state of n chains, d doubles each for mu and var: the batched form of the reference's
 * self._mu / self._var cache (gpyrn/meanfield.py:112-113, 598-607, 643-646).  resize invalidates every entry;
 * set uploads host state and marks it valid; get downloads (any of mu / var / valid may be NULL). */
int gprn_chain_resize(gprn_handle* h, int64_t n);
int gprn_chain_set(gprn_handle* h, int64_t first, int64_t n, const double* mu, const double* var);
int gprn_chain_get(gprn_handle* h, int64_t first, int64_t n, double* mu, double* var, int32_t* valid);
int gprn_chain_invalidate(gprn_handle* h, int64_t first, int64_t n);

/* Covariance-matrix assembly: replaces inference._KMatrix (gpyrn/meanfield.py:413-434),
 * _gp.GP._kernel_matrix / _predict_kernel_matrix (gpyrn/_gp.py:40-62).
 * K_out[n_rows*n_cols] host, row-major.  t_cols == NULL means the square case k(t_rows - t_rows^T)
 * with `nugget` added on the diagonal (WhiteNoise follows quirk covfunc.py:144-148: identity-by-
 * position for square output, constant otherwise). */
int gprn_kmatrix(gprn_handle* h, const int32_t* prog, int prog_len, const double* pars, int n_pars,
                 const double* t_rows, int n_rows, const double* t_cols, int n_cols, double nugget,
                 double* K_out, void* stream);

/* Element-wise kernel evaluation out[i][j] = k(r[i][j]) on an arbitrary lag array (host in, host out):
 * replaces covFunction.__call__(r) (gpyrn/covfunc.py, same lines as the opcode table above) for the
 * host-side kernel objects.  `square` != 0 applies the WhiteNoise identity-by-position rule. */
/* device < 0: the calling thread's current CUDA device. */
int gprn_keval(int device, const int32_t* prog, int prog_len, const double* pars, int n_pars, const double* r,
               int64_t n_rows, int64_t n_cols, int square, double* out);

/* GPRN prediction: replaces inference._Prediction (gpyrn/meanfield.py:1289-1379) and
 * _gp.GP.prediction (gpyrn/_gp.py:107-138).
 *   hyper[n_hyper], mu[d], var[d] host; tstar[T]; mean_at_tstar[p*T] (host-evaluated mean functions)
 *   pred_mean[T*p], pred_var[T*p] row-major (T,p); node_pred[q*T], weight_pred[q*p*T] may be NULL. */
/* Prior draws of all M component GPs on the training epochs: out[m][:] = chol(K_m + nugget*I) z[m][:], with z
 * standard-normal variates supplied by the caller (host, M*N; component order j nodes, then q + j*p + i weights).
 * Replaces inference.sample / _sample_from_gp (gpyrn/meanfield.py:517-539) on K = _tinyNuggetKMatrix
 * (:436-453, nugget 1.25e-12).  The reference draws through scipy's eigen-decomposition (allow_singular=True);
 * the Cholesky factor gives the same distribution whenever K + nugget*I is numerically positive definite and
 * an error (no CPU fallback) when it is not. */
int gprn_sample(gprn_handle* h, const double* hyper, const double* z, double nugget, double* out, void* stream);

int gprn_predict(gprn_handle* h, const double* hyper, const double* mu, const double* var,
                 const double* tstar, int T, const double* mean_at_tstar, double* pred_mean,
                 double* pred_var, double* node_pred, double* weight_pred, void* stream);

/* The same for B hyper-parameter sets (a posterior chain): hyper[B*n_hyper], mu / var [B*d]; mean_at_tstar is
 * [p*T] shared by all sets (mean_shared != 0) or [B*p*T]; outputs [B][T][p], node_pred [B][q][T],
 * weight_pred [B][q*p][T].  Assembly, factorisation and the A^-1 m solves are batched over all GPs of as many
 * sets as fit the workspace. */
int gprn_predict_batched(gprn_handle* h, int B, const double* hyper, const double* mu, const double* var,
                         const double* tstar, int T, const double* mean_at_tstar, int mean_shared,
                         double* pred_mean, double* pred_var, double* node_pred, double* weight_pred, void* stream);

/* Test hook for the factorisation kernels: A[n*n] host SPD (row-major) -> L = chol(A) (lower),
 * X = L^-1 (lower), logdet(A).  Either output may be NULL. */
int gprn_debug_factor(gprn_handle* h, int n, const double* A, double* L_out, double* X_out, double* logdet_out);

/* Stress self-check of the one-launch panel step (last-reader ticket): nmat copies of A are factored reps times
 * and compared bit for bit with the two-launch path; *mismatches_out counts differing (repetition, matrix) pairs. */
int gprn_debug_panel_stress(gprn_handle* h, int n, const double* A, int nmat, int reps, int64_t* mismatches_out);

/* Counters: kernels launched by this handle since creation / last reset, and device time of the
 * last gprn_elbo_batched* call in milliseconds (CUDA events on the launching stream). */
int64_t gprn_launch_count(gprn_handle* h);
int gprn_reset_launch_count(gprn_handle* h);
double gprn_last_elbo_ms(gprn_handle* h);
/* Sum over the evaluations of the last batched call of the iteration counts (for flop accounting). */
int64_t gprn_last_total_iters(gprn_handle* h);
/* CUDA-graph replays (one per lock-step iteration when no slot is being refilled) since the last reset; the
 * kernels inside them are included in gprn_launch_count.  Lock-step rounds of the last batched call. */
int64_t gprn_graph_launch_count(gprn_handle* h);
int64_t gprn_last_rounds(gprn_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* GPRN_B200_H */
