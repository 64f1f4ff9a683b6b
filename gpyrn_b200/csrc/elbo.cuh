// elbo.cuh -- the O(pqN) part of one mean-field iteration: right-hand sides, variational means and
// variances from the factorisation results, the three ELBO terms and the stopping rule.
// Follows gpyrn/meanfield.py:_updateSigMu (:759-865), _expectedLogLike (:923-972),
// _expectedLogPrior (:1019-1065), _entropy (:1085-1093), ELBOcalc loop (:627-649), _initMuVar (:491-510),
// written in the Sigma-free form of SURVEY.md Appendix A.3 (see DESIGN.md "Algorithm").
//
// Workspace slots: the engine (gprn_api.cu: run_pool) keeps `nslot` evaluations in flight; every per-set array
// below is indexed by the SLOT, and slot_set[slot] names the set of the pool (row of `hyper` / `ysub`, entry of
// the result arrays) that the slot currently works on.  "set" in the kernels below is the slot index.
//
// State layout per set (the reference's flat u, meanfield.py:487-488): mu[d], var[d] with
// d = N*q*(p+1); nodes f[j][n] at j*N + n, weights w[i][j][n] at q*N + (i*q + j)*N + n.
// Work vectors per matrix id (= set*M + m; m = j for nodes, q + j*p + i for weights): length Np,
// zero padded.
#pragma once
#include "common.cuh"

namespace gprn {

struct ElboCtx {
    int N, Np, p, q, M, H, d;
    const double* yraw;    // [p][N]
    const double* yerr2;   // [p][N]
    const double* ysub;    // [p][N] shared or [B][p][N] (indexed by the set a slot holds)
    int ysub_shared;
    const double* hyper;   // [B][H] hyper-parameter sets of the whole pool
    const int* slot_set;   // [nslot] pool index of the set each workspace slot currently holds
    const int32_t* par_off;  // [M]
    double *mu, *var;        // [nset][d] current state
    double *mu_new, *var_new;  // [nset][d]
    double *Dv, *bv, *vv, *zv, *uv, *gv;  // [nset*M][Np]
    double *gK;              // [nset*M][Np] diag(K^-1) (q > 1 only)
    double *logdetK, *logdetA;  // [nset*M]
    double *ment, *mlp, *mquad;  // [nset*M] per-matrix entropy / log-prior / quadratic-form pieces
    double *cross_lin;       // [nset] linear part of the cross-node trace (q > 1)
    double *crossbuf;        // [nset][q(q-1)/2][nt*nt] per-tile partials of the cross-node Frobenius norms
    double *hist;            // [nset][3] last three ELBO values
    double *elbo;            // [nset]
    int *iters, *status, *active, *mstatus;  // [nset], [nset], [nset], [nset*M]
    int max_iter;
};

__device__ __forceinline__ const double* hyper_of(const ElboCtx& c, int set) {
    return c.hyper + (size_t)c.slot_set[set] * c.H;
}
__device__ __forceinline__ double variance_at(const ElboCtx& c, int set, int i, int n) {
    double jit = hyper_of(c, set)[c.H - c.p + i];
    return jit * jit + c.yerr2[i * c.N + n];      // meanfield.py:759 (jitters**2 at :618)
}
__device__ __forceinline__ const double* ysub_of(const ElboCtx& c, int set) {
    return c.ysub_shared ? c.ysub : c.ysub + (size_t)c.slot_set[set] * c.p * c.N;
}

// Admission of a set into a workspace slot.  The variational state starts from _initMuVar (meanfield.py:491-510,
// written straight into the flat layout, quirk Q5 included) or, when the chain-state store holds a valid entry
// for the set ('previous', meanfield.py:598-607), from that entry.  Also clears the per-matrix accumulators of the
// slot (log-dets are accumulated by the factorisation kernels).  grid = (number of admitted slots), block = 256.
struct ChainView {        // device view of a chain-state store (all null: no store)
    double* mu;           // [n][d]
    double* var;          // [n][d]
    int* valid;           // [n]
};
__global__ void init_state_kernel(ElboCtx c, const int* __restrict__ slots, ChainView cs) {
    const int set = slots[blockIdx.x];
    const int idx = c.slot_set[set];
    const double* h = hyper_of(c, set);
    double* mu = c.mu + (size_t)set * c.d;
    double* var = c.var + (size_t)set * c.d;
    if (cs.valid && cs.valid[idx]) {
        const double* sm = cs.mu + (size_t)idx * c.d;
        const double* sv = cs.var + (size_t)idx * c.d;
        for (int e = threadIdx.x; e < c.d; e += blockDim.x) {
            mu[e] = sm[e];
            var[e] = sv[e];
        }
    } else {
        double jm = 0.0;
        for (int i = 0; i < c.p; i++) jm += h[c.H - c.p + i];
        jm = jm / c.p;
        for (int e = threadIdx.x; e < c.q * c.N; e += blockDim.x) {
            int j = e / c.N, n = e % c.N;
            double a1 = h[c.par_off[j]];
            double s = 0.0;
            for (int i = 0; i < c.p; i++) {
                double a2 = h[c.par_off[c.q + i]];       // only the first p weight amplitudes (zip truncation)
                double y = c.yraw[i * c.N + n];
                double sg = (y > 0.0) ? 1.0 : ((y < 0.0) ? -1.0 : 0.0);
                s += sqrt((fabs(y) * a1) / a2) * sg;
            }
            mu[e] = s / c.p;
            var[e] = jm;
        }
        for (int e = threadIdx.x; e < c.q * c.p * c.N; e += blockDim.x) {
            int n = e % c.N, ji = e / c.N, j = ji / c.p, i = ji % c.p;   // written in (q,p,N) order
            double a1 = h[c.par_off[j]];
            double a2 = h[c.par_off[c.q + i]];
            double y = c.yraw[i * c.N + n];
            mu[c.q * c.N + e] = sqrt((fabs(y) * a2) / a1);
            var[c.q * c.N + e] = h[c.H - c.p + i];          // jitter, not squared
        }
    }
    if (threadIdx.x < c.M) {
        c.logdetK[(size_t)set * c.M + threadIdx.x] = 0.0;
        c.mstatus[(size_t)set * c.M + threadIdx.x] = 0;
    }
    if (threadIdx.x == 0) {
        c.iters[set] = 0;
        c.status[set] = 0;
        c.active[set] = 1;
    }
}

// Retirement of finished slots: results go to the pool-indexed output arrays; the final state goes back to the
// chain-state store (commit = 2: always; 1: only when the evaluation converged, the reference's caching rule
// meanfield.py:643-646; 0: never).  grid = (number of retired slots), block = 256.
__global__ void retire_kernel(ElboCtx c, const int* __restrict__ slots, double* __restrict__ elbo_out,
                              int* __restrict__ iters_out, int* __restrict__ status_out, int* __restrict__ taken_out,
                              ChainView cs, int commit) {
    const int set = slots[blockIdx.x];
    const int idx = c.slot_set[set];
    const int st = c.status[set];
    if (threadIdx.x == 0) {
        if (elbo_out) elbo_out[idx] = c.elbo[set];
        if (iters_out) iters_out[idx] = c.iters[set];
        if (status_out) status_out[idx] = st;
        if (taken_out) taken_out[idx] = 1;
    }
    if (cs.mu && (commit == 2 || (commit == 1 && st == 0))) {
        const double* mu = c.mu + (size_t)set * c.d;
        const double* var = c.var + (size_t)set * c.d;
        double* dm = cs.mu + (size_t)idx * c.d;
        double* dv = cs.var + (size_t)idx * c.d;
        for (int e = threadIdx.x; e < c.d; e += blockDim.x) {
            dm[e] = mu[e];
            dv[e] = var[e];
        }
        if (threadIdx.x == 0) cs.valid[idx] = 1;
    }
}

// Node right-hand sides (meanfield.py:765, 788-791).  grid = (q, nactive), block = vec_threads.
__global__ void __launch_bounds__(1024) prep_nodes_kernel(ElboCtx c, const int* __restrict__ sets) {
    GPRN_TRACE_SCOPE(TK_OTHER);
    const int set = sets[blockIdx.y], j = blockIdx.x;
    const int N = c.N, q = c.q, p = c.p;
    const double* mu = c.mu + (size_t)set * c.d;
    const double* var = c.var + (size_t)set * c.d;
    const double* muF = mu;
    const double* muW = mu + q * N;
    const double* varW = var + q * N;
    const double* ys = ysub_of(c, set);
    const size_t vo = ((size_t)set * c.M + j) * c.Np;
    for (int n = threadIdx.x; n < c.Np; n += blockDim.x) {
        double D = 0.0, b = 0.0, dd = 0.0;
        if (n < N) {
            for (int i = 0; i < p; i++) {
                double vr = variance_at(c, set, i, n);
                double w = muW[(i * q + j) * N + n];
                dd += (w * w + varW[(i * q + j) * N + n]) / vr;
                double others = 0.0;
                for (int k = 0; k < q; k++)
                    if (k != j) others += muW[(i * q + k) * N + n] * muF[k * N + n];
                b += ((ys[i * N + n] - others) * w) / vr;
            }
            D = 1.0 / dd;
        }
        c.Dv[vo + n] = D;
        c.bv[vo + n] = b;
        c.vv[vo + n] = D * b;
    }
}

// Weight right-hand sides (meanfield.py:838, 847-850, 864).  grid = (q*p, nactive), block = vec_threads.
__global__ void __launch_bounds__(1024) prep_weights_kernel(ElboCtx c, const int* __restrict__ sets) {
    GPRN_TRACE_SCOPE(TK_OTHER);
    const int set = sets[blockIdx.y], ji = blockIdx.x, j = ji / c.p, i = ji % c.p;
    const int N = c.N, q = c.q;
    const double* muFn = c.mu_new + (size_t)set * c.d;
    const double* varFn = c.var_new + (size_t)set * c.d;
    const double* muW = c.mu + (size_t)set * c.d + q * N;
    const double* ys = ysub_of(c, set);
    const size_t vo = ((size_t)set * c.M + q + ji) * c.Np;
    for (int n = threadIdx.x; n < c.Np; n += blockDim.x) {
        double D = 0.0, b = 0.0;
        if (n < N) {
            double vr = variance_at(c, set, i, n);
            double f = muFn[j * N + n];
            double dv = f * f + varFn[j * N + n];
            double others = 0.0;
            for (int k = 0; k < q; k++)
                if (k != j) others += muFn[k * N + n] * muW[(i * q + k) * N + n];
            D = vr / dv;
            b = ((ys[i * N + n] - others) * f) / vr;
        }
        c.Dv[vo + n] = D;
        c.bv[vo + n] = b;
        c.vv[vo + n] = D * b;
    }
}

// After chol(A), X = L^-1, z = X v, u = X^T z, g = colnorm2(X):  mu = v - D u, diag Sigma = D - D^2 g,
// entropy and log-prior pieces of this matrix.  use_identity: quadratic form mu^T K^-1 mu = mu.(b - mu/D)
// (valid when the vector paired with K is the matrix' own mean, i.e. q == 1); otherwise the quadratic
// forms are added by quad_kernel.  grid = (nmat_per_set, nactive), block = vec_threads; first = first matrix index
// of the phase (0 for nodes, q for weights).
__global__ void __launch_bounds__(1024) post_kernel(ElboCtx c, const int* __restrict__ sets, int first, int use_identity) {
    GPRN_TRACE_SCOPE(TK_OTHER);
    __shared__ double red[33];
    const int set = sets[blockIdx.y], m = first + blockIdx.x;
    const int N = c.N, q = c.q, p = c.p;
    const size_t id = (size_t)set * c.M + m;
    const size_t vo = id * c.Np;
    size_t so;   // offset of this matrix' vector inside the flat state
    if (m < q) so = (size_t)m * N;
    else { int ji = m - q, j = ji / p, i = ji % p; so = (size_t)q * N + (size_t)(i * q + j) * N; }
    double* mun = c.mu_new + (size_t)set * c.d + so;
    double* varn = c.var_new + (size_t)set * c.d + so;
    double s_logD = 0.0, s_Dg = 0.0, s_quad = 0.0;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        double D = c.Dv[vo + n], g = c.gv[vo + n], v = c.vv[vo + n], u = c.uv[vo + n];
        double mval = v - D * u;
        mun[n] = mval;
        varn[n] = D - (D * D) * g;
        s_logD += log(D);
        s_Dg += D * g;
        if (use_identity) s_quad += mval * (c.bv[vo + n] - mval / D);
    }
    s_logD = block_sum(s_logD, red);
    s_Dg = block_sum(s_Dg, red);
    s_quad = block_sum(s_quad, red);
    if (threadIdx.x == 0) {
        double ldK = c.logdetK[id], ldA = c.logdetA[id];
        c.ment[id] = 0.5 * (ldK - ldA + s_logD);
        c.mlp[id] = -0.5 * ldK - 0.5 * (s_quad + s_Dg);
        if (c.mstatus[id]) c.status[set] = 1;
    }
}

// -0.5 * ||z||^2 added to the log-prior accumulator (z = X_K m computed by trmv_lower).
// grid = (nmat_per_set, nactive), block = 256.
__global__ void quad_kernel(ElboCtx c, const int* __restrict__ sets, int first) {
    GPRN_TRACE_SCOPE(TK_OTHER);
    __shared__ double red[33];
    const int set = sets[blockIdx.y], m = first + blockIdx.x;
    const size_t vo = ((size_t)set * c.M + m) * c.Np;
    double s = 0.0;
    for (int n = threadIdx.x; n < c.N; n += blockDim.x) s = fma(c.zv[vo + n], c.zv[vo + n], s);
    s = block_sum(s, red);
    if (threadIdx.x == 0) c.mquad[(size_t)set * c.M + m] = -0.5 * s;
}

// Copy the vector that the reference pairs with K_m in the prior's quadratic form into vv (zero padded):
// nodes: mu_f[j];  weights (j,i): row (j*p+i) of the (p,q,N) array viewed as (q,p,N)  (quirk Q4, :1021,1050).
__global__ void gather_quad_vec_kernel(ElboCtx c, const int* __restrict__ sets, int first) {
    const int set = sets[blockIdx.y], m = first + blockIdx.x;
    const size_t vo = ((size_t)set * c.M + m) * c.Np;
    const double* src = c.mu_new + (size_t)set * c.d + (size_t)m * c.N;   // m-th N-vector of the flat state
    for (int n = threadIdx.x; n < c.Np; n += blockDim.x) c.vv[vo + n] = n < c.N ? src[n] : 0.0;
}

// Cross-node trace, linear part (quirk Q3): -0.5 * sum_{k<j} sum_n D_k[n] * diag(K_j^-1)[n].
// grid = (nactive), block = 256.
__global__ void cross_linear_kernel(ElboCtx c, const int* __restrict__ sets) {
    GPRN_TRACE_SCOPE(TK_OTHER);
    __shared__ double red[33];
    const int set = sets[blockIdx.x];
    double s = 0.0;
    for (int j = 1; j < c.q; j++)
        for (int k = 0; k < j; k++) {
            const double* Dk = c.Dv + ((size_t)set * c.M + k) * c.Np;
            const double* gKj = c.gK + ((size_t)set * c.M + j) * c.Np;
            for (int n = threadIdx.x; n < c.N; n += blockDim.x) s = fma(Dk[n], gKj[n], s);
        }
    s = block_sum(s, red);
    if (threadIdx.x == 0) c.cross_lin[set] = -0.5 * s;
}

// Cross-node trace, quadratic part: +0.5 * || X_Kj D_k X_Ak^T ||_F^2 for k < j.
// C[a][b] = sum_{n <= min(a,b)} XK_j[a][n] D_k[n] XA_k[b][n]; one CTA per 64x64 tile of C.
// grid = (nt*nt, npairs, nactive), block = 128, dynamic shared memory 2*TILE_SMEM.
// TR: the factors are stored as transposed tiles (latency path, mid.cuh): the operand tiles are transposed on load.
template <bool TR>
__global__ void __launch_bounds__(128) cross_frob_kernel(ElboCtx c, const double* __restrict__ XK,
                                                         const double* __restrict__ XA,
                                                         const int* __restrict__ sets) {
    GPRN_TRACE_SCOPE(TK_CROSS_FROB);
    extern __shared__ double smem[];
    double* As = smem;
    double* Bs = smem + NB * LDT;
    __shared__ double red[33];
    const int nt = c.Np / NB;
    const int ta = blockIdx.x / nt, tb = blockIdx.x % nt;
    int pair = blockIdx.y, j = 1, k = 0;          // enumerate (j,k), k<j
    while (pair >= j) { pair -= j; j++; }
    k = pair;
    const int set = sets[blockIdx.z];
    const size_t Np = c.Np;
    const double* xk = XK + ((size_t)set * c.M + j) * Np * Np;
    const double* xa = XA + ((size_t)set * c.M + k) * Np * Np;
    const double* Dk = c.Dv + ((size_t)set * c.M + k) * Np;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wm = warp >> 1, wn = warp & 1;
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
    const int kmax = min(ta, tb);
    __shared__ double sc[NB];
    for (int kt = 0; kt <= kmax; kt++) {
        load_tile<TR, false>(As, xk + (size_t)(ta * NB) * Np + kt * NB, Np, tid, 128);
        load_tile<TR, false>(Bs, xa + (size_t)(tb * NB) * Np + kt * NB, Np, tid, 128);
        if (tid < NB) sc[tid] = Dk[kt * NB + tid];
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        mma_tile_scaled(acc, As, Bs, sc, wm, wn, lane);
        __syncthreads();
    }
    // rows/cols >= N belong to the identity padding and must not be counted
    const int r = lane >> 2, cc = lane & 3;
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            int ga = ta * NB + wm * 32 + a * 8 + r, gb = tb * NB + wn * 32 + b * 8 + 2 * cc;
            if (ga < c.N) {
                if (gb < c.N) s = fma(acc[a][b][0], acc[a][b][0], s);
                if (gb + 1 < c.N) s = fma(acc[a][b][1], acc[a][b][1], s);
            }
        }
    s = block_sum(s, red);
    if (tid == 0) c.crossbuf[((size_t)set * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = 0.5 * s;
}

// Likelihood term, ELBO assembly, stopping rule, state commit; loop_cond != 0: the condition of the WHILE node this
// kernel runs in.  grid = (nactive), block = 1024 (fixed: the block sums
// depend on it).
__global__ void __launch_bounds__(1024) elbo_finish_kernel(ElboCtx c, const int* __restrict__ sets,
                                                           cudaGraphConditionalHandle loop_cond) {
    GPRN_TRACE_SCOPE(TK_OTHER);
    __shared__ double red[33];
    const int set = sets[blockIdx.x];
    const int N = c.N, q = c.q, p = c.p;
    const double* muFn = c.mu_new + (size_t)set * c.d;
    const double* varFn = c.var_new + (size_t)set * c.d;
    const double* muWn = muFn + q * N;
    const double* varWn = varFn + q * N;
    double s_log = 0.0, s_res = 0.0, s_val = 0.0;
    for (int e = threadIdx.x; e < p * N; e += blockDim.x) {
        int i = e / N, n = e % N;
        double vr = variance_at(c, set, i, n);
        s_log += log((2.0 * M_PI) * vr);                      // :925
        double om = 0.0;
        for (int j = 0; j < q; j++) om += muWn[(i * q + j) * N + n] * muFn[j * N + n];
        double rs = c.yraw[i * N + n] - om;                   // raw y (quirk Q1, :940)
        s_res += (rs * rs) / vr;
        for (int j = 0; j < q; j++) {
            double sf = varFn[j * N + n], sw = varWn[(i * q + j) * N + n];
            double mw = muWn[(i * q + j) * N + n], mf = muFn[j * N + n];
            s_val += (sf * (mw * mw) + sw * (mf * mf) + sf * sw) / vr;   // :968-971
        }
    }
    s_log = block_sum(s_log, red);
    s_res = block_sum(s_res, red);
    s_val = block_sum(s_val, red);
    // per-matrix pieces summed in a fixed order (deterministic: no floating-point atomics on this path)
    double s_cross = 0.0;
    if (q > 1) {
        const int ncross = (q * (q - 1) / 2) * (c.Np / NB) * (c.Np / NB);
        const double* cb = c.crossbuf + (size_t)set * ncross;
        for (int e = threadIdx.x; e < ncross; e += blockDim.x) s_cross += cb[e];
        s_cross = block_sum(s_cross, red);
    }
    // commit the new state (the reference carries new_mu/new_var into the next iteration, :636)
    const int commit = c.max_iter > 0;
    if (commit) {
        double* mu = c.mu + (size_t)set * c.d;
        double* var = c.var + (size_t)set * c.d;
        for (int e = threadIdx.x; e < c.d; e += blockDim.x) {
            mu[e] = muFn[e];
            var[e] = varFn[e];
        }
    }
    if (threadIdx.x == 0) {
        const double LOG2PI = log(2.0 * M_PI);
        const double MN = (double)c.M * (double)N;
        double ll = -0.5 * s_log - 0.5 * s_res - 0.5 * s_val;
        double lp = 0.0, ent = 0.0;
#pragma unroll 4
        for (int m = 0; m < c.M; m++) {
            const size_t id = (size_t)set * c.M + m;
            ent += c.ment[id];
            lp += c.mlp[id];
            if (q > 1) lp += c.mquad[id];
        }
        if (q > 1) lp += c.cross_lin[set] + s_cross;
        lp -= 0.5 * MN * LOG2PI;                     // :1064
        ent += 0.5 * MN * (1.0 + LOG2PI);            // :1092
        double elbo = (ll + lp + ent) / q;                                       // :709
        if (c.status[set] == 1) elbo = nan("");
        c.elbo[set] = elbo;
        double* hs = c.hist + set * 3;
        hs[0] = hs[1]; hs[1] = hs[2]; hs[2] = elbo;
        int it = c.iters[set];
        if (commit) it += 1;
        c.iters[set] = it;
        int done = 0;
        if (!commit) { done = 1; c.status[set] = 2; }      // max_iter == 0: the loop body never runs (:634,648)
        else if (c.status[set] == 1) done = 1;            // not PD: stop at once instead of spinning to max_iter
        else {
            if (it > 3) {                                   // :640-646
                double mean = ((hs[0] + hs[1]) + hs[2]) / 3.0;
                double d0 = hs[0] - mean, d1 = hs[1] - mean, d2 = hs[2] - mean;
                double sd = sqrt(((d0 * d0 + d1 * d1) + d2 * d2) / 3.0);
                double crit = fabs(sd / mean);
                if (crit < 1e-3 && crit != 0.0) done = 1;
            }
            if (!done && it >= c.max_iter) { done = 1; c.status[set] = 2; }
        }
        if (done) {
            c.active[set] = 0;
            // device-resident loop (iteration_loop_graph): the first finished set hands control back to the host
            if (loop_cond) cudaGraphSetConditional(loop_cond, 0);
        }
    }
}

}  // namespace gprn
