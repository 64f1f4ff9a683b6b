"""Mean-field variational inference for GPRNs (Nguyen & Bonilla 2013) on a B200.

Host-side mirror of the reference's ``gpyrn.meanfield.inference`` (gpyrn/meanfield.py:92-1403): same
constructor, ``set_components``, parameter-vector bookkeeping, ``ELBO`` / ``ELBOcalc`` / ``nELBO`` /
``optimize`` / ``_Prediction`` / ``predict`` signatures and return shapes.  Everything numerical --
covariance assembly, Cholesky factorisations, the variational mu/Sigma updates, the ELBO terms, the
GP predictive -- runs in the CUDA library behind the C ABI of ``include/gprn_b200.h``; this module
only keeps the Python objects, evaluates the mean functions (O(pN), host by design) and marshals
small arrays.  There is no CPU fallback.

Additions over the reference: ``ELBO_batch(parameters[B, n])`` evaluates B hyper-parameter sets in one
device call (the data-parallel entry used for sweeps / MCMC walkers) with an optional device-resident per-chain
warm-start state, ``optimize_batch`` runs many Nelder-Mead starts in lock-step on top of it, ``logposterior_batch`` is
the vectorised log-posterior for ensemble samplers, ``Prediction_batch`` predicts for a chain of hyper-parameter sets,
and ``device=`` selects the GPU.
"""
import ctypes
import threading
import time as time_module
from itertools import chain

import numpy as np

from . import _lib, covfunc, meanfunc


_chol_handles = {}                       # device ordinal -> scratch `inference` whose handle runs _cholNugget
_chol_lock = threading.Lock()


def _cholNugget(matrix, device=0):
    """
    Cholesky decomposition of a symmetric positive definite matrix (reference ``_cholNugget``, meanfield.py:71-88).

    The factorisation runs in the blocked device kernels of ``csrc/factor.cuh`` -- the ones every ELBO iteration and
    prediction uses -- through ``gprn_debug_factor``; only the lower triangle of ``matrix`` is
    read.  As in the reference no nugget is added, and a matrix that is not positive definite gives a NaN factor
    (what ``jnp.linalg.cholesky`` returns) instead of raising.

    Returns:
        L (n, n) lower-triangular factor, nugget (always 0.0)
    """
    A = _lib.f64(matrix)
    if A.ndim != 2 or A.shape[0] != A.shape[1] or A.shape[0] < 1:
        raise ValueError(f'_cholNugget: expected a square matrix, got shape {A.shape}')
    n = A.shape[0]
    L = np.empty((n, n))
    with _chol_lock:
        g = _chol_handles.get(device)
        if g is None:
            g = _chol_handles[device] = inference(1, np.arange(4.0), np.zeros(4), np.ones(4), device=device)
        rc = _lib.lib().gprn_debug_factor(g._h(), n, _lib.dptr(A), _lib.dptr(L), None, None)
        if rc != 0:
            msg = _lib.lib().gprn_last_error().decode()
            if 'not positive definite' not in msg:
                raise _lib.GprnError(msg)
            L.fill(np.nan)
    return L, 0.0


class inference:
    """
    Mean-field variational inference for GPRNs.

    Args:
        q: number of latent node functions
        time: time coordinates, shape (N,)
        *args: observed data  y1, y1error, y2, y2error, ...
        device: CUDA device ordinal (keyword only, default 0)
    """

    def __init__(self, q: int, time, *args, device: int = 0):
        self.q = q
        self.time = time
        self.N = self.time.size
        msg = 'Number of observed data arrays should be even: y1, y1error, ...'
        assert len(args) > 0 and len(args) % 2 == 0, msg
        msg = 'Output arrays should all have the same dimensions as time'
        assert np.all(np.array([len(a) for a in args]) == self.N), msg
        self.p = len(args) // 2
        self.qp = self.q * self.p
        self.d = self.N * self.q * (self.p + 1)
        self.tt = np.tile(time, self.p)
        self.y = np.array(args[::2], dtype=float)
        self.yerr = np.array(args[1::2], dtype=float)
        self.yerr2 = self.yerr ** 2
        self._components_set = False
        self._frozen_mask = np.array([])
        self._mu, self._var = None, None
        self._mu_var_iters = 0
        self.update_muvar_after = 50
        self.elbo_max_iter = 5000
        self.device = device
        self._handle = None
        self._model_sig = None

    # ------------------------------------------------------------------------------------------
    # device handle
    # ------------------------------------------------------------------------------------------
    def _h(self):
        if self._handle is None:
            h = ctypes.c_void_p()
            t = _lib.f64(self.time)
            y = _lib.f64(self.y)
            e = _lib.f64(self.yerr)
            _lib.check(_lib.lib().gprn_create(self.device, self.N, self.p, self.q, _lib.dptr(t), _lib.dptr(y),
                                              _lib.dptr(e), ctypes.byref(h)))
            self._handle = h
        return self._handle

    def close(self):
        """Release the device workspace."""
        if self._handle is not None:
            _lib.lib().gprn_destroy(self._handle)
            self._handle = None
            self._model_sig = None
            self._n_chains = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _bind_model(self, nodes, weights):
        """Serialise kernel structure (not values) and hand it to the library when it changed."""
        progs_n = [k.program() for k in nodes]
        progs_w = [k.program() for k in weights]
        sig = (tuple(map(tuple, progs_n)), tuple(map(tuple, progs_w)))
        n_hyper = int(sum(k.pars.size for k in chain(nodes, weights)) + self.p)
        if sig != self._model_sig:
            def pack(progs):
                off = np.zeros(len(progs) + 1, dtype=np.int32)
                off[1:] = np.cumsum([len(p) for p in progs])
                return np.array([t for p in progs for t in p], dtype=np.int32), off
            pn, on = pack(progs_n)
            pw, ow = pack(progs_w)
            _lib.check(_lib.lib().gprn_set_model(self._h(), _lib.iptr(pn), _lib.iptr(on), _lib.iptr(pw),
                                                 _lib.iptr(ow), n_hyper))
            self._model_sig = sig
        return n_hyper

    @staticmethod
    def _hyper_vector(nodes, weights, jitters):
        parts = [np.asarray(k.pars, dtype=float).ravel() for k in chain(nodes, weights)]
        parts.append(np.asarray(jitters, dtype=float).ravel())
        return np.concatenate(parts)

    # ------------------------------------------------------------------------------------------
    # components and parameters (host bookkeeping, reference meanfield.py:136-379)
    # ------------------------------------------------------------------------------------------
    def set_components(self, nodes, weights, means, jitters):
        """Set the q nodes, q*p weights (index j*p+i: node j -> output i), p means and p jitters."""
        if isinstance(nodes, covfunc.covFunction):
            nodes = [nodes]
        if len(nodes) != self.q:
            raise ValueError(f'Wrong number of nodes provided, expected {self.q} got {len(nodes)}')
        if isinstance(weights, covfunc.covFunction):
            weights = [weights]
        if len(weights) != self.qp:
            raise ValueError(f'Wrong number of weights provided, expected {self.qp} got {len(weights)}')
        if isinstance(means, (int, float, meanfunc.meanFunction)):
            means = [means]
        if isinstance(jitters, (int, float)):
            jitters = [jitters]
        self.nodes = nodes
        self.weights = weights
        self.means = means
        self.jitters = np.array(jitters, dtype=float)
        self._components_set = True

    def _require_components(self):
        assert self._components_set, 'GPRN components not set, use set_components'

    def get_parameters(self, nodes=None, weights=None, means=None, jitters=None, include_frozen=False):
        """Values of all GPRN parameters: nodes, weights, means, jitters (in that order)."""
        given = [nodes, weights, means, jitters]
        if not self._components_set and all(g is None for g in given):
            raise ValueError('Cannot get parameters. Provide arguments or run set_components before.')
        if self._components_set:
            nodes, weights, means, jitters = self.nodes, self.weights, self.means, self.jitters
        chunks = []
        for group in (nodes, weights, means):
            if group is not None:
                chunks += [c.get_parameters() for c in group]
        if jitters is not None:
            chunks += [np.array([j]) for j in jitters]
        flat = np.concatenate(chunks).ravel()
        return flat if include_frozen else flat[~self.frozen_mask]

    def set_parameters(self, parameters):
        """Set values for all (or all non-frozen) GPRN parameters."""
        self._require_components()
        parameters = np.atleast_1d(np.array(parameters, dtype=float))
        current = self.get_parameters(include_frozen=True)
        n_all = self.n_parameters
        n_free = n_all - int(self.frozen_mask.sum())
        if parameters.size == n_all:
            full = parameters.copy()
            full[self.frozen_mask] = current[self.frozen_mask]
        elif parameters.size == n_free:
            full = current.copy()
            full[~self.frozen_mask] = parameters
        else:
            msg = f'Wrong number of parameters provided: got {parameters.size}, '
            msg += f'expected {n_all}' if n_all == n_free else f'expected {n_all} (all) or {n_free} (not frozen)'
            raise ValueError(msg)
        rest = full
        for component in chain(self.nodes, self.weights, self.means):
            rest = component.set_parameters(rest)
        self.jitters = rest

    @property
    def n_parameters(self):
        """Total number of parameters."""
        self._require_components()
        n = sum(c.pars.size for c in chain(self.nodes, self.weights, self.means))
        return n + self.jitters.size

    @property
    def parameters_dict(self):
        """Dictionary with parameter names and values."""
        self._require_components()
        out = {}
        for label, group in (('node', self.nodes), ('weight', self.weights), ('mean', self.means)):
            for i, comp in enumerate(group, start=1):
                for name, val in zip(comp._param_names, comp.pars):
                    out[f'{label}{i}.{name}'] = val
        for i, jit in enumerate(self.jitters, start=1):
            out[f'jitter{i}'] = jit
        return out

    def _set_frozen(self, value, index, name):
        self.frozen_mask
        if index is None and name is None:
            raise ValueError('Provide either index or name')
        if name is None:
            self._frozen_mask[index] = value
            return
        names = list(self.parameters_dict.keys())
        if '*' in name:
            stem = name.replace('*', '')
            for i, known in enumerate(names):
                if stem in known:
                    self._frozen_mask[i] = value
        else:
            assert name in names, f'Name "{name}" not found in parameters_dict'
            self._frozen_mask[names.index(name)] = value

    def freeze_parameter(self, index=None, name=None):
        """Freeze (do not fit for) a parameter by index or name ('node1*' freezes all of node 1)."""
        self._set_frozen(True, index, name)

    def thaw_parameter(self, index=None, name=None):
        """Thaw (free) a parameter by index or name."""
        self._set_frozen(False, index, name)

    def freeze_all_parameters(self):
        self._frozen_mask = np.ones(self._frozen_mask.size, dtype=bool)

    def thaw_all_parameters(self):
        self._frozen_mask = np.zeros(self._frozen_mask.size, dtype=bool)

    fix_parameter = freeze_parameter
    fix_all_parameters = freeze_all_parameters
    free_parameter = thaw_parameter
    free_all_parameters = thaw_all_parameters

    @property
    def frozen_mask(self):
        """Boolean mask for the frozen parameters."""
        self._require_components()
        if self._frozen_mask.size == 0:
            self._frozen_mask = np.full(self.n_parameters, False, dtype=bool)
        return self._frozen_mask

    @frozen_mask.setter
    def frozen_mask(self, mask):
        raise NotImplementedError('Do not set frozen_mask, use thaw_parameter/freeze_parameter')

    def _get_components(self, nodes=None, weights=None, means=None, jitters=None):
        if all(i is None for i in (nodes, weights, means, jitters)) and not self._components_set:
            raise ValueError('GPRN components not set, use set_components')
        nodes = self.nodes if nodes is None else nodes
        weights = self.weights if weights is None else weights
        means = self.means if means is None else means
        jitters = self.jitters if jitters is None else jitters
        if isinstance(nodes, covfunc.covFunction):
            nodes = [nodes]
        if isinstance(weights, covfunc.covFunction):
            weights = [weights]
        return nodes, weights, means, jitters

    # ------------------------------------------------------------------------------------------
    # host pieces of the path
    # ------------------------------------------------------------------------------------------
    def _mean(self, means, time=None):
        """Concatenated mean-function values (p*T,), zeros where a mean is None (reference :382-411)."""
        t = self.time if time is None else time
        out = np.zeros(self.p * t.size)
        for i, mf in enumerate(means):
            if mf is None:
                continue
            if isinstance(mf, (int, float)):
                out[i * t.size:(i + 1) * t.size] = mf
            else:
                out[i * t.size:(i + 1) * t.size] = mf(t)
        return out

    def _kmat(self, kernel, t_rows, t_cols, nugget):
        prog = np.array(kernel.program(), dtype=np.int32)
        pars = _lib.f64(kernel.pars)
        tr = _lib.f64(t_rows)
        if t_cols is None:
            out = np.empty((tr.size, tr.size))
            tc, nc = None, 0
        else:
            tc = _lib.f64(t_cols)
            nc = tc.size
            out = np.empty((tr.size, tc.size))
        _lib.check(_lib.lib().gprn_kmatrix(self._h(), _lib.iptr(prog), prog.size, _lib.dptr(pars), pars.size,
                                           _lib.dptr(tr), tr.size, _lib.dptr(tc), nc, nugget, _lib.dptr(out), None))
        return out

    def _KMatrix(self, kernel, time=None):
        """k(t - t^T) + 1e-6 I, assembled on the device (reference :413-434)."""
        return self._kmat(kernel, self.time if time is None else time, None, 1e-6)

    def _tinyNuggetKMatrix(self, kernel, time=None):
        """k(t - t^T) + 1.25e-12 I (reference :436-453)."""
        return self._kmat(kernel, self.time if time is None else time, None, 1.25e-12)

    def _predictKMatrix(self, kernel, time):
        """k(time - self.time^T) without nugget (reference :455-471)."""
        return self._kmat(kernel, np.atleast_1d(time), self.time, 0.0)

    def _u_to_fhatW(self, u):
        """Flat variational vector -> nodes (1,q,N) and weights (p,q,N) (reference :473-489)."""
        f = u[:self.q * self.N].reshape((1, self.q, self.N))
        w = u[self.q * self.N:].reshape((self.p, self.q, self.N))
        return f, w

    def _initMuVar(self, nodes, weights, jitter):
        """Initial variational means / variances (reference :491-510), computed by the device kernel."""
        _, mu, var, _, _ = self._run_elbo(nodes, weights, None, jitter, 0, None, None)
        return mu.ravel(), var.ravel()

    def _randomMuVar(self):
        return np.random.randn(self.d, 1), np.random.rand(self.d, 1)

    def sample(self, time=None, z=None, nugget=1.25e-12):
        """
        Draws from the GP priors of the nodes and weights on ``self.time`` (reference :517-539; like the reference,
        the ``time`` argument is accepted and ignored).

        Args:
            z: optional standard-normal variates, array (q + q*p, N), nodes first; default ``np.random.randn``
            nugget: diagonal term (reference ``_tinyNuggetKMatrix``: 1.25e-12).  The draw is ``chol(K + nugget I) z``
                on the device; the reference's eigen-decomposition route (``allow_singular=True``) is the same
                distribution.  A kernel matrix that is not numerically positive definite raises ``GprnError``.

        Returns:
            node_samples (q, N), weight_samples (q*p, N)
        """
        nodes, weights, _, jitters = self._get_components()
        self._bind_model(nodes, weights)
        M = self.q + self.qp
        z = np.random.randn(M, self.N) if z is None else np.asarray(z, dtype=float).reshape(M, self.N)
        hyper = _lib.f64(self._hyper_vector(nodes, weights, jitters))
        out = np.empty((M, self.N))
        _lib.check(_lib.lib().gprn_sample(self._h(), _lib.dptr(hyper), _lib.dptr(_lib.f64(z)), float(nugget),
                                          _lib.dptr(out), None))
        return out[:self.q].copy(), out[self.q:].copy()

    def _sample_from_gp(self, kernel, time=None, z=None):
        """One draw from ``kernel`` on ``time`` (default ``self.time``; reference :517-530)."""
        time = self.time if time is None else np.atleast_1d(np.asarray(time, dtype=float))
        g = inference(1, time, np.zeros_like(time), np.ones_like(time), device=self.device)
        try:
            g.set_components(kernel, covfunc.WhiteNoise(1.0), meanfunc.Constant(0.0), 0.1)   # weight: K = I, unused
            nodes, _ = g.sample(z=None if z is None else np.vstack([np.asarray(z, dtype=float).ravel(),
                                                                    np.zeros(time.size)]))
        finally:
            g.close()
        return nodes[0]

    # ------------------------------------------------------------------------------------------
    # ELBO
    # ------------------------------------------------------------------------------------------
    def _run_elbo(self, nodes, weights, means, jitters, max_iter, mu0, var0):
        """One device evaluation.  Returns (elbo, mu[1+p,q,N], var[1+p,q,N], iters, status)."""
        self._bind_model(nodes, weights)
        hyper = _lib.f64(self._hyper_vector(nodes, weights, jitters))
        if means is None:
            ysub = _lib.f64(self.y)
        else:
            ysub = _lib.f64(np.concatenate(self.y) - self._mean(means)).reshape(self.p, self.N)
        mu = np.empty(self.d)
        var = np.empty(self.d)
        init_mode = 0
        if mu0 is not None:
            mu[:] = np.asarray(mu0, dtype=float).ravel()
            var[:] = np.asarray(var0, dtype=float).ravel()
            init_mode = 1
        elbo = np.empty(1)
        iters = np.zeros(1, dtype=np.int32)
        status = np.zeros(1, dtype=np.int32)
        _lib.check(_lib.lib().gprn_elbo_batched(self._h(), 1, _lib.dptr(hyper), _lib.dptr(ysub), 1, init_mode,
                                                _lib.dptr(mu), _lib.dptr(var), -1 if max_iter is None else max_iter,
                                                _lib.dptr(elbo), _lib.iptr(iters), _lib.iptr(status), None))
        shape = (1 + self.p, self.q, self.N)
        return float(elbo[0]), mu.reshape(shape), var.reshape(shape), int(iters[0]), int(status[0])

    @property
    def ELBO(self):
        """The evidence lower bound for the GPRN."""
        return self.ELBOcalc()[0]

    def ELBOcalc(self, nodes=None, weights=None, means=None, jitters=None, max_iter=None, mu=None, var=None):
        """
        Calculate the evidence lower bound by iterating the closed-form variational updates.

        Args:
            nodes, weights, means, jitters: components (default: those given to ``set_components``)
            max_iter: maximum number of iterations (default 10000)
            mu, var: 'init' (default), 'random', 'previous', or arrays of size d with the initial
                variational means / variances

        Returns:
            ELBO, mu (1+p, q, N), var (1+p, q, N), number of iterations
        """
        nodes, weights, means, jitters = self._get_components(nodes, weights, means, jitters)
        mu0 = var0 = None
        if mu is None or var is None:
            mu = var = 'init'
        if isinstance(mu, str) or isinstance(var, str):
            if mu == 'previous' or var == 'previous':
                if self._mu is not None:
                    mu0, var0 = self._mu, self._var
            elif mu == 'random' and var == 'random':
                mu0, var0 = self._randomMuVar()
            elif not (mu == 'init' and var == 'init'):
                raise ValueError("mu and var must be 'init', 'random', 'previous' or arrays")
        else:
            mu0, var0 = mu, var
        elbo, mu_out, var_out, it, status = self._run_elbo(nodes, weights, means, jitters, max_iter, mu0, var0)
        if status == 2:
            print('\nMax iterations reached')
        elif status == 0:
            self._mu, self._var = mu_out, var_out
        return elbo, mu_out, var_out, it

    def ELBOaux(self, Kf=None, Kw=None, Lf=None, Lw=None, y=None, jitt2=None, mu=None, var=None):
        """One fixed-point iteration from (mu, var).  The matrix arguments of the reference signature are
        accepted and ignored (the device rebuilds them from the current components); the Sigma outputs
        are None because the device path never forms them (Sigma-free updates)."""
        nodes, weights, means, jitters = self._get_components()
        if jitt2 is not None:
            jitters = np.sqrt(np.asarray(jitt2, dtype=float))
        self._bind_model(nodes, weights)
        elbo, mu_out, var_out, _, _ = self._run_elbo(nodes, weights, means, jitters, 1, mu, var)
        return elbo, mu_out, var_out, None, None

    def _full_parameter_matrix(self, parameters):
        """(B, n_parameters) matrix from rows holding all parameters or only the free ones (the rule of
        ``set_parameters``: frozen entries keep their current values either way)."""
        P = np.atleast_2d(np.asarray(parameters, dtype=float))
        n_all = self.n_parameters
        frozen = self.frozen_mask
        n_free = n_all - int(frozen.sum())
        current = self.get_parameters(include_frozen=True)
        if P.shape[1] == n_all:
            full = P.copy()
            full[:, frozen] = current[frozen]
        elif P.shape[1] == n_free:
            full = np.tile(current, (P.shape[0], 1))
            full[:, ~frozen] = P
        else:
            msg = f'Wrong number of parameters provided: got {P.shape[1]}, '
            msg += f'expected {n_all}' if n_all == n_free else f'expected {n_all} (all) or {n_free} (not frozen)'
            raise ValueError(msg)
        return full

    def _split_parameter_matrix(self, P):
        """Full parameter rows -> (hyper [B,H] = kernel parameters + jitters, y - mean [p*N] or [B,p*N], shared flag)."""
        B = P.shape[0]
        n_kernel = sum(k.pars.size for k in chain(self.nodes, self.weights))
        n_mean = sum(m.pars.size for m in self.means if isinstance(m, meanfunc.meanFunction))
        hyper = _lib.f64(np.concatenate([P[:, :n_kernel], P[:, n_kernel + n_mean:]], axis=1))
        saved = [m.pars.copy() if isinstance(m, meanfunc.meanFunction) else None for m in self.means]
        try:
            if n_mean == 0 or np.all(P[:, n_kernel:n_kernel + n_mean] == P[0, n_kernel:n_kernel + n_mean]):
                self._assign_means(P[0, n_kernel:n_kernel + n_mean])
                ysub = _lib.f64(np.concatenate(self.y) - self._mean(self.means))
                shared = 1
            else:
                ysub = np.empty((B, self.p * self.N))
                for b in range(B):
                    self._assign_means(P[b, n_kernel:n_kernel + n_mean])
                    ysub[b] = np.concatenate(self.y) - self._mean(self.means)
                ysub = _lib.f64(ysub)
                shared = 0
        finally:
            self._restore_means(saved)
        return hyper, ysub, shared

    def ELBO_batch(self, parameters, max_iter=None, return_info=False, mu=None, var=None, return_state=False,
                   state=None, slots=0, work_source=None):
        """
        ELBO of B hyper-parameter sets in one device call (continuous batching over the GPU's workspace slots).

        Args:
            parameters: array (B, n_parameters) or (B, number of free parameters), ``get_parameters`` order
            max_iter: per-evaluation iteration cap (default 10000)
            return_info: also return (iterations, status) arrays
            mu, var: optional per-set initial variational state, host arrays (B, d); default 'init'
            return_state: also return the final (mu, var), host arrays (B, d)
            state: ``'previous'`` keeps one variational state per row ON THE DEVICE between calls (row b <-> chain
                b): an evaluation starts from its chain's last converged state, or from 'init' when there is none,
                and stores its own state when it converges -- the batched form of ``nELBO``'s warm start
                (reference :598-607, :643-646), without moving B*d doubles per call.  ``'previous-always'`` stores
                the state whether or not the evaluation converged.  ``reset_chain_state()`` forgets all chains.
            slots: evaluations kept in flight (0: as many as fit the device workspace)
            work_source: optional callable returning the next row index to evaluate, or -1 when none is left
                (several processes sharing one counter split the rows dynamically; see
                ``gpyrn_b200.distributed``).  Rows not handed out by it are not evaluated: their outputs are 0 and
                the extra return value ``taken`` (bool, (B,)) is False.

        Returns:
            elbo (B,) [, iterations (B,), status (B,)] [, mu (B, d), var (B, d)] [, taken (B,)]
        """
        self._require_components()
        P = self._full_parameter_matrix(parameters)
        B = P.shape[0]
        self._bind_model(self.nodes, self.weights)
        hyper, ysub, shared = self._split_parameter_matrix(P)
        elbo = np.empty(B)
        iters = np.zeros(B, dtype=np.int32)
        status = np.zeros(B, dtype=np.int32)
        L = _lib.lib()
        mi = -1 if max_iter is None else max_iter
        if state is None and work_source is None and not slots:
            init_mode, mu_io, var_io = 0, None, None
            if mu is not None or var is not None:
                if mu is None or var is None:
                    raise ValueError('provide both mu and var, or neither')
                mu_io = _lib.f64(np.array(mu, dtype=float).reshape(B, self.d))
                var_io = _lib.f64(np.array(var, dtype=float).reshape(B, self.d))
                init_mode = 1
            elif return_state:
                mu_io, var_io = np.empty((B, self.d)), np.empty((B, self.d))
            _lib.check(L.gprn_elbo_batched(self._h(), B, _lib.dptr(hyper), _lib.dptr(ysub), shared, init_mode,
                                           _lib.dptr(mu_io), _lib.dptr(var_io), mi, _lib.dptr(elbo),
                                           _lib.iptr(iters), _lib.iptr(status), None))
            out = (elbo, iters, status) if return_info else (elbo,)
            if return_state:
                out = out + (mu_io, var_io)
            return out if len(out) > 1 else out[0]
        # pool entry: device-resident chain state and / or an external work source
        if state not in (None, 'previous', 'previous-always'):
            raise ValueError("state must be None, 'previous' or 'previous-always'")
        if (mu is None) != (var is None):
            raise ValueError('provide both mu and var, or neither')
        state_mode = {None: 0, 'previous': 1, 'previous-always': 2}[state]
        if state_mode == 0 and (mu is not None or return_state):
            state_mode = 2                          # host state in / out goes through the store as well
            self.reset_chain_state()
        if state_mode:
            self._ensure_chain_state(B)
            if mu is not None:
                _lib.check(L.gprn_chain_set(self._h(), 0, B, _lib.dptr(_lib.f64(np.array(mu).reshape(B, self.d))),
                                            _lib.dptr(_lib.f64(np.array(var).reshape(B, self.d)))))
        taken = np.zeros(B, dtype=np.int32)
        cb = None
        if work_source is not None:
            cb = _lib.NEXT_SET_FN(lambda _user: int(work_source()))
        _lib.check(L.gprn_elbo_pool(self._h(), B, hyper.ctypes.data, 0, _lib.dptr(ysub), shared,
                                    ctypes.cast(cb, ctypes.c_void_p) if cb is not None else None, None, int(slots),
                                    state_mode, mi, elbo.ctypes.data, iters.ctypes.data, status.ctypes.data,
                                    taken.ctypes.data, 0, None))
        out = (elbo, iters, status) if return_info else (elbo,)
        if return_state:
            mu_o, var_o = np.empty((B, self.d)), np.empty((B, self.d))
            _lib.check(L.gprn_chain_get(self._h(), 0, B, _lib.dptr(mu_o), _lib.dptr(var_o), None))
            out = out + (mu_o, var_o)
        if work_source is not None:
            out = out + (taken.astype(bool),)
        return out if len(out) > 1 else out[0]

    def _ensure_chain_state(self, B):
        if getattr(self, '_n_chains', 0) != B:
            _lib.check(_lib.lib().gprn_chain_resize(self._h(), B))
            self._n_chains = B

    def reset_chain_state(self):
        """Forget the device-resident per-chain variational states: the next ``ELBO_batch(state='previous')``
        starts every row from 'init'."""
        if getattr(self, '_n_chains', 0):
            _lib.check(_lib.lib().gprn_chain_invalidate(self._h(), 0, self._n_chains))

    def get_chain_state(self):
        """Host copy of the device-resident chain states: mu (B, d), var (B, d), valid (B,) bool."""
        B = getattr(self, '_n_chains', 0)
        if not B:
            raise ValueError("no chain state: call ELBO_batch(..., state='previous') first")
        mu, var, valid = np.empty((B, self.d)), np.empty((B, self.d)), np.zeros(B, dtype=np.int32)
        _lib.check(_lib.lib().gprn_chain_get(self._h(), 0, B, _lib.dptr(mu), _lib.dptr(var), _lib.iptr(valid)))
        return mu, var, valid.astype(bool)

    def _assign_means(self, values):
        rest = np.asarray(values, dtype=float)
        for m in self.means:
            if isinstance(m, meanfunc.meanFunction) and m.pars.size:
                n = m.pars.size
                m._assign(rest[:n])
                rest = rest[n:]

    def _restore_means(self, saved):
        for m, s in zip(self.means, saved):
            if s is not None:
                m._assign(s)

    def nELBO(self, parameters, max_iter=None):
        """Negative ELBO for given values of the (free) parameters; warm-started from the last converged state."""
        self._require_components()
        self.set_parameters(parameters)
        start = time_module.time()
        elbo, _, _, _ = self.ELBOcalc(self.nodes, self.weights, self.means, self.jitters, max_iter=max_iter,
                                      mu='previous', var='previous')
        end = time_module.time()
        print(f'ELBO={elbo:7.2f} (took {1e3 * (end - start):5.2f} ms){20 * " "}', end='\r', flush=True)
        return -elbo

    def _select_vars(self, vars):
        if vars is None:
            return
        if isinstance(vars, str):
            if '-' in vars:
                self.thaw_parameter(name='*')
                self.freeze_parameter(name=vars.replace('-', ''))
            else:
                self.freeze_parameter(name='*')
                self.thaw_parameter(name=vars)
        elif isinstance(vars, list):
            self.freeze_parameter(name='*')
            for v in vars:
                self.thaw_parameter(name=v)
        else:
            raise ValueError(f'`vars` should be str or list, got {type(vars)}')

    def optimize(self, vars=None, **kwargs):
        """Maximise the ELBO with scipy.optimize.minimize (Nelder-Mead by default); ``vars`` selects the
        free parameters ('name', '-name' or a list of names)."""
        from scipy.optimize import minimize
        self._select_vars(vars)
        kwargs.setdefault('method', 'Nelder-Mead')
        res = minimize(self.nELBO, self.get_parameters(), **kwargs)
        self.set_parameters(res.x)
        return res

    def nELBO_batch(self, parameters, max_iter=None):
        """Negative ELBO of B parameter rows, each warm-started from its own chain's last converged state kept on the
        device (row b <-> chain b): the batched ``nELBO`` (reference :1095-1111).  The object's own parameters are
        left untouched."""
        return -self.ELBO_batch(parameters, max_iter=max_iter, state='previous')

    def optimize_batch(self, starts, vars=None, max_iter=None, **kwargs):
        """
        Multi-start optimisation: S independent ``scipy.optimize.minimize`` runs (Nelder-Mead by default, exactly the
        driver of ``optimize``, reference :1114-1152) advanced in LOCK-STEP -- every round, the one objective
        evaluation each live start is waiting for goes to the GPU as a single ``ELBO_batch`` call, each start
        warm-started from its own device-resident variational state (what ``nELBO``'s ``'previous'`` does for a
        single run).  The optimiser logic is scipy's own, run as S coroutines, so a start returns what a sequential
        ``optimize`` from the same point would.

        Args:
            starts: array (S, number of free parameters) of initial points
            vars: as ``optimize``
            max_iter: iteration cap of every ELBO evaluation (default 10000)
            **kwargs: passed to ``scipy.optimize.minimize``

        Returns:
            list of S ``OptimizeResult``; the object's parameters are set to the best one.
        """
        from scipy.optimize import minimize
        self._require_components()
        self._select_vars(vars)
        kwargs.setdefault('method', 'Nelder-Mead')
        X0 = np.atleast_2d(np.asarray(starts, dtype=float))
        S = X0.shape[0]
        n_free = self.n_parameters - int(self.frozen_mask.sum())
        if X0.shape[1] != n_free:
            raise ValueError(f'starts must have {n_free} columns (the free parameters), got {X0.shape[1]}')
        cond = threading.Condition()
        request = [None] * S           # point a start is waiting on
        value = [None] * S
        done = [False] * S
        results = [None] * S
        errors = []
        self.reset_chain_state()
        self._ensure_chain_state(S)

        def objective(s):
            def f(x):
                with cond:
                    request[s] = np.array(x, dtype=float)
                    cond.notify_all()
                    while value[s] is None and not errors:
                        cond.wait()
                    if errors:
                        raise RuntimeError('optimize_batch aborted')
                    v, value[s] = value[s], None
                return v
            return f

        def run(s):
            try:
                results[s] = minimize(objective(s), X0[s], **kwargs)
            except BaseException as e:          # noqa: BLE001 - reported by the coordinator
                with cond:
                    errors.append(e)
            finally:
                with cond:
                    done[s] = True
                    cond.notify_all()

        threads = [threading.Thread(target=run, args=(s,), daemon=True) for s in range(S)]
        for t in threads:
            t.start()
        P = np.tile(X0[0], (S, 1))
        self.n_batch_calls = 0
        try:
            while True:
                with cond:
                    while not errors and not all(done[s] or request[s] is not None for s in range(S)):
                        cond.wait()
                    if errors or all(done):
                        break
                    live = [s for s in range(S) if request[s] is not None]
                    for s in live:
                        P[s] = request[s]
                        request[s] = None
                it = iter(live)
                elbo, taken = self.ELBO_batch(P, max_iter=max_iter, state='previous',
                                              work_source=lambda: next(it, -1))
                self.n_batch_calls += 1
                with cond:
                    for s in live:
                        value[s] = -float(elbo[s])
                    cond.notify_all()
        except BaseException as e:              # noqa: BLE001
            with cond:
                errors.append(e)
                cond.notify_all()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        best = min(range(S), key=lambda s: results[s].fun if np.isfinite(results[s].fun) else np.inf)
        self.set_parameters(results[best].x)
        return results

    def logposterior_batch(self, thetas, priors, names=None, max_iter=100, rows=None):
        """Vectorised log-posterior for ensemble samplers (emcee ``vectorize=True``): rows of ``thetas`` are walker
        positions in the free parameters; returns (log prior + ELBO, ELBO), each (B,).  One ``ELBO_batch`` call for
        the rows with a finite prior, capped at ``max_iter`` iterations like the reference's ``logposterior``
        (:1214-1219), warm-started per row from the device-resident chain state.  ``rows``: evaluate only these
        row indices (the others come back as -inf) -- an ensemble sampler that moves half of its walkers passes the
        whole ensemble and the moving half, so that row = walker = chain on the device."""
        self._require_components()
        thetas = np.atleast_2d(np.asarray(thetas, dtype=float))
        if names is None:
            names = np.array(list(self.parameters_dict.keys()))[~self.frozen_mask]
        lp = np.full(thetas.shape[0], -np.inf)
        todo = range(thetas.shape[0]) if rows is None else [int(r) for r in rows]
        for r in todo:
            lp[r] = sum(priors[n].logpdf(v) for v, n in zip(thetas[r], names))
        ok = np.isfinite(lp)
        elbo = np.full(thetas.shape[0], -np.inf)
        if ok.any():
            live = np.flatnonzero(ok)
            it = iter(live.tolist())
            # rows outside the prior support are never handed out; give them a harmless in-support stand-in
            rows = np.where(ok[:, None], thetas, thetas[live[0]])
            e, _ = self.ELBO_batch(rows, max_iter=max_iter, state='previous', work_source=lambda: next(it, -1))
            elbo[live] = e[live]
        total = np.where(ok, lp + elbo, -np.inf)
        return total, elbo

    def mcmc(self, priors, p0=None, vars=None, niter=500, filename="gprn.h5", seed=None, **kwargs):
        """Posterior sampling of the free hyper-parameters with an affine-invariant ensemble sampler
        (reference :1154-1286: ``2 * ndim`` walkers, log-posterior = log prior + ELBO capped at 100 iterations, the ELBO
        kept as blob, autocorrelation-time convergence check every 10 steps).

        The walkers that move in a half step are ONE batched device call (``logposterior_batch``), each warm-started
        from its own device-resident variational state.  The sampler is ``gpyrn_b200.sampler.EnsembleSampler`` (the
        stretch move of emcee's default configuration, which the reference uses; emcee itself is not needed); the
        chain is written to ``filename`` -- ``gprn.h5`` in emcee's HDF5 backend layout like the reference's (:1253-1255;
        ``gpyrn_b200.h5chain`` writes the container, h5py is not needed), a ``.npz`` archive for any other extension,
        ``None``: no file.
        Returns the sampler (``get_chain``, ``get_log_prob``, ``get_blobs``, ``acceptance_fraction``,
        ``get_autocorr_time`` as in emcee).  ``kwargs`` go to the sampler (``a``)."""
        from .sampler import EnsembleSampler, backend_for
        self._require_components()
        if vars is not None and not isinstance(vars, (str, list)):
            raise ValueError(f'`vars` should be str or list, got {type(vars)}')
        self._select_vars(vars)
        names = np.array(list(self.parameters_dict.keys()))[~self.frozen_mask]
        rng = np.random.default_rng(seed)

        def prior_rvs():
            return np.array([priors[n].rvs(random_state=rng) for n in names])

        def logprior(theta):
            return sum(priors[n].logpdf(v) for v, n in zip(theta, names))

        def logposterior(coords, rows):           # the whole ensemble + the rows that move: row = walker = chain
            total, elbo = self.logposterior_batch(coords, priors, names=names, max_iter=100, rows=rows)
            bad = np.isnan(total)                 # a covariance matrix that is not positive definite: reject the move
            return np.column_stack([np.where(bad, -np.inf, total), np.where(bad, -np.inf, elbo)])

        ndim = len(names)
        nwalkers = 2 * ndim
        print(f'Setting up sampler (parameters: {ndim}, walkers: {nwalkers})')
        if p0 is None:
            p0 = np.array([prior_rvs() for _ in range(nwalkers)])
        else:                                      # a small ball around p0, as the reference draws it (:1238-1246)
            sigma = []
            for n in names:
                sd = priors[n].std
                sigma.append(float(sd() if callable(sd) else sd))
            # emcee.utils.sample_ellipsoid(p0, covmat, size): normal draws with `diag(sigma) / 100` as COVARIANCE
            p0 = rng.multivariate_normal(np.asarray(p0, dtype=float), np.diag(sigma) / 100, size=nwalkers)
            for i, pt in enumerate(p0):
                if np.isneginf(logprior(pt)):
                    p0[i] = prior_rvs()
        print('initial values for parameters are set')
        self.reset_chain_state()
        backend = backend_for(filename) if filename else None
        sampler = EnsembleSampler(nwalkers, ndim, logposterior, seed=rng.integers(2 ** 31), backend=backend,
                                  rows_aware=True, **kwargs)
        old_tau = np.inf
        for sample in sampler.sample(p0, iterations=niter):
            if sampler.iteration % 10:
                continue
            print(sample.log_prob.max())
            tau = sampler.get_autocorr_time(tol=0)          # an estimate even if it is not trustworthy yet
            converged = np.all(tau * 100 < sampler.iteration) & np.all(np.abs(old_tau - tau) / tau < 0.01)
            if converged:
                print('MCMC converged!')
                break
            old_tau = tau
        if backend is not None and sampler.iteration:
            backend.save(sampler, force=True)
        return sampler

    # ------------------------------------------------------------------------------------------
    # prediction
    # ------------------------------------------------------------------------------------------
    def _Prediction(self, nodes=None, weights=None, means=None, jitters=None, tstar=None, mu=None, var=None,
                    separate=False):
        """
        Predictive mean and variance of the GPRN at ``tstar`` (reference :1289-1379).

        Returns:
            mean (T, p), variance (T, p) and, with ``separate``, an object array [node means (q,T),
            weight means (q*p,T)].
        """
        nodes = self.nodes if nodes is None else nodes
        weights = self.weights if weights is None else weights
        means = self.means if means is None else means
        jitters = self.jitters if jitters is None else jitters
        tstar = self.time if tstar is None else np.atleast_1d(np.asarray(tstar, dtype=float))
        if mu is None and var is None:
            if self._mu is None and self._var is None:
                mu, var = self._initMuVar(nodes, weights, jitters)
            else:
                mu, var = self._mu, self._var
        self._bind_model(nodes, weights)
        hyper = _lib.f64(self._hyper_vector(nodes, weights, jitters))
        T = tstar.size
        mean_t = _lib.f64(self._mean(means, tstar))
        ts = _lib.f64(tstar)
        muf = _lib.f64(np.asarray(mu).ravel())
        varf = _lib.f64(np.asarray(var).ravel())
        pm = np.empty((T, self.p))
        pv = np.empty((T, self.p))
        npred = np.empty((self.q, T))
        wpred = np.empty((self.qp, T))
        _lib.check(_lib.lib().gprn_predict(self._h(), _lib.dptr(hyper), _lib.dptr(muf), _lib.dptr(varf),
                                           _lib.dptr(ts), T, _lib.dptr(mean_t), _lib.dptr(pm), _lib.dptr(pv),
                                           _lib.dptr(npred), _lib.dptr(wpred), None))
        if separate:
            sep = np.empty(2, dtype=object)
            sep[0], sep[1] = npred, wpred
            return pm, pv, sep
        return pm, pv

    def Prediction_batch(self, parameters, tstar=None, max_iter=None, return_info=False):
        """
        Predictive mean / variance for B hyper-parameter sets (a posterior chain): each set's
        variational state is converged on the device (``ELBO_batch``) and then fed to the predictive
        (``_Prediction``, reference :1289-1379).  The object's own parameters are left untouched.

        Args:
            parameters: array (B, n_parameters) or (B, number of free parameters), ``get_parameters`` order
            tstar: test epochs (default: the training epochs)

        Returns:
            mean (B, T, p), variance (B, T, p) [, elbo (B,), iterations (B,), status (B,)]
        """
        self._require_components()
        P = self._full_parameter_matrix(parameters)
        B = P.shape[0]
        tstar = self.time if tstar is None else np.atleast_1d(np.asarray(tstar, dtype=float))
        elbo, iters, status, mu, var = self.ELBO_batch(P, max_iter=max_iter, return_info=True, return_state=True)
        n_kernel = sum(k.pars.size for k in chain(self.nodes, self.weights))
        n_mean = sum(m.pars.size for m in self.means if isinstance(m, meanfunc.meanFunction))
        T = tstar.size
        ts = _lib.f64(tstar)
        hyper = _lib.f64(np.concatenate([P[:, :n_kernel], P[:, n_kernel + n_mean:]], axis=1))
        saved = [m.pars.copy() if isinstance(m, meanfunc.meanFunction) else None for m in self.means]
        try:
            if n_mean == 0 or np.all(P[:, n_kernel:n_kernel + n_mean] == P[0, n_kernel:n_kernel + n_mean]):
                self._assign_means(P[0, n_kernel:n_kernel + n_mean])
                mean_t, shared = _lib.f64(self._mean(self.means, tstar)), 1
            else:
                mean_t, shared = np.empty((B, self.p * T)), 0
                for b in range(B):
                    self._assign_means(P[b, n_kernel:n_kernel + n_mean])
                    mean_t[b] = self._mean(self.means, tstar)
        finally:
            self._restore_means(saved)
        pm, pv = np.empty((B, T, self.p)), np.empty((B, T, self.p))
        # one device call: assembly / factorisation / solves batched over the GPs of all sets that fit the workspace
        _lib.check(_lib.lib().gprn_predict_batched(self._h(), B, _lib.dptr(hyper), _lib.dptr(_lib.f64(mu)),
                                                   _lib.dptr(_lib.f64(var)), _lib.dptr(ts), T, _lib.dptr(mean_t), shared,
                                                   _lib.dptr(pm), _lib.dptr(pv), None, None, None))
        if return_info:
            return pm, pv, elbo, iters, status
        return pm, pv

    def predict(self, tstar=None, nn=1000):
        """GPRN prediction at ``tstar`` (default: ``nn`` points spanning the data +/- 20 %).
        Returns tstar, mean, standard deviation, separate node/weight predictions."""
        if tstar is None:
            mi, ma = self.time.min(), self.time.max()
            span = ma - mi
            tstar = np.linspace(mi - 0.2 * span, ma + 0.2 * span, nn)
        aa, vv, bb = self._Prediction(tstar=tstar, separate=True)
        return tstar, aa, np.sqrt(vv), bb
