"""Mean-field variational inference for GPRNs (Nguyen & Bonilla 2013) on a B200.

Host-side mirror of the reference's ``gpyrn.meanfield.inference`` (gpyrn/meanfield.py:92-1403): same
constructor, ``set_components``, parameter-vector bookkeeping, ``ELBO`` / ``ELBOcalc`` / ``nELBO`` /
``optimize`` / ``_Prediction`` / ``predict`` signatures and return shapes.  Everything numerical --
covariance assembly, Cholesky factorisations, the variational mu/Sigma updates, the ELBO terms, the
GP predictive -- runs in the CUDA library behind the C ABI of ``include/gprn_b200.h``; this module
only keeps the Python objects, evaluates the mean functions (O(pN), host by design) and marshals
small arrays.  There is no CPU fallback.

Additions over the reference: ``ELBO_batch(parameters[B, n])`` evaluates B hyper-parameter sets in one
device call (the data-parallel entry used for sweeps / MCMC walkers), and ``device=`` selects the GPU.
"""
import ctypes
import time as time_module
from itertools import chain

import numpy as np

from . import _lib, covfunc, meanfunc


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

    def ELBO_batch(self, parameters, max_iter=None, return_info=False, mu=None, var=None, return_state=False):
        """
        ELBO of B hyper-parameter sets in one device call.

        Args:
            parameters: array (B, n_parameters) or (B, number of free parameters), ``get_parameters`` order
            max_iter: per-evaluation iteration cap (default 10000)
            return_info: also return (iterations, status) arrays
            mu, var: optional per-set initial variational state, arrays (B, d) -- the batched form of
                ``ELBOcalc(mu='previous')`` for chains / walkers that carry their own state; default 'init'
            return_state: also return the final (mu, var), arrays (B, d), to be fed back on the next call

        Returns:
            elbo (B,) [, iterations (B,), status (B,)] [, mu (B, d), var (B, d)]
        """
        self._require_components()
        P = np.atleast_2d(np.asarray(parameters, dtype=float))
        B = P.shape[0]
        n_all = self.n_parameters
        if P.shape[1] != n_all:
            full = np.tile(self.get_parameters(include_frozen=True), (B, 1))
            full[:, ~self.frozen_mask] = P
            P = full
        n_kernel = sum(k.pars.size for k in chain(self.nodes, self.weights))
        n_mean = sum(m.pars.size for m in self.means if isinstance(m, meanfunc.meanFunction))
        self._bind_model(self.nodes, self.weights)
        hyper = _lib.f64(np.concatenate([P[:, :n_kernel], P[:, n_kernel + n_mean:]], axis=1))
        if n_mean == 0 or np.all(P[:, n_kernel:n_kernel + n_mean] == P[0, n_kernel:n_kernel + n_mean]):
            saved = [m.pars.copy() if isinstance(m, meanfunc.meanFunction) else None for m in self.means]
            self._assign_means(P[0, n_kernel:n_kernel + n_mean])
            ysub = _lib.f64(np.concatenate(self.y) - self._mean(self.means))
            self._restore_means(saved)
            shared = 1
        else:
            saved = [m.pars.copy() if isinstance(m, meanfunc.meanFunction) else None for m in self.means]
            ysub = np.empty((B, self.p * self.N))
            for b in range(B):
                self._assign_means(P[b, n_kernel:n_kernel + n_mean])
                ysub[b] = np.concatenate(self.y) - self._mean(self.means)
            self._restore_means(saved)
            ysub = _lib.f64(ysub)
            shared = 0
        elbo = np.empty(B)
        iters = np.zeros(B, dtype=np.int32)
        status = np.zeros(B, dtype=np.int32)
        init_mode, mu_io, var_io = 0, None, None
        if mu is not None or var is not None:
            if mu is None or var is None:
                raise ValueError('provide both mu and var, or neither')
            mu_io = _lib.f64(np.array(mu, dtype=float).reshape(B, self.d))
            var_io = _lib.f64(np.array(var, dtype=float).reshape(B, self.d))
            init_mode = 1
        elif return_state:
            mu_io, var_io = np.empty((B, self.d)), np.empty((B, self.d))
        _lib.check(_lib.lib().gprn_elbo_batched(self._h(), B, _lib.dptr(hyper), _lib.dptr(ysub), shared, init_mode,
                                                _lib.dptr(mu_io), _lib.dptr(var_io),
                                                -1 if max_iter is None else max_iter, _lib.dptr(elbo),
                                                _lib.iptr(iters), _lib.iptr(status), None))
        out = (elbo, iters, status) if return_info else (elbo,)
        if return_state:
            out = out + (mu_io, var_io)
        return out if len(out) > 1 else out[0]

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

    def mcmc(self, priors, p0=None, vars=None, niter=500, **kwargs):
        """Posterior sampling of the hyper-parameters with emcee (the sampler itself is a host driver
        outside this package's scope; it needs the optional ``emcee`` dependency)."""
        try:
            from emcee import EnsembleSampler
        except ImportError as e:
            raise ImportError('inference.mcmc needs the optional dependency `emcee`') from e
        self._require_components()
        self._select_vars(vars)
        names = np.array(list(self.parameters_dict.keys()))[~self.frozen_mask]

        def logprior(theta):
            return sum(priors[n].logpdf(v) for v, n in zip(theta, names))

        def logposterior(theta):
            lp = logprior(theta)
            if np.isneginf(lp):
                return -np.inf, -np.inf
            elbo = -self.nELBO(theta, max_iter=100)
            return lp + elbo, elbo

        ndim = len(names)
        nwalkers = 2 * ndim
        if p0 is None:
            p0 = np.array([[priors[n].rvs() for n in names] for _ in range(nwalkers)])
        sampler = EnsembleSampler(nwalkers, ndim, logposterior, **kwargs)
        sampler.run_mcmc(p0, niter, progress=False)
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
        P = np.atleast_2d(np.asarray(parameters, dtype=float))
        B = P.shape[0]
        if P.shape[1] != self.n_parameters:
            full = np.tile(self.get_parameters(include_frozen=True), (B, 1))
            full[:, ~self.frozen_mask] = P
            P = full
        tstar = self.time if tstar is None else np.atleast_1d(np.asarray(tstar, dtype=float))
        elbo, iters, status, mu, var = self.ELBO_batch(P, max_iter=max_iter, return_info=True, return_state=True)
        n_kernel = sum(k.pars.size for k in chain(self.nodes, self.weights))
        n_mean = sum(m.pars.size for m in self.means if isinstance(m, meanfunc.meanFunction))
        T = tstar.size
        ts = _lib.f64(tstar)
        pm, pv = np.empty((B, T, self.p)), np.empty((B, T, self.p))
        saved = [m.pars.copy() if isinstance(m, meanfunc.meanFunction) else None for m in self.means]
        try:
            for b in range(B):
                self._assign_means(P[b, n_kernel:n_kernel + n_mean])
                mean_t = _lib.f64(self._mean(self.means, tstar))
                hyper = _lib.f64(np.concatenate([P[b, :n_kernel], P[b, n_kernel + n_mean:]]))
                _lib.check(_lib.lib().gprn_predict(self._h(), _lib.dptr(hyper), _lib.dptr(_lib.f64(mu[b])),
                                                   _lib.dptr(_lib.f64(var[b])), _lib.dptr(ts), T, _lib.dptr(mean_t),
                                                   _lib.dptr(pm[b]), _lib.dptr(pv[b]), None, None, None))
        finally:
            self._restore_means(saved)
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
