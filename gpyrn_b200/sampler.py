"""Affine-invariant ensemble sampler for ``inference.mcmc`` (the caller of the batched log-posterior path).

The reference drives ``emcee.EnsembleSampler`` with one ``nELBO`` per walker and step and stores the chain with emcee's
HDF5 backend (gpyrn/meanfield.py:1154-1286).  Neither ``emcee`` nor ``h5py`` exists in this image, and the point of the
B200 path is that the walkers of a step are ONE batched device call, so this module carries the few lines of emcee
that ``mcmc`` uses, vectorised from the start:

* ``EnsembleSampler`` -- Goodman & Weare (2010) stretch move (scale a = 2) on two half-ensembles that are updated in
  turn (emcee's default ``StretchMove`` / red-blue split): every half step proposes for half of the walkers and
  evaluates all proposals with one call of the log-probability function.  ``sample`` is a generator like emcee's,
  ``get_chain`` / ``get_log_prob`` / ``get_blobs`` / ``acceptance_fraction`` / ``get_autocorr_time`` follow its meaning.
* ``integrated_time`` -- the integrated autocorrelation time with Sokal's automatic window (c = 5), the estimator
  behind ``sampler.get_autocorr_time(tol=0)`` that the reference's convergence check polls (:1268-1283).
* ``HDFBackend`` -- chain, log-probabilities and blobs in emcee's HDF5 layout (``gprn.h5`` of the reference, written by
  ``gpyrn_b200.h5chain``: h5py is not needed); ``NpzBackend`` -- the same arrays in a ``.npz`` file.

The log-probability function may be *row aware*: it then receives the coordinates of the whole ensemble plus the
indices of the rows to evaluate, and returns full-length arrays.  ``inference.mcmc`` uses that so that the
device-resident variational state (one warm start per chain, ``ELBO_batch(state='previous')``) stays keyed by walker.
Pure numpy; no device code here.
"""
import numpy as np


class State:
    """Ensemble position: ``coords`` (nwalkers, ndim), ``log_prob`` (nwalkers,), ``blobs`` (nwalkers, nblobs) or None."""

    def __init__(self, coords, log_prob=None, blobs=None):
        self.coords = np.array(coords, dtype=float)
        self.log_prob = None if log_prob is None else np.array(log_prob, dtype=float)
        self.blobs = None if blobs is None else np.array(blobs, dtype=float)


def _autocorr_1d(x):
    """Normalised autocorrelation function of a 1-d series (FFT, zero padded to the next power of two)."""
    x = np.asarray(x, dtype=float)
    n = 1
    while n < x.size:
        n <<= 1
    f = np.fft.fft(x - np.mean(x), n=2 * n)
    acf = np.fft.ifft(f * np.conjugate(f))[: x.size].real
    if acf[0] == 0.0:
        return np.ones_like(acf)
    return acf / acf[0]


def integrated_time(x, c=5, tol=50, quiet=True):
    """Integrated autocorrelation time per dimension of a chain ``x`` (nsteps, nwalkers, ndim): the autocorrelation
    functions of the walkers are averaged, tau(M) = 2 sum_{m<=M} rho(m) - 1, and M is the first lag with
    M >= c tau(M) (Sokal).  ``tol``: the estimate is trusted when the chain is longer than ``tol`` tau; with
    ``quiet`` (or ``tol = 0``) an untrusted estimate is returned anyway, otherwise ``ValueError``."""
    x = np.atleast_1d(np.asarray(x, dtype=float))
    if x.ndim == 1:
        x = x[:, None, None]
    elif x.ndim == 2:
        x = x[:, :, None]
    nsteps, nwalkers, ndim = x.shape
    tau = np.empty(ndim)
    for d in range(ndim):
        f = np.zeros(nsteps)
        for k in range(nwalkers):
            f += _autocorr_1d(x[:, k, d])
        f /= nwalkers
        taus = 2.0 * np.cumsum(f) - 1.0
        m = np.arange(taus.size) < c * taus
        window = int(np.argmin(m)) if np.any(~m) else taus.size - 1
        tau[d] = taus[window]
    if tol > 0 and not quiet and np.any(tol * tau > nsteps):
        raise ValueError(f"the chain is shorter than {tol} times the integrated autocorrelation time: {tau}")
    return tau


class NpzBackend:
    """Chain storage in one ``.npz`` file: ``chain`` (nsteps, nwalkers, ndim), ``log_prob`` (nsteps, nwalkers),
    ``blobs`` (nsteps, nwalkers, nblobs), ``accepted`` (nwalkers,), ``iteration``.  The numpy-native alternative to
    :class:`HDFBackend`."""

    def __init__(self, filename="gprn.npz", every=50):
        self.filename = filename
        self.every = max(1, int(every))
        self.nwalkers = self.ndim = 0

    def reset(self, nwalkers, ndim):
        self.nwalkers, self.ndim = int(nwalkers), int(ndim)

    def save(self, sampler, force=False):
        if not force and sampler.iteration % self.every:
            return
        blobs = sampler.get_blobs()
        np.savez_compressed(self.filename, chain=sampler.get_chain(), log_prob=sampler.get_log_prob(),
                            blobs=np.zeros((sampler.iteration, self.nwalkers, 0)) if blobs is None else blobs,
                            accepted=sampler.naccepted, iteration=sampler.iteration)

    @staticmethod
    def load(filename):
        z = np.load(filename)
        return {k: z[k] for k in z.files}


class HDFBackend(NpzBackend):
    """The reference's chain file: emcee's ``HDFBackend`` layout in an HDF5 container (``gprn.h5``,
    meanfield.py:1253-1255) -- group ``name`` with the attributes ``version`` / ``nwalkers`` / ``ndim`` / ``has_blobs`` /
    ``iteration`` and the datasets ``accepted``, ``chain``, ``log_prob``, ``blobs``.  Written by
    ``gpyrn_b200.h5chain`` (no h5py in this image): contiguous datasets, the file is rewritten at every save.  Like the
    reference's ``be.reset(nwalkers, ndim)``, ``reset`` empties the file."""

    def __init__(self, filename="gprn.h5", name="mcmc", every=50):
        super().__init__(filename, every)
        self.name = name

    def reset(self, nwalkers, ndim):
        from .h5chain import write_chain
        super().reset(nwalkers, ndim)
        write_chain(self.filename, np.zeros((0, self.nwalkers, self.ndim)), np.zeros((0, self.nwalkers)),
                    np.zeros(self.nwalkers), name=self.name)

    def save(self, sampler, force=False):
        from .h5chain import write_chain
        if not force and sampler.iteration % self.every:
            return
        write_chain(self.filename, sampler.get_chain(), sampler.get_log_prob(), sampler.naccepted, sampler.get_blobs(),
                    name=self.name)

    @staticmethod
    def load(filename, name="mcmc"):
        from .h5chain import read_chain
        return read_chain(filename, name)


def backend_for(filename, **kwargs):
    """``.h5`` / ``.hdf5`` -> :class:`HDFBackend` (the reference's format), anything else -> :class:`NpzBackend`."""
    if str(filename).lower().endswith(('.h5', '.hdf5', '.hdf')):
        return HDFBackend(filename, **kwargs)
    return NpzBackend(filename, **kwargs)


class EnsembleSampler:
    """Stretch-move ensemble sampler.

    Args:
        nwalkers, ndim: ensemble size (even, at least 2 ndim as in emcee) and dimension.
        log_prob_fn: vectorised log-probability.  Plain: ``f(coords (n, ndim)) -> (n,)`` or ``(n, 1 + nblobs)`` with
            the log-probability in column 0 (emcee ``vectorize=True``).  With ``rows_aware``:
            ``f(coords (nwalkers, ndim), rows) -> arrays of length nwalkers`` of which only ``rows`` are read.
        a: stretch scale.  seed: seed of the ``numpy.random.Generator``.  backend: object with ``reset`` / ``save``.
    """

    def __init__(self, nwalkers, ndim, log_prob_fn, a=2.0, seed=None, backend=None, rows_aware=False):
        if nwalkers % 2 or nwalkers < 2 * ndim:
            raise ValueError("the number of walkers must be even and at least twice the dimension")
        self.nwalkers, self.ndim, self.a = int(nwalkers), int(ndim), float(a)
        self.log_prob_fn, self.rows_aware = log_prob_fn, rows_aware
        self.rng = np.random.default_rng(seed)
        self.backend = backend
        self.reset()

    def reset(self):
        self.iteration = 0
        self.naccepted = np.zeros(self.nwalkers)
        self._chain, self._log_prob, self._blobs = [], [], []
        self.nevals = 0                      # log-probability calls (each one batched device call in inference.mcmc)
        if self.backend is not None:
            self.backend.reset(self.nwalkers, self.ndim)

    # ---- evaluation ------------------------------------------------------------------------
    def _evaluate(self, coords, rows):
        """log-probability (and blobs) of ``coords[rows]``: (len(rows),), (len(rows), nblobs) or None."""
        self.nevals += 1
        out = self.log_prob_fn(coords, rows) if self.rows_aware else self.log_prob_fn(coords[rows])
        out = np.asarray(out, dtype=float)
        if self.rows_aware:
            out = out[rows]
        if out.ndim == 1:
            lp, blobs = out, None
        else:
            lp, blobs = out[:, 0], (out[:, 1:] if out.shape[1] > 1 else None)
        if np.any(np.isnan(lp)):
            raise ValueError("the log-probability function returned NaN")
        return lp, blobs

    # ---- sampling --------------------------------------------------------------------------
    def sample(self, initial_state, iterations=1, store=True):
        """Generator over ``iterations`` steps from ``initial_state`` (coords array or ``State``); yields the ``State``
        after each step.  A step updates the two halves of a freshly shuffled split one after the other."""
        state = initial_state if isinstance(initial_state, State) else State(initial_state)
        if state.coords.shape != (self.nwalkers, self.ndim):
            raise ValueError(f"initial coordinates must have shape {(self.nwalkers, self.ndim)}")
        if state.log_prob is None:
            rows = np.arange(self.nwalkers)
            state.log_prob, state.blobs = self._evaluate(state.coords, rows)
        if not np.all(np.isfinite(state.log_prob)):
            raise ValueError("the initial state has walkers with a non-finite log-probability")
        half = self.nwalkers // 2
        for _ in range(int(iterations)):
            perm = self.rng.permutation(self.nwalkers)
            for S, C in ((perm[:half], perm[half:]), (perm[half:], perm[:half])):
                S = np.sort(S)
                zz = ((self.a - 1.0) * self.rng.random(S.size) + 1.0) ** 2 / self.a
                partner = state.coords[self.rng.choice(C, size=S.size)]
                prop = state.coords.copy()
                prop[S] = partner - (partner - state.coords[S]) * zz[:, None]
                lp_new, blobs_new = self._evaluate(prop, S)
                lnpdiff = (self.ndim - 1.0) * np.log(zz) + lp_new - state.log_prob[S]
                accept = np.log(self.rng.random(S.size)) < lnpdiff
                acc = S[accept]
                state.coords[acc] = prop[acc]
                state.log_prob[acc] = lp_new[accept]
                if blobs_new is not None:
                    state.blobs[acc] = blobs_new[accept]
                self.naccepted[acc] += 1
            self.iteration += 1
            if store:
                self._chain.append(state.coords.copy())
                self._log_prob.append(state.log_prob.copy())
                if state.blobs is not None:
                    self._blobs.append(state.blobs.copy())
                if self.backend is not None:
                    self.backend.save(self)
            yield state

    def run_mcmc(self, initial_state, nsteps, **kwargs):
        state = None
        for state in self.sample(initial_state, iterations=nsteps, **kwargs):
            pass
        if self.backend is not None and nsteps:
            self.backend.save(self, force=True)
        return state

    # ---- results ---------------------------------------------------------------------------
    @staticmethod
    def _view(a, flat, discard, thin):
        a = np.asarray(a)[discard::thin]
        return a.reshape((-1,) + a.shape[2:]) if flat else a

    def get_chain(self, flat=False, discard=0, thin=1):
        c = np.array(self._chain).reshape(len(self._chain), self.nwalkers, self.ndim)
        return self._view(c, flat, discard, thin)

    def get_log_prob(self, flat=False, discard=0, thin=1):
        return self._view(np.array(self._log_prob).reshape(len(self._log_prob), self.nwalkers), flat, discard, thin)

    def get_blobs(self, flat=False, discard=0, thin=1):
        if not self._blobs:
            return None
        return self._view(np.array(self._blobs), flat, discard, thin)

    @property
    def acceptance_fraction(self):
        return self.naccepted / max(1, self.iteration)

    def get_autocorr_time(self, discard=0, thin=1, **kwargs):
        return thin * integrated_time(self.get_chain(discard=discard, thin=thin), **kwargs)
