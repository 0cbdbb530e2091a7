"""Drop-in for the reference's exact-GP "delta map" plugin
(policy_transportation/models/gaussian_process.py:16-126) with every posterior computation on the B200 engine.

Same constructor, methods, return layouts and post-fit attributes; the L-BFGS-B driver and sklearn's restart logic
(sklearn:_gpr.py:299-344, 658-674) stay on the host exactly as in the reference, with the objective
(log marginal likelihood + gradient) served by `gptb_lml`.
"""
from __future__ import annotations

import warnings
from operator import itemgetter

import numpy as np
import scipy.optimize
from sklearn.base import clone
from sklearn.utils import check_random_state

from . import _lib
from .kernel_spec import check_supported, kernel_kind, map_gradient, read_params

GPR_CHOLESKY_LOWER = True


class _Regressor:
    """What callers reach through `GaussianProcess.gp` in the reference (an sklearn GaussianProcessRegressor):
    `alpha`, `kernel_`, `L_`, `alpha_`, `log_marginal_likelihood_value_`, `log_marginal_likelihood()`."""

    def __init__(self, owner, kernel, alpha, optimizer, n_restarts_optimizer, n_targets):
        self._o = owner
        self.kernel = kernel
        self.alpha = alpha
        self.optimizer = optimizer
        self.n_restarts_optimizer = n_restarts_optimizer
        self.n_targets = n_targets
        self.kernel_ = None
        self.log_marginal_likelihood_value_ = None

    @property
    def L_(self):
        o = self._o
        if getattr(o, "_L_cache", None) is not None:
            return o._L_cache
        if o._engine.spatial:
            # the handle holds the factor of the Morton-ordered system; sklearn's L_ is the factor in the caller's order:
            # factorise once more on a scratch handle in natural order (attribute access only, not on any hot path)
            scratch = _lib.Engine(o._engine.device)
            try:
                scratch.set_kernel_kind(kernel_kind(self.kernel_))
                scratch.set_train(o.X, o.Y)
                c, ell, s2 = read_params(self.kernel_, scratch.d)
                info, _ = scratch.factorize(c, ell, s2, self.alpha, want_lml=False)
                if info > 0:
                    raise np.linalg.LinAlgError("kernel matrix is not positive definite")
                o._L_cache = scratch.export_L()      # kept until the next fit / append: repeated attribute reads cost nothing
                return o._L_cache
            finally:
                scratch.close()
        o._ensure_fitted_factor()
        return o._engine.export_L()

    @property
    def alpha_(self):
        self._o._ensure_fitted_factor()     # an LML evaluation at another theta may have overwritten the handle's factor
        return self._o._engine.export_alpha()

    @property
    def X_train_(self):
        return self._o.X

    @property
    def y_train_(self):
        return self._o.Y

    def log_marginal_likelihood(self, theta=None, eval_gradient=False, clone_kernel=True):
        """sklearn:_gpr.py:541-656 -- evaluated on the GPU for `theta` (log-transformed, active hyper-parameters)."""
        o = self._o
        if theta is None:
            if eval_gradient:
                raise ValueError("Gradient can only be evaluated for theta!=None")
            return self.log_marginal_likelihood_value_
        kernel = self.kernel_.clone_with_theta(theta) if clone_kernel else self.kernel_
        if not clone_kernel:
            kernel.theta = theta
        c, ell, s2 = read_params(kernel, o._engine.d)
        info, lml, g = o._engine.lml(c, ell, s2, self.alpha, want_grad=eval_gradient)
        o._factor_theta = None          # the handle now holds the factorisation for `theta`, not the fitted one
        if info > 0:                    # not positive definite: sklearn:_gpr.py:590-593
            return (-np.inf, np.zeros_like(theta)) if eval_gradient else -np.inf
        if eval_gradient:
            return lml, map_gradient(kernel, g, o._engine.d)
        return lml


class GaussianProcess:
    MAX_CONCURRENT_FITS = 12                             # engine handles used at once by parallel_restarts (see _optimise_concurrently)

    def __init__(self, kernel, alpha=1e-10, optimizer='fmin_l_bfgs_b', n_restarts_optimizer=5, n_targets=None, device=None,
                 variance_mode=None, spatial=None, parallel_restarts=None):
        check_supported(kernel)
        if optimizer is None:
            n_restarts_optimizer = 0                     # gaussian_process.py:18-21 (sklearn default)
        self.gp = _Regressor(self, kernel, alpha, optimizer, n_restarts_optimizer, n_targets)
        self.kernel = kernel
        self.alpha = alpha
        self._device = device
        # how |L^-1 k*|^2 is evaluated: "fp64" (DMMA, default), "int8x5|6|7" (INT8-sliced tcgen05 path, 7-bit digit planes) or
        # "int8w4|5|6" (same path, 8-bit digit planes: fewer plane products for the same accuracy; include/gptb200.h);
        # the environment variable GPTB_VARIANCE_MODE sets the default for objects that do not say
        import os as _os
        self._variance_mode = variance_mode or _os.environ.get("GPTB_VARIANCE_MODE", "fp64")
        # spatial mode (include/gptb200.h: Morton-ordered training points, sorted query batches, zero-plane skipping); GPTB_SPATIAL=1
        self._spatial = bool(int(_os.environ.get("GPTB_SPATIAL", "0"))) if spatial is None else bool(spatial)
        # small-N fits are launch-latency bound (tens of microsecond kernels, a host round trip per objective evaluation): the optimiser
        # restarts are independent L-BFGS-B runs, so they can run concurrently, one engine handle (= one stream set) per run, the
        # GPU interleaving their kernels.  Same start points (drawn up front from the global RNG in the same order), same runs, same
        # arg-min as the sequential loop of sklearn:_gpr.py:312-337.  Opt-in (GPTB_PARALLEL_RESTARTS=1 sets the default).
        self._parallel_restarts = bool(int(_os.environ.get("GPTB_PARALLEL_RESTARTS", "0"))) if parallel_restarts is None else bool(parallel_restarts)
        self._restart_engines = []
        self._engine_obj = None
        self._factor_theta = None
        self._K_inv = None

    # -- engine handle (created lazily so that constructing the object never touches CUDA, like the reference) ----
    @property
    def _engine(self):
        if self._engine_obj is None:
            dev = self._device
            if dev is None:
                import os
                dev = int(os.environ.get("LOCAL_RANK", "0")) if os.environ.get("GPTB_DEVICE") is None else int(os.environ["GPTB_DEVICE"])
            self._engine_obj = _lib.Engine(dev)
            vm = str(self._variance_mode).lower()
            if vm != "fp64":
                self._engine_obj.set_variance_mode(*_lib.parse_variance_mode(vm))
            if self._spatial:
                self._engine_obj.set_spatial(True)
        return self._engine_obj

    # -- fit -------------------------------------------------------------------------------------------------------
    def fit(self, X, Y):
        X = np.asarray(X, dtype=np.float64)
        Y = np.asarray(Y, dtype=np.float64)
        if Y.ndim == 1:
            Y = Y[:, None]
        self.X = X
        self.Y = Y
        self.n_features = np.shape(self.X)[1]
        self.n_samples = np.shape(self.X)[0]
        self.n_outputs = np.shape(self.Y)[1]
        mask = np.isnan(self.Y).any(axis=1)              # gaussian_process.py:33-35 (quirk Q9)
        self.X = self.X[~mask]
        self.Y = self.Y[~mask]

        gp = self.gp
        if gp.n_targets is not None and self.Y.shape[1] != gp.n_targets:     # sklearn:_gpr.py:268-273
            raise ValueError(
                "The number of targets seen in `y` is different from the parameter `n_targets`. "
                f"Got {self.Y.shape[1]} != {gp.n_targets}.")
        eng = self._engine
        eng.set_kernel_kind(kernel_kind(gp.kernel))
        eng.set_train(self.X, self.Y)
        gp.kernel_ = clone(gp.kernel)
        rng = check_random_state(None)                   # the global numpy RNG, as in the reference (quirk Q10)
        kernel_ = gp.kernel_
        read_params(kernel_, eng.d)                      # validates the length-scale dimensionality early

        if gp.optimizer is not None and kernel_.n_dims > 0:
            def obj_func(theta, eval_gradient=True):
                if eval_gradient:
                    lml, grad = gp.log_marginal_likelihood(theta, eval_gradient=True, clone_kernel=False)
                    return -lml, -grad
                return -gp.log_marginal_likelihood(theta, clone_kernel=False)

            if gp.n_restarts_optimizer > 0 and not np.isfinite(kernel_.bounds).all():
                raise ValueError("Multiple optimizer restarts (n_restarts_optimizer>0) requires that all bounds are finite.")
            if self._parallel_restarts and gp.n_restarts_optimizer > 0:
                optima = self._optimise_concurrently(kernel_, rng)
            else:
                optima = [self._constrained_optimization(obj_func, kernel_.theta, kernel_.bounds)]
                bounds = kernel_.bounds
                for _ in range(gp.n_restarts_optimizer):
                    theta_initial = rng.uniform(bounds[:, 0], bounds[:, 1])
                    optima.append(self._constrained_optimization(obj_func, theta_initial, bounds))
            lml_values = list(map(itemgetter(1), optima))
            kernel_.theta = optima[int(np.argmin(lml_values))][0]
            kernel_._check_bounds_params()
            gp.log_marginal_likelihood_value_ = -np.min(lml_values)
            want_lml = False
        else:
            kernel_.theta = kernel_.theta                # normalises params to numpy values (quirk Q12)
            want_lml = True

        c, ell, s2 = read_params(kernel_, eng.d)
        info, lml = eng.factorize(c, ell, s2, gp.alpha, want_lml=want_lml)   # sklearn:_gpr.py:347-367
        if info > 0:
            raise np.linalg.LinAlgError(
                f"The kernel, {kernel_}, is not returning a positive definite matrix "
                f"({info}-th leading minor). Try gradually increasing the 'alpha' parameter of your "
                "GaussianProcessRegressor estimator.")
        if want_lml:
            gp.log_marginal_likelihood_value_ = lml
        self._factor_theta = np.array(kernel_.theta, copy=True)
        self._K_inv = None
        self._L_cache = None

        self.kernel = kernel_
        prm = self.kernel.get_params()
        self.kernel_params_ = [prm['k1__k2__length_scale'], prm['k1']]
        self.noise_var_ = gp.alpha + prm['k2__noise_level']
        self.prior_var = prm['k1__k1__constant_value']
        print('lenghtscales', prm['k1__k2__length_scale'])

    def _optimise_concurrently(self, kernel_, rng):
        """Run 0 (from the kernel's theta) and the restarts as concurrent L-BFGS-B runs, one scratch engine each."""
        from concurrent.futures import ThreadPoolExecutor
        gp, eng = self.gp, self._engine
        bounds = kernel_.bounds
        starts = [np.array(kernel_.theta, copy=True)] + [rng.uniform(bounds[:, 0], bounds[:, 1]) for _ in range(gp.n_restarts_optimizer)]
        # at most MAX_CONCURRENT_FITS handles factorise at once: each factorisation keeps a spine kernel of eight CTAs that meet at a
        # barrier on the GPU, and the handles of one device must never hold all of its SMs with half-started ones
        n_eng = min(len(starts), self.MAX_CONCURRENT_FITS)
        while len(self._restart_engines) < n_eng - 1:
            self._restart_engines.append(_lib.Engine(eng.device))
        engines = [eng] + self._restart_engines[:n_eng - 1]
        for e in engines[1:]:
            e.set_kernel_kind(kernel_kind(kernel_))
            e.set_train(self.X, self.Y)
        import queue
        free = queue.SimpleQueue()
        for e in engines:
            free.put(e)

        def run(i):
            e = free.get()
            try:
                def obj(theta, eval_gradient=True):
                    k = kernel_.clone_with_theta(theta)              # per-evaluation clone: the runs share no mutable kernel object
                    c, ell, s2 = read_params(k, e.d)
                    info, lml, g = e.lml(c, ell, s2, gp.alpha, want_grad=eval_gradient)
                    if info > 0:
                        return (np.inf, np.zeros_like(theta)) if eval_gradient else np.inf
                    return (-lml, -map_gradient(k, g, e.d)) if eval_gradient else -lml

                return self._constrained_optimization(obj, starts[i], bounds)
            finally:
                free.put(e)

        with ThreadPoolExecutor(max_workers=n_eng) as pool:
            optima = list(pool.map(run, range(len(starts))))
        self._factor_theta = None
        return optima

    def _constrained_optimization(self, obj_func, initial_theta, bounds):
        gp = self.gp
        if gp.optimizer == "fmin_l_bfgs_b":              # sklearn:_gpr.py:658-668
            res = scipy.optimize.minimize(obj_func, initial_theta, method="L-BFGS-B", jac=True, bounds=bounds)
            if res.status != 0:
                from sklearn.exceptions import ConvergenceWarning
                warnings.warn(f"lbfgs failed to converge (status={res.status}): {res.message}", ConvergenceWarning)
            return res.x, res.fun
        if callable(gp.optimizer):
            theta_opt, func_min = gp.optimizer(obj_func, initial_theta, bounds=bounds)
            return theta_opt, func_min
        raise ValueError(f"Unknown optimizer {gp.optimizer}.")

    def _ensure_fitted_factor(self):
        """LML evaluations through `gp.log_marginal_likelihood` overwrite the handle's factor; restore the fitted one."""
        if not hasattr(self, "X"):
            raise RuntimeError("GaussianProcess is not fitted")
        if self._factor_theta is None:
            c, ell, s2 = read_params(self.kernel, self._engine.d)
            info, _ = self._engine.factorize(c, ell, s2, self.gp.alpha, want_lml=False)
            if info > 0:
                raise np.linalg.LinAlgError("kernel matrix is not positive definite")
            self._factor_theta = np.array(self.kernel.theta, copy=True)

    @property
    def K_inv(self):
        """(c R + (alpha + s2) I)^-1, gaussian_process.py:42-43 -- built on demand from the Cholesky factor."""
        if self._K_inv is None:
            self._ensure_fitted_factor()
            self._K_inv = self._engine.export_Kinv()
        return self._K_inv

    # -- queries ---------------------------------------------------------------------------------------------------
    def _query(self, x, flags, vel=None):
        self._ensure_fitted_factor()
        x = np.asarray(x, dtype=np.float64)
        if x.ndim == 1:
            raise ValueError("Expected 2D array, got 1D array instead")
        return self._engine.query(x, flags, vel)

    def predict(self, x, return_std=False, return_cov=False):
        squeeze = self.n_outputs == 1                                     # sklearn squeezes single-target outputs (_gpr.py:452-456,495-499)
        if return_std == True:                                            # noqa: E712  (reference semantics)
            o = self._query(x, _lib.MEAN | _lib.STD)
            if squeeze:
                return o["mean"][:, 0], o["std"][:, 0]
            return o["mean"], o["std"]
        if return_cov == True:                                            # noqa: E712
            mean, cov = self._joint(x)
            if squeeze:
                return mean[:, 0], cov
            # sklearn:_gpr.py:470-478 -- the covariance is replicated over the outputs
            return mean, np.repeat(cov[:, :, np.newaxis], self.n_outputs, axis=2)
        mean = self._query(x, _lib.MEAN)["mean"]
        return mean[:, 0] if squeeze else mean

    def _joint(self, x):
        self._ensure_fitted_factor()
        x = np.asarray(x, dtype=np.float64)
        return self._engine.query_cov(x)

    def samples(self, x):
        """10 joint posterior samples, seed 0, layout (10, M, n_outputs) (gaussian_process.py:57-60 over sklearn's sample_y,
        _gpr.py:502-539).  Mean and covariance come from the GPU; the draw itself is numpy's RandomState(0) multivariate
        normal exactly as sklearn calls it, so the samples match the reference's stream."""
        mean, cov = self._joint(x)
        rng = check_random_state(0)
        n_samples = 10
        if self.n_outputs == 1:
            y = rng.multivariate_normal(mean[:, 0], cov, n_samples).T                     # (M, 10), as sklearn returns it
            return np.transpose(y, (2, 0, 1))     # the reference transposes three axes here and fails for one output; so do we
        cols = [rng.multivariate_normal(mean[:, t], cov, n_samples).T[:, np.newaxis] for t in range(self.n_outputs)]
        y = np.hstack(cols)                                                                # (M, p, 10)
        return np.transpose(y, (2, 0, 1))

    def derivative(self, x, return_var=False):
        """Jacobian of the posterior mean, layout (M, n_outputs, n_features) (quirk Q4), and optionally the variance of
        each Jacobian entry (identical across outputs, quirk Q5)."""
        if return_var == True:                                            # noqa: E712
            o = self._query(x, _lib.JAC | _lib.JACVAR)
            return o["jac"], o["jacvar"]
        return self._query(x, _lib.JAC)["jac"]

    def derivative_of_variance(self, x):
        return self._query(x, _lib.DVAR)["dvar"]

    # -- rank-1 growth of the training set at the fitted hyper-parameters ---------------------------------------------------------
    def append(self, x, y):
        """Add ONE training pair (x (d,), y (p,)) to the fitted model without re-optimising the hyper-parameters: what a re-`fit` with
        `optimizer=None` (or an all-fixed kernel) on the extended set computes, as an O(N^2) update of L, L^-1 and alpha on the device
        (include/gptb200.h gptb_append_point) instead of an O(N^3) re-factorisation.  The greedy active-learning loop of
        models/gaussian_process_al.py:41-55 is the caller."""
        self._ensure_fitted_factor()
        x = np.asarray(x, dtype=np.float64).reshape(1, -1)
        y = np.asarray(y, dtype=np.float64).reshape(1, -1)
        info, lml = self._engine.append_point(x, y, want_lml=True)
        if info > 0:
            raise np.linalg.LinAlgError(f"The kernel, {self.kernel}, is not returning a positive definite matrix ({info}-th leading minor).")
        self.X = np.vstack([self.X, x])
        self.Y = np.vstack([self.Y, y])
        self.n_samples = self.X.shape[0]
        self.gp.log_marginal_likelihood_value_ = lml
        self._K_inv = None
        self._L_cache = None

    # -- minimum-variance stabilisation (plot_utils.py:283-317): the two query shapes of the reference's plotting helpers -------------
    def minimum_variance_field(self, x, gain=2.0):
        """vel - gain * std * g / |g| with g = derivative_of_variance (plot_utils.plot_vector_field_minvar:286-289, gain 2): one fused
        query (mean, std and the variance gradient share the regenerated k(x*, X)); returns (M, n_outputs)."""
        o = self._query(x, _lib.MEAN | _lib.STD | _lib.DVAR)
        grad = o["dvar"].T
        return o["mean"] - gain * o["std"] * grad / np.linalg.norm(grad, axis=1).reshape(-1, 1)

    def rollout_min_variance(self, start, steps=1000, gain=1.0):
        """Rollouts pos <- pos + mean(pos) - gain * std(pos) * g / |g| from every row of `start` (plot_utils.plot_traj_evolution:298-310 is
        the one-start-point, gain = 1 case); the loop runs on the device.  Returns the positions after each step, (steps, K, d)."""
        self._ensure_fitted_factor()
        return self._engine.rollout_min_variance(np.atleast_2d(np.asarray(start, dtype=np.float64)), steps, gain)
