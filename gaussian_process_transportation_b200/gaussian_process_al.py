"""Drop-in for policy_transportation/models/gaussian_process_al.py:15-107: the exact GP with a greedy active-learning
subset ("GP-AL").  When the training set is larger than `n_samples_max`, a random 10 % seed set is grown one point at a
time by the largest predictive std over the remaining pool (file:26-55), re-fitting the GP after every addition exactly as
the reference does (sklearn re-optimises the hyper-parameters from the initial kernel at every `fit`).  Every fit and every
pool-wide std evaluation runs on the B200 engine; the host keeps the index bookkeeping and numpy's global RNG draw.
Note the return layouts of THIS class differ from models/gaussian_process.py: `predict` always returns (mean, std) and
`derivative` returns (dy_dx (M, d, p), dsigma_dx (M, d, 1)) (file:71-107).
"""
import numpy as np

from . import _lib
from .gaussian_process import GaussianProcess as _ExactGP


class GaussianProcess():
    def __init__(self, kernel, alpha=1e-10, n_restarts_optimizer=5, n_samples_max=20000, device=None, variance_mode=None):
        self._mk = lambda restarts: _ExactGP(kernel=kernel, alpha=alpha, n_restarts_optimizer=restarts, device=device,
                                             variance_mode=variance_mode)
        self._model = self._mk(n_restarts_optimizer)
        self.gp = self._model.gp
        self.kernel = kernel
        self.alpha = alpha
        self.n_samples_max = n_samples_max

    def fit(self, X, Y):
        import contextlib, io
        self.X = np.asarray(X, dtype=np.float64)
        self.Y = np.asarray(Y, dtype=np.float64)
        self.n_features = np.shape(self.X)[1]
        self.n_samples = np.shape(self.X)[0]
        if self.n_samples > self.n_samples_max:
            print("Starting Active Learning")
            n_initial = int(0.1 * self.n_samples_max)
            X_tmp, Y_tmp = np.copy(self.X), np.copy(self.Y)
            initial_idx = np.random.choice(range(self.n_samples), size=n_initial, replace=False)     # global RNG, file:33
            X_sample, Y_sample = X_tmp[initial_idx], Y_tmp[initial_idx]
            X_tmp = np.delete(X_tmp, initial_idx, axis=0)
            Y_tmp = np.delete(Y_tmp, initial_idx, axis=0)
            gp_active = self._mk(0)                               # GaussianProcessRegressor(kernel, alpha): optimiser on, no restarts
            quiet = io.StringIO()
            with contextlib.redirect_stdout(quiet):
                gp_active.fit(X_sample, Y_sample)
            self.n_samples_batch = int(self.n_samples_max / 20)
            for _ in range(int(self.n_samples_max - n_initial)):
                # the constant sqrt(noise) shift of our predict() does not move the argmax (file:41-46)
                _, std = gp_active.predict(X_tmp, return_std=True)
                std = np.reshape(std, (-1, Y_tmp.shape[1]))
                query_idx = np.argmax(std[:, 0])
                X_sample = np.vstack([X_sample, X_tmp[query_idx]])
                Y_sample = np.vstack([Y_sample, Y_tmp[query_idx]])
                X_tmp = np.delete(X_tmp, query_idx, axis=0)
                Y_tmp = np.delete(Y_tmp, query_idx, axis=0)
                if gp_active.gp.optimizer is None or gp_active.gp.kernel_.n_dims == 0:
                    # nothing to optimise (all hyper-parameters fixed): the re-fit is a rank-1 update of the factor on the device
                    gp_active.append(X_sample[-1], Y_sample[-1])
                else:
                    with contextlib.redirect_stdout(quiet):
                        gp_active.fit(X_sample, Y_sample)
            mean = gp_active.predict(self.X)
            error = np.mean(np.sum(np.abs(mean - self.X), axis=1))        # sic: compared against the INPUTS (file:57)
            print("error:", error)
            self.n_samples = self.n_samples_max
            self.X, self.Y = np.copy(X_sample), np.copy(Y_sample)
        self._model.fit(self.X, self.Y)
        self.kernel = self._model.kernel
        self.kernel_params_ = self._model.kernel_params_
        self.noise_var_ = self._model.noise_var_
        self.max_var = self.kernel.get_params()['k1__k1__constant_value'] + self.noise_var_

    @property
    def K_inv(self):
        return self._model.K_inv

    def predict(self, x):
        return self._model.predict(x, return_std=True)

    def derivative(self, x):
        """dy_dx[i] = (dk*/dx) alpha  -> (M, d, p);  dsigma_dx[i] = -2 (dk*/dx) K^-1 k*  -> (M, d, 1)   (file:71-107)"""
        o = self._model._query(x, _lib.JAC | _lib.DVAR)
        return np.transpose(o["jac"], (0, 2, 1)), np.transpose(o["dvar"])[:, :, np.newaxis]
