"""Golden vectors for the SURVEY.md section-8 "next" row f4, generated from the UNMODIFIED reference (build container
only): `GaussianProcessTransportationDiffeo` (transportation/gaussian_process_transportation_diffeomorphic.py) and the
active-learning GP (models/gaussian_process_al.py).  TEST INFRASTRUCTURE.

    python oracle/make_golden_f4.py

optuna, quaternion and matplotlib are absent from this image; empty stand-ins are registered so the modules import.  The
orientation branch (needs numpy-quaternion) and `optimize_diffeomorphism` (needs optuna) therefore cannot be run through
the reference and have no golden; `diffeomorphism_error` is driven with a fixed-value trial object instead of a study.
"""
import contextlib
import io
import os
import pickle
import sys
import types
import warnings

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.make_golden import REF, OUT, import_reference, kparams  # noqa: E402


class FixedTrial:
    def __init__(self, value):
        self.value = value

    def suggest_float(self, name, lo, hi, log=False):
        return self.value


def main():
    warnings.filterwarnings("ignore")
    for name in ["optuna", "quaternion"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    pt, resample = import_reference()
    from sklearn.gaussian_process.kernels import RBF, WhiteKernel, ConstantKernel as C
    from policy_transportation.transportation.gaussian_process_transportation_diffeomorphic import GaussianProcessTransportationDiffeo
    from policy_transportation.models.gaussian_process_al import GaussianProcess as GPAL
    quiet = io.StringIO()

    # ---- Diffeo flow on the 2-D demo (fixed theta): apply, samples, invertibility error ----------------------
    data = np.load(os.path.join(REF, "example/2D/data/example.npz"))
    X = resample(data["demo"], num_points=200)
    S = resample(data["floor"], num_points=20)
    T = resample(data["newfloor"], num_points=20)
    dX = np.zeros((len(X), 2))
    dX[:-1] = X[1:] - X[:-1]
    k = C(constant_value=10) * RBF(4 * np.ones(2)) + WhiteKernel(0.01)
    g = GaussianProcessTransportationDiffeo(kernel_transport=k)
    g.source_distribution, g.target_distribution = S, T
    g.training_traj, g.training_delta = X.copy(), dX.copy()
    with contextlib.redirect_stdout(quiet):
        g.fit_transportation(optimize=False)
        g.apply_transportation()
        samples = g.sample_transportation()
    out = dict(S=S, T=T, traj_in=X, delta_in=dX, c=10.0, ell=4 * np.ones(2), s2=0.01, traj_out=g.training_traj, std=g.std,
               delta_out=g.training_delta, var_vel=g.var_vel_transported, traj_rotated=g.traj_rotated, delta_map_mean=g.delta_map_mean,
               samples=samples)
    g2 = GaussianProcessTransportationDiffeo(kernel_transport=k)
    g2.source_distribution, g2.target_distribution = S, T
    g2.training_traj = X.copy()
    with contextlib.redirect_stdout(quiet):
        g2.fit_transportation(optimize=False)
        out["invertibility_error"] = g2.check_invertibility()
    out["traj_rotated_inv"] = g2.traj_rotated_inv
    # one objective evaluation of the length-scale-bound study with the trial value pinned (LML optimisation + 5 restarts inside)
    g3 = GaussianProcessTransportationDiffeo(kernel_transport=k)
    g3.source_distribution, g3.target_distribution = S, T
    g3.training_traj = X.copy()
    np.random.seed(0)
    with contextlib.redirect_stdout(quiet):
        out["diffeo_error_ml5"] = g3.diffeomorphism_error(FixedTrial(5.0))
    kp = kparams(g3.kernel_transport)
    out.update(diffeo_c=kp["c"], diffeo_ell=kp["ell"], diffeo_s2=kp["s2"], diffeo_lml=g3.gp_delta_map.gp.log_marginal_likelihood_value_)
    np.savez_compressed(os.path.join(OUT, "f4_diffeo2d.npz"), **out)
    print("diffeo 2d: invertibility error", out["invertibility_error"], "objective(max_ls=5)", out["diffeo_error_ml5"], g3.kernel_transport)

    # ---- Diffeo flow on the shipped 3-D clouds (fixed theta) ----------------------------------------------------
    S3 = np.asarray(pickle.load(open(os.path.join(REF, "distributions/source.pkl"), "rb")), dtype=float)
    T3 = np.asarray(pickle.load(open(os.path.join(REF, "distributions/target.pkl"), "rb")), dtype=float)
    last = np.load(os.path.join(REF, "data/last.npz"))
    traj, dlt = last["training_traj"], last["training_delta"]
    k3 = C(0.05) * RBF(length_scale=[0.08, 0.1, 0.12]) + WhiteKernel(1e-4)
    g = GaussianProcessTransportationDiffeo(kernel_transport=k3)
    g.source_distribution, g.target_distribution = S3, T3
    g.training_traj, g.training_delta = traj.copy(), dlt.copy()
    with contextlib.redirect_stdout(quiet):
        g.fit_transportation(optimize=False)
        g.apply_transportation()
    np.savez_compressed(os.path.join(OUT, "f4_diffeo3d.npz"), S=S3, T=T3, traj_in=traj, delta_in=dlt, c=0.05, ell=np.array([0.08, 0.1, 0.12]),
                        s2=1e-4, traj_out=g.training_traj, std=g.std, delta_out=g.training_delta, var_vel=g.var_vel_transported)
    print("diffeo 3d done")

    # ---- active-learning GP: 90 points, budget 24 (fixed hyper-parameters so the greedy order is deterministic) -----
    from oracle.gp_oracle import synthetic_pairs, helix_queries
    Sx, Tx = synthetic_pairs(90, 2, seed=4)
    Y = Tx - Sx
    kal = C(0.5, constant_value_bounds="fixed") * RBF(np.array([0.2, 0.3]), length_scale_bounds="fixed") + WhiteKernel(1e-3, noise_level_bounds="fixed")
    al = GPAL(kernel=kal, n_restarts_optimizer=0, n_samples_max=24)
    np.random.seed(3)
    with contextlib.redirect_stdout(quiet):
        al.fit(Sx, Y)
    xq, _ = helix_queries(40, 2)
    mean, std = al.predict(xq)
    dy, ds = al.derivative(xq)
    np.savez_compressed(os.path.join(OUT, "f4_al_fixed.npz"), X=Sx, Y=Y, c=0.5, ell=np.array([0.2, 0.3]), s2=1e-3, n_samples_max=24, seed=3,
                        X_sel=al.X, Y_sel=al.Y, xq=xq, mean=mean, std=std, dy_dx=dy, dsigma_dx=ds, max_var=al.max_var)
    print("al fixed: selected", al.X.shape)
    # optimised variant (hyper-parameters re-fitted after every added point, as the class does by default)
    kal2 = C(0.5) * RBF([0.2, 0.3]) + WhiteKernel(1e-3)
    al2 = GPAL(kernel=kal2, n_restarts_optimizer=0, n_samples_max=20)
    np.random.seed(3)
    with contextlib.redirect_stdout(quiet):
        al2.fit(Sx, Y)
    kp = kparams(al2.kernel)
    mean2, std2 = al2.predict(xq)
    np.savez_compressed(os.path.join(OUT, "f4_al_optimised.npz"), X=Sx, Y=Y, k0_c=0.5, k0_ell=np.array([0.2, 0.3]), k0_s2=1e-3, n_samples_max=20,
                        seed=3, X_sel=al2.X, Y_sel=al2.Y, xq=xq, mean=mean2, std=std2, c=kp["c"], ell=kp["ell"], s2=kp["s2"],
                        lml=al2.gp.log_marginal_likelihood_value_)
    print("al optimised:", al2.kernel)


if __name__ == "__main__":
    main()
