"""Generate golden vectors from the UNMODIFIED reference (runs only in the build container, where /root/reference
exists).  Outputs small .npz fixtures under tests/golden/ (inputs + reference outputs).  TEST INFRASTRUCTURE.

    python oracle/make_golden.py            # regenerates every fixture

The reference imports matplotlib and `Quaternion` at module import (gaussian_process.py:11-13,
policy_transportation.py:9); both are absent here, so empty stand-ins are registered first.
Orientation transport therefore cannot be run through the reference (dependency absent) and has no golden.
"""
import os
import pickle
import sys
import types
import warnings

import numpy as np

REF = os.environ.get("GPT_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "mpl_toolkits", "mpl_toolkits.mplot3d", "Quaternion"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
    sys.path.insert(0, REF)
    import policy_transportation as pt
    from policy_transportation.utils import resample
    return pt, resample


def kparams(k):
    p = k.get_params()
    return dict(c=float(p["k1__k1__constant_value"]), ell=np.atleast_1d(np.asarray(p["k1__k2__length_scale"], dtype=float)),
                s2=float(p["k2__noise_level"]))


def run_gpt(pt, kernel, S, T, traj, delta, do_scale, seed, optimizer="fmin_l_bfgs_b"):
    from policy_transportation.transportation.policy_transportation import PolicyTransportation
    np.random.seed(seed)
    if optimizer == "fmin_l_bfgs_b":
        g = pt.GaussianProcessTransportation(kernel_transport=kernel)
    else:  # deterministic variant: same flow, optimizer=None on the delta map
        g = pt.GaussianProcessTransportation(kernel_transport=kernel)
        g.method = PolicyTransportation(pt.GaussianProcess(kernel=kernel, optimizer=None))
    g.source_distribution, g.target_distribution = S, T
    g.training_traj = traj
    if delta is not None:
        g.training_delta = delta
    g.fit_transportation(do_scale=do_scale, do_rotation=True)
    g.apply_transportation()
    gp = g.method.delta_map
    kp = kparams(gp.kernel)
    out = dict(S=S, T=T, traj_in=traj, do_scale=np.array(do_scale), c=kp["c"], ell=kp["ell"], s2=kp["s2"],
               lml=gp.gp.log_marginal_likelihood_value_, traj_out=g.training_traj, std=g.std,
               R=g.method.affine_transform.rotation_matrix, scale=np.array(float(g.method.affine_transform.scale)),
               S_centroid=g.method.affine_transform.S_centroid, T_centroid=g.method.affine_transform.T_centroid,
               alpha_=gp.gp.alpha_, Ldiag=np.diag(gp.gp.L_).copy())
    if delta is not None:
        out.update(delta_in=delta, delta_out=g.training_delta, var_vel=g.var_vel_transported)
    return out, g


def main():
    warnings.filterwarnings("ignore")
    os.makedirs(OUT, exist_ok=True)
    pt, resample = import_reference()
    from sklearn.gaussian_process.kernels import RBF, WhiteKernel, ConstantKernel as C

    # ---- C1: the 2D demo (example/2D/surface_generalization.py:27-78) ------------------------------------
    data = np.load(os.path.join(REF, "example/2D/data/example.npz"))
    X = resample(data["demo"], num_points=400)
    S = resample(data["floor"], num_points=20)
    T = resample(data["newfloor"], num_points=20)
    dX = np.zeros((len(X), 2))
    dX[:-1] = X[1:] - X[:-1]
    k = C(constant_value=10) * RBF(4 * np.ones(2)) + WhiteKernel(0.01)
    out, g = run_gpt(pt, k, S, T, X, dX, False, seed=0)
    np.savez_compressed(os.path.join(OUT, "c1_demo2d_optimised.npz"), k0_c=10.0, k0_ell=4 * np.ones(2), k0_s2=0.01, **out)
    print("c1 fitted", g.method.delta_map.kernel, out["lml"])
    # deterministic twin at the fitted theta
    kf = C(constant_value=out["c"]) * RBF(out["ell"]) + WhiteKernel(out["s2"])
    out2, g2 = run_gpt(pt, kf, S, T, X, dX, False, seed=0, optimizer=None)
    gp = g2.method.delta_map
    xr = g2.method.affine_transform.predict(X)
    J, Jv = gp.derivative(xr, return_var=True)
    mean, std = gp.predict(xr, return_std=True)
    out2.update(xq=xr, mean=mean, std_raw=std, J=J, Jvar=Jv, dvar=gp.derivative_of_variance(xr),
                X_train=gp.X, Y_train=gp.Y)
    np.savez_compressed(os.path.join(OUT, "c1_demo2d_fixed.npz"), **out2)

    # ---- C2: shipped 3D clouds (distributions/*.pkl 834x3; data/last.npz) ---------------------------------
    S3 = pickle.load(open(os.path.join(REF, "distributions/source.pkl"), "rb"))
    T3 = pickle.load(open(os.path.join(REF, "distributions/target.pkl"), "rb"))
    S3, T3 = np.asarray(S3, dtype=float), np.asarray(T3, dtype=float)
    last = np.load(os.path.join(REF, "data/last.npz"))
    traj, dlt, ori = last["training_traj"], last["training_delta"], last["training_ori"]
    kdef = C(0.1) * RBF(length_scale=[0.1]) + WhiteKernel(0.0001)
    out, g = run_gpt(pt, kdef, S3, T3, traj, dlt, False, seed=0)
    out["ori_in"] = ori
    np.savez_compressed(os.path.join(OUT, "c2_clouds3d_optimised.npz"), k0_c=0.1, k0_ell=np.array([0.1]), k0_s2=1e-4, **out)
    print("c2 fitted", g.method.delta_map.kernel, out["lml"])
    kf = C(constant_value=out["c"]) * RBF(out["ell"]) + WhiteKernel(out["s2"])
    out2, g2 = run_gpt(pt, kf, S3, T3, traj, dlt, False, seed=0, optimizer=None)
    gp = g2.method.delta_map
    xr = g2.method.affine_transform.predict(traj)
    J, Jv = gp.derivative(xr, return_var=True)
    mean, std = gp.predict(xr, return_std=True)
    out2.update(xq=xr, mean=mean, std_raw=std, J=J, Jvar=Jv, dvar=gp.derivative_of_variance(xr), ori_in=ori)
    np.savez_compressed(os.path.join(OUT, "c2_clouds3d_fixed.npz"), **out2)
    # scale variant
    out3, _ = run_gpt(pt, kf, S3, T3 * 1.3 + 0.1, traj, dlt, True, seed=0, optimizer=None)
    np.savez_compressed(os.path.join(OUT, "c2_clouds3d_scaled.npz"), **out3)

    # ---- synthetic ARD / isotropic GPs with LML + gradient at several theta --------------------------------
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from oracle.gp_oracle import synthetic_pairs, helix_queries
    for tag, n, d, ell, c, s2 in [("ard300", 300, 3, [0.1, 0.15, 0.2], 0.1, 1e-4), ("iso500", 500, 3, 0.12, 0.05, 1e-4),
                                  ("ard2d200", 200, 2, [0.2, 0.3], 0.5, 1e-3), ("ard1000", 1000, 3, [0.1, 0.1, 0.1], 0.1, 1e-4)]:
        Sx, Tx = synthetic_pairs(n, d, seed=1)
        Y = Tx - Sx
        kern = C(c) * RBF(ell) + WhiteKernel(s2)
        gp = pt.GaussianProcess(kernel=kern, optimizer=None)
        gp.fit(Sx, Y)
        xq, _ = helix_queries(64, d)
        mean, std = gp.predict(xq, return_std=True)
        J, Jv = gp.derivative(xq, return_var=True)
        thetas, lmls, grads = [], [], []
        rng = np.random.default_rng(5)
        th0 = gp.gp.kernel_.theta
        for i in range(4):
            th = th0 + (0.0 if i == 0 else 1.0) * rng.normal(0, 0.4, th0.shape)
            v, gr = gp.gp.log_marginal_likelihood(th, eval_gradient=True)
            thetas.append(th); lmls.append(v); grads.append(gr)
        np.savez_compressed(os.path.join(OUT, f"syn_{tag}.npz"), X=Sx, Y=Y, c=c, ell=np.atleast_1d(np.asarray(ell, float)), s2=s2,
                            xq=xq, mean=mean, std=std, J=J, Jvar=Jv, dvar=gp.derivative_of_variance(xq),
                            alpha_=gp.gp.alpha_, Ldiag=np.diag(gp.gp.L_).copy(), thetas=np.array(thetas), lmls=np.array(lmls),
                            grads=np.array(grads), K_inv_diag=np.diag(gp.K_inv).copy())
        print(tag, "done")


if __name__ == "__main__":
    main()
