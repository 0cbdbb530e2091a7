"""CPU: the oracle restatements against the golden vectors produced by the unmodified reference
(oracle/make_golden.py).  Tolerances: the two restatements differ from the reference only by the inv-vs-Cholesky
route (quirk Q3), i.e. by ~cond(K)*eps."""
import os
import warnings

import numpy as np
import pytest

from oracle.gp_oracle import ChoGP, OracleGPT, SkGaussianProcess, OraclePolicyTransportation

warnings.filterwarnings("ignore")


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


@pytest.mark.parametrize("name", ["syn_ard300.npz", "syn_iso500.npz", "syn_ard2d200.npz", "syn_ard1000.npz"])
def test_chogp_matches_reference(golden_dir, name):
    g = load(golden_dir, name)
    gp = ChoGP(float(g["c"]), g["ell"], float(g["s2"])).fit(g["X"], g["Y"])
    mean, std = gp.predict(g["xq"], return_std=True)
    J, Jv = gp.derivative(g["xq"], return_var=True)
    assert rel(gp.alpha, g["alpha_"]) < 1e-9
    assert rel(np.diag(gp.L), g["Ldiag"]) < 1e-12
    assert rel(mean, g["mean"]) < 1e-10
    assert np.max(np.abs(std - g["std"])) / np.sqrt(float(g["c"]) + float(g["s2"])) < 1e-9
    assert rel(J, g["J"]) < 1e-9
    assert rel(Jv, g["Jvar"]) < 1e-7
    assert rel(gp.derivative_of_variance(g["xq"]), g["dvar"]) < 1e-7


@pytest.mark.parametrize("name", ["syn_ard300.npz", "syn_iso500.npz", "syn_ard2d200.npz"])
def test_chogp_lml_and_gradient(golden_dir, name):
    g = load(golden_dir, name)
    d = g["X"].shape[1]
    iso = g["ell"].size == 1
    for th, lml, grad in zip(g["thetas"], g["lmls"], g["grads"]):
        c = np.exp(th[0]); ell = np.exp(th[1:-1]); s2 = np.exp(th[-1])
        gp = ChoGP(c, ell, s2).fit(g["X"], g["Y"])
        v, gc, gl, gs = gp.lml(eval_gradient=True)
        mine = np.concatenate([[gc], [gl.sum()] if iso else gl, [gs]])
        assert abs(v - lml) <= 1e-10 * abs(lml)
        assert rel(mine, grad) < 1e-8
        assert (len(ell) == 1) == iso and d >= 1


@pytest.mark.parametrize("name", ["c1_demo2d_fixed.npz", "c2_clouds3d_fixed.npz", "c2_clouds3d_scaled.npz"])
def test_transport_flow_restatement(golden_dir, name):
    from sklearn.gaussian_process.kernels import RBF, WhiteKernel, ConstantKernel as C
    g = load(golden_dir, name)
    kern = C(float(g["c"])) * RBF(g["ell"]) + WhiteKernel(float(g["s2"]))
    t = OracleGPT(kern, optimizer=None)
    t.source_distribution, t.target_distribution = g["S"], g["T"]
    t.training_traj, t.training_delta = g["traj_in"], g["delta_in"]
    t.fit_transportation(do_scale=bool(g["do_scale"]))
    t.apply_transportation()
    assert rel(t.method.affine_transform.rotation_matrix, g["R"]) < 1e-14
    assert rel(t.training_traj, g["traj_out"]) < 1e-12
    assert rel(t.std, g["std"]) < 1e-9
    assert rel(t.training_delta, g["delta_out"]) < 1e-10
    assert rel(t.var_vel_transported, g["var_vel"]) < 1e-8


def test_chogp_flow_matches_reference_c2(golden_dir):
    """Cholesky-only restatement through the transport flow (differs from the reference by quirk Q3 only)."""
    g = load(golden_dir, "c2_clouds3d_fixed.npz")

    class Map:
        def __init__(self):
            self.gp = ChoGP(float(g["c"]), g["ell"], float(g["s2"]))
        def fit(self, X, Y): self.gp.fit(X, Y)
        def predict(self, x, return_std=False): return self.gp.predict(x, return_std=return_std)
        def derivative(self, x, return_var=False): return self.gp.derivative(x, return_var=return_var)

    m = OraclePolicyTransportation(Map())
    m.fit(g["S"], g["T"])
    xt, std = m.transport(g["traj_in"])
    v, vv = m.transport_velocity(g["traj_in"], g["delta_in"])
    assert rel(xt, g["traj_out"]) < 1e-11
    assert np.max(np.abs(std - g["std"])) / np.sqrt(float(g["c"]) + float(g["s2"])) < 1e-8
    assert rel(v, g["delta_out"]) < 1e-9
    assert rel(vv, g["var_vel"]) < 1e-6


def test_quaternion_restatement_self_consistency():
    from oracle.gp_oracle import quat_from_matrix_nonorthogonal, quat_mul
    rng = np.random.default_rng(0)
    q = rng.normal(size=(20, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
                  np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
                  np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], -2)
    qq = quat_from_matrix_nonorthogonal(R)
    s = np.sign(np.sum(qq * q, axis=1))[:, None]
    assert np.max(np.abs(qq * s - q)) < 1e-12
    one = np.array([1.0, 0, 0, 0])
    assert np.allclose(quat_mul(one, q), q)


# ---------------------------------------------------------------------------------------------------------------
# SURVEY section-8 "next" row f4: the oracle restatements of the Diffeo flow and the active-learning GP against
# goldens from the unmodified reference (oracle/make_golden_f4.py)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["f4_diffeo2d.npz", "f4_diffeo3d.npz"])
def test_oracle_diffeo_flow_matches_reference_golden(golden_dir, name):
    import contextlib, io
    from sklearn.gaussian_process.kernels import RBF, WhiteKernel, ConstantKernel as C
    from oracle.gp_oracle import OracleDiffeo
    g = np.load(os.path.join(golden_dir, name))
    k = C(float(g["c"])) * RBF(g["ell"]) + WhiteKernel(float(g["s2"]))
    t = OracleDiffeo(k)
    t.source_distribution, t.target_distribution = g["S"], g["T"]
    t.training_traj, t.training_delta = g["traj_in"].copy(), g["delta_in"].copy()
    with contextlib.redirect_stdout(io.StringIO()):
        t.fit_transportation(optimize=False)
        t.apply_transportation()
    assert np.allclose(t.training_traj, g["traj_out"], rtol=1e-11, atol=1e-13)
    assert np.allclose(t.training_delta, g["delta_out"], rtol=1e-10, atol=1e-13)
    assert np.allclose(t.std, g["std"], rtol=0, atol=1e-10)
    assert np.allclose(t.var_vel_transported, g["var_vel"], rtol=1e-8, atol=1e-16)
    if "invertibility_error" in g.files:
        t2 = OracleDiffeo(k)
        t2.source_distribution, t2.target_distribution = g["S"], g["T"]
        t2.training_traj = g["traj_in"].copy()
        with contextlib.redirect_stdout(io.StringIO()):
            t2.fit_transportation(optimize=False)
            assert abs(t2.check_invertibility() - float(g["invertibility_error"])) < 1e-9 * float(g["invertibility_error"])


def test_oracle_active_learning_gp_matches_reference_golden(golden_dir):
    import contextlib, io
    from sklearn.gaussian_process.kernels import RBF, WhiteKernel, ConstantKernel as C
    from oracle.gp_oracle import OracleGPAL
    g = np.load(os.path.join(golden_dir, "f4_al_fixed.npz"))
    k = C(float(g["c"]), constant_value_bounds="fixed") * RBF(g["ell"], length_scale_bounds="fixed") + \
        WhiteKernel(float(g["s2"]), noise_level_bounds="fixed")
    np.random.seed(int(g["seed"]))
    with contextlib.redirect_stdout(io.StringIO()):
        al = OracleGPAL(k, n_restarts_optimizer=0, n_samples_max=int(g["n_samples_max"])).fit(g["X"], g["Y"])
    assert np.array_equal(al.X, g["X_sel"])
    mean, std = al.predict(g["xq"])
    dy, ds = al.derivative(g["xq"])
    assert np.allclose(mean, g["mean"], rtol=1e-11, atol=1e-13) and np.allclose(std, g["std"], rtol=0, atol=1e-10)
    assert np.allclose(dy, g["dy_dx"], rtol=1e-10, atol=1e-12) and np.allclose(ds, g["dsigma_dx"], rtol=1e-7, atol=1e-12)


def test_digit_plane_identity_and_truncation_bound():
    """The 8-bit digit split of csrc/digits.cuh restated in numpy: digits stay in int8, they reproduce Q exactly (carries included), and a
    sliced dot product with S = 5 planes is within the 40-bit truncation bound of the FP64 one."""
    from oracle.digit_planes import combine8, digit_scale8, sliced_dot, split8
    rng = np.random.default_rng(0)
    for S in (4, 5, 6):
        x = np.concatenate([rng.standard_normal(4000) * 10.0 ** rng.integers(-12, 1, 4000), [0.0, 1.0, -1.0, 0.999999, -0.999999]])
        scale = digit_scale8(np.abs(x).max())
        planes, q = split8(x, S, scale)
        assert planes.dtype == np.int8 and planes.shape == (S, x.size)
        assert np.array_equal(combine8(planes), q)                                  # exact, including the byte carries
        assert np.max(np.abs(q * (scale / 256.0 ** S) - x)) <= 0.5 * scale / 256.0 ** S + 4e-16 * scale          # half a unit of the last digit (+ FP64 rounding)
    a = rng.random(4096) * 0.1
    b = rng.standard_normal(4096) * np.exp(-rng.random(4096) * 20)
    sa, sb = digit_scale8(0.1), digit_scale8(np.abs(b).max())
    pa, _ = split8(a, 5, sa)
    pb, _ = split8(b, 5, sb)
    ref = float(np.dot(a, b))
    assert abs(sliced_dot(pa, pb, sa, sb) - ref) < 4096 * 8 * 2.0 ** -40 * sa * sb


def test_cholesky_oracle_against_headline_size_golden(golden_dir):
    """The Cholesky-only restatement (what the GPU tests at N = 16384 / 32768 are checked against) pinned at the headline size: the
    N = 4096 golden of the benchmark's hyper-parameters, produced by the unmodified reference (oracle/make_golden_large.py)."""
    from oracle.gp_oracle import ChoGP
    tr = np.load(os.path.join(golden_dir, "n4096_train.npz"))
    g = np.load(os.path.join(golden_dir, "n4096_c3.npz"))
    gp = ChoGP(float(g["c"]), g["ell"], float(g["s2"])).fit(tr["X"], tr["Y"])
    xq = tr["xq"][::4]
    mean, std = gp.predict(xq, return_std=True)
    J = gp.derivative(xq)
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
    assert rel(mean, g["mean"][::4]) < 1e-10 and rel(J, g["J"][::4]) < 1e-9
    assert np.max(np.abs(std[:, 0] - g["std0"][::4])) / np.sqrt(float(g["c"]) + float(g["s2"])) < 1e-9
    assert abs(gp.lml() - float(g["lml"])) < 1e-9 * abs(float(g["lml"]))
