"""CPU (`-m "not gpu"`): the C-ABI library loads and exports every symbol the header declares, the host-side kernel
parsing / gradient mapping, and loud failure without a CUDA device (no compute calls here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from gaussian_process_transportation_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "gptb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gptb_[a-z_A-Z0-9]+)\s*\(", hdr))
    assert declared, "no prototypes found in include/gptb200.h"
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/gptb200.h but not exported"
    assert declared == set(_lib.SYMBOLS), "ctypes binding and header disagree"
    assert _lib.load().gptb_version() >= 100


def test_no_cuda_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gaussian_process_transportation_b200 import _lib
    with pytest.raises(_lib.GptbError):
        _lib.Engine(0)


def test_kernel_parsing_and_gradient_mapping():
    from sklearn.gaussian_process.kernels import RBF, Matern, WhiteKernel, ConstantKernel as C
    from gaussian_process_transportation_b200.kernel_spec import check_supported, map_gradient, read_params, UnsupportedKernel
    k = C(0.1) * RBF(length_scale=[0.1]) + WhiteKernel(1e-4)
    check_supported(k)
    c, ell, s2 = read_params(k, 3)
    assert c == 0.1 and s2 == 1e-4 and np.allclose(ell, [0.1] * 3)
    g = np.array([1.0, 2.0, 3.0, 4.0, 5.0])
    assert np.allclose(map_gradient(k, g, 3), [1.0, 9.0, 5.0])                      # isotropic: summed
    k2 = C(0.1) * RBF(length_scale=[0.1, 0.2, 0.3]) + WhiteKernel(1e-4)
    assert np.allclose(map_gradient(k2, g, 3), g)
    k3 = C(0.1, constant_value_bounds="fixed") * RBF([0.1, 0.2, 0.3]) + WhiteKernel(1e-4, noise_level_bounds="fixed")
    assert np.allclose(map_gradient(k3, g, 3), [2.0, 3.0, 4.0]) and k3.theta.size == 3
    from gaussian_process_transportation_b200.kernel_spec import kernel_kind
    assert kernel_kind(k) == 0
    assert kernel_kind(C(0.1) * Matern(0.1, nu=1.5) + WhiteKernel(1e-4)) == 1
    assert kernel_kind(C(0.1) * Matern([0.1, 0.2], nu=2.5) + WhiteKernel(1e-4)) == 2
    assert kernel_kind(C(0.1) * Matern(0.1, nu=0.5) + WhiteKernel(1e-4)) == 3
    with pytest.raises(UnsupportedKernel):
        check_supported(C(0.1) * Matern(0.1, nu=3.5) + WhiteKernel(1e-4))
    with pytest.raises(UnsupportedKernel):
        check_supported(RBF(0.1))
    with pytest.raises(ValueError):
        read_params(C(1.0) * RBF([1.0, 2.0]) + WhiteKernel(1.0), 3)


def test_affine_transform_matches_oracle():
    from gaussian_process_transportation_b200 import AffineTransform
    from oracle.gp_oracle import OracleAffine, synthetic_pairs
    for d, scale in [(2, False), (3, False), (3, True)]:
        S, T = synthetic_pairs(50, d, seed=3)
        a, b = AffineTransform(do_scale=scale), OracleAffine(do_scale=scale)
        a.fit(S, T); b.fit(S, T * (1.4 if scale else 1.0))
        if not scale:
            assert np.array_equal(a.rotation_matrix, b.rotation_matrix)
            assert np.array_equal(a.predict(S), b.predict(S))
        assert a.derivative(S).shape == (50, d, d)
        assert abs(np.linalg.det(a.rotation_matrix) - 1.0) < 1e-12


def test_oracle_quaternion_restatement_is_self_consistent():
    """The Bar-Itzhack restatement in the oracle (numpy-quaternion itself is absent): on orthogonal matrices it returns the generating
    quaternion, against scipy's Rotation, up to sign."""
    from scipy.spatial.transform import Rotation
    from oracle.gp_oracle import quat_from_matrix_nonorthogonal, quat_mul
    rot = Rotation.random(30, random_state=0)
    a = quat_from_matrix_nonorthogonal(rot.as_matrix())
    b = np.roll(rot.as_quat(), 1, axis=1)                       # scipy is (x, y, z, w)
    s = np.sign(np.sum(a * b, axis=1))[:, None]
    assert np.max(np.abs(a * s - b)) < 1e-12
    q = np.roll(Rotation.random(30, random_state=1).as_quat(), 1, axis=1)
    prod = np.roll((rot * Rotation.from_quat(np.roll(q, -1, axis=1))).as_quat(), 1, axis=1)
    mine = quat_mul(b, q)
    s = np.sign(np.sum(mine * prod, axis=1))[:, None]
    assert np.max(np.abs(mine * s - prod)) < 1e-12


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) prints one JSON line with the
    contract keys; smallest workload so the CPU suite stays fast."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ["impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"]:
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["vs_baseline"] is None
    # "reference" where baseline/_ref holds the unmodified reference (build container, GPU box), "port" on a bare checkout
    from baseline.reference_shim import available
    assert line["cpu_baseline"]["kind"] == ("reference" if available() else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_variance_mode_names():
    from gaussian_process_transportation_b200 import _lib
    assert _lib.parse_variance_mode("fp64") == (0, 6)
    assert _lib.parse_variance_mode("int8x6") == (1, 6)
    assert _lib.parse_variance_mode("INT8W5") == (2, 5)
    for bad in ("int8", "int8y5", "fp32", "int8x"):
        with pytest.raises(ValueError):
            _lib.parse_variance_mode(bad)


def test_bench_alignment_matches_the_package_and_the_oracle():
    """bench.py aligns the synthetic clouds with its own few lines of Kabsch (so that the GPU arm imports nothing from oracle/);
    they must be the reference's AffineTransform."""
    import contextlib, io, sys
    sys.path.insert(0, ROOT)
    import bench
    from gaussian_process_transportation_b200 import AffineTransform
    from oracle.gp_oracle import OracleAffine
    S, T = bench.synthetic_pairs(200, 3, seed=0)
    R, Sc, Tc = bench.kabsch(S, T)
    with contextlib.redirect_stdout(io.StringIO()):
        a = AffineTransform(); a.fit(S, T)
        b = OracleAffine(); b.fit(S, T)
    assert np.allclose(R, a.rotation_matrix, atol=1e-14) and np.allclose(R, b.rotation_matrix, atol=1e-14)
    _, _, _, _, _, Sr, D = bench.aligned_training_set(200)
    assert np.allclose(Sr, a.predict(S), atol=1e-14) and np.allclose(D, T - a.predict(S), atol=1e-14)
    from gaussian_process_transportation_b200 import _lib
    assert _lib.parse_variance_mode("int8w5p") == (3, 5)


def test_unmodified_reference_and_oracle_port_agree():
    """baseline/_ref (the unmodified reference, when installed by build()) against the oracle port on a small problem: the CPU arm of
    bench.py times the former, the tests on the GPU box use the latter."""
    import contextlib, io, sys, warnings
    sys.path.insert(0, ROOT)
    from baseline.reference_shim import available, import_reference
    if not available():
        pytest.skip("baseline/_ref is not installed on this checkout")
    from sklearn.gaussian_process.kernels import RBF, WhiteKernel, ConstantKernel as C
    from oracle.gp_oracle import SkGaussianProcess, synthetic_pairs
    warnings.filterwarnings("ignore")
    pt = import_reference()
    S, T = synthetic_pairs(150, 3, seed=2)
    k = C(0.1) * RBF([0.1, 0.15, 0.2]) + WhiteKernel(1e-4)
    xq = np.random.default_rng(0).random((40, 3))
    outs = []
    for cls in (pt.GaussianProcess, SkGaussianProcess):
        gp = cls(kernel=k, optimizer=None)
        with contextlib.redirect_stdout(io.StringIO()):
            gp.fit(S, T - S)
        m, s = gp.predict(xq, return_std=True)
        J, V = gp.derivative(xq, return_var=True)
        outs.append((m, s, J, V, gp.derivative_of_variance(xq)))
    for a, b in zip(*outs):
        assert np.max(np.abs(np.asarray(a) - np.asarray(b))) < 1e-12
