"""ctypes binding of libgptb200.so (include/gptb200.h).  There is no fallback: a missing library or a missing CUDA
device raises immediately."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgptb200.so")

MEAN, STD, JAC, JACVAR, AFFINE_IN, TRANSPORT, VELOCITY, JPHI, DVAR = 0x1, 0x2, 0x4, 0x8, 0x10, 0x20, 0x40, 0x80, 0x100

_lib = None
_dp = C.POINTER(C.c_double)

SYMBOLS = {
    "gptb_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "gptb_destroy": (None, [C.c_void_p]),
    "gptb_last_error": (C.c_char_p, [C.c_void_p]),
    "gptb_version": (C.c_int, []),
    "gptb_set_train": (C.c_int, [C.c_void_p, _dp, _dp, C.c_int64, C.c_int, C.c_int]),
    "gptb_set_kernel_kind": (C.c_int, [C.c_void_p, C.c_int]),
    "gptb_factorize": (C.c_int, [C.c_void_p, C.c_double, _dp, C.c_double, C.c_double, _dp]),
    "gptb_append_point": (C.c_int, [C.c_void_p, _dp, _dp, _dp]),
    "gptb_lml": (C.c_int, [C.c_void_p, C.c_double, _dp, C.c_double, C.c_double, C.c_int, _dp, _dp]),
    "gptb_set_variance_mode": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "gptb_prepare_variance": (C.c_int, [C.c_void_p]),
    "gptb_set_variance_guard": (C.c_int, [C.c_void_p, C.c_double]),
    "gptb_variance_guard_report": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), _dp, _dp, _dp]),
    "gptb_set_affine": (C.c_int, [C.c_void_p, _dp, C.c_double, _dp, _dp]),
    "gptb_query": (C.c_int, [C.c_void_p, _dp, C.c_int64, C.c_uint32, _dp] + [_dp] * 9),
    "gptb_query_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32, C.c_void_p] + [C.c_void_p] * 9),
    "gptb_query_grid": (C.c_int, [C.c_void_p, _dp, _dp, C.POINTER(C.c_int64), C.c_int64, C.c_int64, C.c_uint32, _dp, C.c_int64, _dp]),
    "gptb_rollout_min_variance": (C.c_int, [C.c_void_p, _dp, C.c_int64, C.c_int, C.c_double, _dp]),
    "gptb_query_cov": (C.c_int, [C.c_void_p, _dp, C.c_int64, _dp, _dp]),
    "gptb_transport_orientation": (C.c_int, [C.c_void_p, _dp, _dp, C.c_int64, _dp, _dp]),
    "gptb_transport_orientation_diffeo": (C.c_int, [C.c_void_p, _dp, _dp, C.c_int64, _dp]),
    "gptb_transport_stiffness": (C.c_int, [C.c_void_p, _dp, _dp, C.c_int64, _dp, _dp]),
    "gptb_export_L": (C.c_int, [C.c_void_p, _dp]),
    "gptb_export_alpha": (C.c_int, [C.c_void_p, _dp]),
    "gptb_export_Kinv": (C.c_int, [C.c_void_p, _dp]),
    "gptb_state_alloc": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int]),
    "gptb_state_buffer": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "gptb_state_commit": (C.c_int, [C.c_void_p]),
    "gptb_launch_count": (C.c_int64, [C.c_void_p]),
    "gptb_kernel_time": (C.c_int, [C.c_void_p, C.c_int, _dp, C.POINTER(C.c_int64)]),
    "gptb_timing_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "gptb_timing_reset": (C.c_int, [C.c_void_p]),
    "gptb_stream": (C.c_void_p, [C.c_void_p]),
    "gptb_set_trailing_variant": (C.c_int, [C.c_void_p, C.c_int]),
    "gptb_set_workspace_limit": (C.c_int, [C.c_void_p, C.c_int64]),
    "gptb_set_query_pipeline": (C.c_int, [C.c_void_p, C.c_int]),
    "gptb_set_spatial": (C.c_int, [C.c_void_p, C.c_int]),
    "gptb_executed_products": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_int]),
    "gptb_set_debug_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "gptb_debug_read_profile": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "gptb_test_gemm_nt": (C.c_int, [C.c_void_p, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "gptb_test_potrf_tile": (C.c_int, [C.c_void_p, _dp, _dp, _dp, C.POINTER(C.c_int)]),
}


def load():
    """Load libgptb200.so and declare every prototype of include/gptb200.h.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("GPTB_LIB_PATH", LIB_PATH)       # developer override: A/B of two builds of the same library
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `make` (or __graft_entry__.build()). "
            "gaussian_process_transportation_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        if not hasattr(lib, name) and path != LIB_PATH:
            continue                                      # an older build may lack newer entry points
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def as_f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def ptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


class GptbError(RuntimeError):
    pass


def parse_variance_mode(name):
    """'fp64' -> (0, 6); 'int8xS' -> (1, S) (7-bit digit planes); 'int8wS' -> (2, S) (8-bit digit planes); 'int8w5p' -> (3, 5)
    (five 8-bit planes plus the first dropped diagonal of plane products)."""
    vm = str(name).lower()
    if vm == "fp64":
        return 0, 6
    if vm == "int8w5p":
        return 3, 5
    if len(vm) == 6 and vm[:4] == "int8" and vm[4] in "xw" and vm[5].isdigit():
        return (1 if vm[4] == "x" else 2), int(vm[5])
    raise ValueError(f"unknown variance_mode {name!r} (use 'fp64', 'int8x5'..'int8x7', 'int8w4'..'int8w6' or 'int8w5p')")


class Engine:
    """One handle = one device + stream + model state (not thread-safe, cf. include/gptb200.h)."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.gptb_create(int(device), C.byref(h))
        if rc != 0:
            raise GptbError(
                f"gptb_create(device={device}) failed with status {rc}"
                + (": no CUDA device is visible and this package has no CPU fallback" if rc == -4 else ""))
        self.h = h
        self.device = int(device)
        self.N = self.d = self.p = 0
        self.spatial = False

    def close(self):
        if getattr(self, "h", None):
            self.lib.gptb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ---------------------------------------------------------------------------------------------
    def _check(self, rc, what):
        if rc < 0:
            raise GptbError(f"{what}: {self.lib.gptb_last_error(self.h).decode()} (status {rc})")
        return rc

    def error(self):
        return self.lib.gptb_last_error(self.h).decode()

    # -- model -----------------------------------------------------------------------------------------------
    def set_train(self, X, Y):
        X = as_f64(X)
        Y = as_f64(Y)
        if Y.ndim == 1:
            Y = Y[:, None]
        self.N, self.d = X.shape
        self.p = Y.shape[1]
        self._check(self.lib.gptb_set_train(self.h, ptr(X), ptr(Y), self.N, self.d, self.p), "gptb_set_train")

    def set_kernel_kind(self, kind):
        self._check(self.lib.gptb_set_kernel_kind(self.h, int(kind)), "gptb_set_kernel_kind")

    def _ell(self, ell):
        e = np.atleast_1d(np.asarray(ell, dtype=np.float64)).ravel()
        if e.size == 1:
            e = np.repeat(e, self.d)
        if e.size != self.d:
            raise ValueError(f"length_scale has {e.size} entries for {self.d} input dimensions")
        return np.ascontiguousarray(e)

    def factorize(self, c, ell, s2, jitter, want_lml=True):
        """Returns (info, lml); info > 0 means not positive definite."""
        e = self._ell(ell)
        out = C.c_double(float("nan"))
        rc = self.lib.gptb_factorize(self.h, float(c), ptr(e), float(s2), float(jitter), C.byref(out) if want_lml else None)
        self._check(rc, "gptb_factorize")
        return rc, out.value

    def append_point(self, x, y, want_lml=False):
        """Rank-1 append of one training point at the current hyper-parameters; returns (info, lml)."""
        x, y = as_f64(x).ravel(), as_f64(y).ravel()
        if x.size != self.d or y.size != self.p:
            raise ValueError(f"append_point needs x ({self.d},) and y ({self.p},)")
        out = C.c_double(float("nan"))
        rc = self.lib.gptb_append_point(self.h, ptr(x), ptr(y), C.byref(out) if want_lml else None)
        self._check(rc, "gptb_append_point")
        if rc == 0:
            self.N += 1
        return rc, out.value

    def lml(self, c, ell, s2, jitter, want_grad=True):
        """Returns (info, lml, grad[2+d]) with grad = d/dlog [c, ell_0.., s2]."""
        e = self._ell(ell)
        out = C.c_double(float("nan"))
        g = np.zeros(2 + self.d)
        rc = self.lib.gptb_lml(self.h, float(c), ptr(e), float(s2), float(jitter), int(bool(want_grad)), C.byref(out), ptr(g))
        self._check(rc, "gptb_lml")
        return rc, out.value, g

    def set_variance_mode(self, mode, slices=6):
        """0 = FP64 DMMA (default); 1 = INT8-sliced tcgen05 path with `slices` 7-bit digit planes (5..7);
        2 = the same path with `slices` 8-bit digit planes (4..6).  A string ("fp64", "int8x6", "int8w5") is parsed."""
        if isinstance(mode, str):
            mode, slices = parse_variance_mode(mode)
        self._check(self.lib.gptb_set_variance_mode(self.h, int(mode), int(slices)), "gptb_set_variance_mode")

    def set_variance_guard(self, threshold):
        """Probe threshold of the INT8-sliced path's run-time accuracy guard (relative to sqrt(c + s2); default 2e-8, 0 = off)."""
        self._check(self.lib.gptb_set_variance_guard(self.h, float(threshold)), "gptb_set_variance_guard")

    def variance_guard(self):
        """What the guard decided for the current model: requested / used digit planes (0 = FP64 path), probe errors, threshold."""
        rq, us, ex = C.c_int(0), C.c_int(0), C.c_int(0)
        e1, e0, th = C.c_double(0.0), C.c_double(0.0), C.c_double(0.0)
        self._check(self.lib.gptb_variance_guard_report(self.h, C.byref(rq), C.byref(us), C.byref(ex), C.byref(e1), C.byref(e0), C.byref(th)),
                    "gptb_variance_guard_report")
        return {"requested_slices": rq.value, "used_slices": us.value, "used_extra_diagonal": ex.value, "probe_err": e1.value, "first_err": e0.value,
                "threshold": th.value}

    def prepare_variance(self):
        self._check(self.lib.gptb_prepare_variance(self.h), "gptb_prepare_variance")

    def set_affine(self, R=None, s=1.0, Sbar=None, Tbar=None):
        if R is None:
            self._check(self.lib.gptb_set_affine(self.h, None, 1.0, None, None), "gptb_set_affine")
            return
        R, Sbar, Tbar = as_f64(R), as_f64(Sbar), as_f64(Tbar)
        self._check(self.lib.gptb_set_affine(self.h, ptr(R), float(s), ptr(Sbar), ptr(Tbar)), "gptb_set_affine")

    def query(self, x, flags, vel=None):
        """Host-pointer query; returns a dict of numpy arrays in the reference layouts."""
        x = as_f64(x)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise ValueError(f"query points must have shape (M, {self.d}), got {x.shape}")
        M, d, p = x.shape[0], self.d, self.p
        o = {}
        if flags & MEAN: o["mean"] = np.empty((M, p))
        if flags & STD: o["std"] = np.empty((M, p))
        if flags & JAC: o["jac"] = np.empty((M, p, d))
        if flags & JACVAR: o["jacvar"] = np.empty((M, p, d))
        if flags & TRANSPORT: o["xhat"] = np.empty((M, d))
        if flags & VELOCITY:
            o["vhat"] = np.empty((M, d))
            if flags & JACVAR: o["vvar"] = np.empty((M, p))
        if flags & JPHI: o["jphi"] = np.empty((M, d, d))
        if flags & DVAR: o["dvar"] = np.empty((d, M))
        v = as_f64(vel) if vel is not None else None
        if M == 0:
            return o
        g = o.get
        rc = self.lib.gptb_query(self.h, ptr(x), M, int(flags), ptr(v), ptr(g("mean")), ptr(g("std")), ptr(g("jac")), ptr(g("jacvar")),
                                 ptr(g("xhat")), ptr(g("vhat")), ptr(g("vvar")), ptr(g("jphi")), ptr(g("dvar")))
        self._check(rc, "gptb_query")
        return o

    def query_grid(self, origin, step, dims, flags, first=0, count=None, sample_stride=0):
        """Dense lattice query with on-device point generation and reduction (include/gptb200.h gptb_query_grid).  Returns
        {"columns": names, "stats": (ncol, 4) [sum, sumsq, min, max], "sample_index": lattice indices, "sample": (n, ncol)}."""
        origin, step = as_f64(origin), as_f64(step)
        dims_a = (C.c_int64 * len(dims))(*[int(v) for v in dims])
        total = int(np.prod([int(v) for v in dims]))
        count = total - first if count is None else int(count)
        p, d = self.p, self.d
        cols = ([f"mean{o}" for o in range(p)] if flags & MEAN else []) + (["std"] if flags & STD else []) + \
               ([f"jac{o}{a}" for o in range(p) for a in range(d)] if flags & JAC else [])
        stats = np.zeros((len(cols), 4))
        j0 = -(-first // sample_stride) if sample_stride > 0 else 0
        j1 = -(-(first + count) // sample_stride) if sample_stride > 0 else 0
        sample = np.zeros((max(j1 - j0, 0), len(cols)))
        self._check(self.lib.gptb_query_grid(self.h, ptr(origin), ptr(step), dims_a, int(first), count, int(flags), ptr(stats), int(sample_stride),
                                             ptr(sample) if sample.size else None), "gptb_query_grid")
        return {"columns": cols, "stats": stats, "sample_index": np.arange(j0, j1) * sample_stride if sample_stride > 0 else np.zeros(0, int), "sample": sample}

    def rollout_min_variance(self, start, steps, gain=1.0):
        """K minimum-variance stabilised rollouts of `steps` steps on the device (include/gptb200.h); returns (steps, K, d)."""
        start = as_f64(start)
        if start.ndim != 2 or start.shape[1] != self.d:
            raise ValueError(f"start points must have shape (K, {self.d}), got {start.shape}")
        traj = np.empty((int(steps), start.shape[0], self.d))
        self._check(self.lib.gptb_rollout_min_variance(self.h, ptr(start), start.shape[0], int(steps), float(gain), ptr(traj)), "gptb_rollout_min_variance")
        return traj

    def query_cov(self, x):
        """Posterior mean (M,p) and joint covariance (M,M) on the device (k(x,x) + s2 I - K* K^-1 K*^T)."""
        x = as_f64(x)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise ValueError(f"query points must have shape (M, {self.d}), got {x.shape}")
        M = x.shape[0]
        mean, cov = np.empty((M, self.p)), np.empty((M, M))
        if M:
            self._check(self.lib.gptb_query_cov(self.h, ptr(x), M, ptr(mean), ptr(cov)), "gptb_query_cov")
        return mean, cov

    def transport_orientation(self, pos, ori):
        """q_hat = quat(Jphi(pos)) (x) ori on the device; returns (ori_out (M,4), jphi (M,3,3))."""
        pos, ori = as_f64(pos), as_f64(ori)
        M = pos.shape[0]
        if pos.shape != (M, 3) or ori.shape != (M, 4):
            raise ValueError(f"orientation transport needs pos (M,3) and ori (M,4), got {pos.shape} and {ori.shape}")
        out, jphi = np.empty((M, 4)), np.empty((M, 3, 3))
        if M:
            self._check(self.lib.gptb_transport_orientation(self.h, ptr(pos), ptr(ori), M, ptr(out), ptr(jphi)), "gptb_transport_orientation")
        return out, jphi

    def transport_orientation_diffeo(self, pos, ori):
        """q_hat = quat(I + Jpsi(gamma(pos))) (x) (quat(R) (x) ori) on the device (the diffeomorphic variant's composition)."""
        pos, ori = as_f64(pos), as_f64(ori)
        M = pos.shape[0]
        if pos.shape != (M, 3) or ori.shape != (M, 4):
            raise ValueError(f"orientation transport needs pos (M,3) and ori (M,4), got {pos.shape} and {ori.shape}")
        out = np.empty((M, 4))
        if M:
            self._check(self.lib.gptb_transport_orientation_diffeo(self.h, ptr(pos), ptr(ori), M, ptr(out)), "gptb_transport_orientation_diffeo")
        return out

    def transport_stiffness(self, pos, stiff):
        """K_hat = Jphi(pos) K Jphi(pos)^T on the device; returns (stiff_out (M,d,d), jphi (M,d,d))."""
        pos, stiff = as_f64(pos), as_f64(stiff)
        M, d = pos.shape
        if stiff.shape != (M, d, d):
            raise ValueError(f"stiffness transport needs pos (M,d) and stiffness (M,d,d), got {pos.shape} and {stiff.shape}")
        out, jphi = np.empty((M, d, d)), np.empty((M, d, d))
        if M:
            self._check(self.lib.gptb_transport_stiffness(self.h, ptr(pos), ptr(stiff), M, ptr(out), ptr(jphi)), "gptb_transport_stiffness")
        return out, jphi

    def query_dev(self, x_ptr, M, flags, vel_ptr=0, mean=0, std=0, jac=0, jacvar=0, xhat=0, vhat=0, vvar=0, jphi=0, dvar=0):
        rc = self.lib.gptb_query_dev(self.h, x_ptr, M, int(flags), vel_ptr or None, mean or None, std or None, jac or None, jacvar or None,
                                     xhat or None, vhat or None, vvar or None, jphi or None, dvar or None)
        self._check(rc, "gptb_query_dev")

    def export_L(self):
        L = np.empty((self.N, self.N))
        self._check(self.lib.gptb_export_L(self.h, ptr(L)), "gptb_export_L")
        return L

    def export_alpha(self):
        a = np.empty((self.N, self.p))
        self._check(self.lib.gptb_export_alpha(self.h, ptr(a)), "gptb_export_alpha")
        return a

    def export_Kinv(self):
        K = np.empty((self.N, self.N))
        self._check(self.lib.gptb_export_Kinv(self.h, ptr(K)), "gptb_export_Kinv")
        return K

    # -- instrumentation -------------------------------------------------------------------------------------
    def launch_count(self):
        return int(self.lib.gptb_launch_count(self.h))

    def timing(self, on):
        self.lib.gptb_timing_enable(self.h, int(on))

    def timing_reset(self):
        self.lib.gptb_timing_reset(self.h)

    def kernel_time(self, which):
        ms = C.c_double(0.0)
        n = C.c_int64(0)
        self._check(self.lib.gptb_kernel_time(self.h, which, C.byref(ms), C.byref(n)), "gptb_kernel_time")
        return ms.value, n.value

    def stream(self):
        return self.lib.gptb_stream(self.h)

    def set_trailing_variant(self, variant):
        self._check(self.lib.gptb_set_trailing_variant(self.h, int(variant)), "gptb_set_trailing_variant")

    def set_spatial(self, on):
        """Morton-ordered training points + sorted query batches + zero-plane skipping (include/gptb200.h); before set_train."""
        self._check(self.lib.gptb_set_spatial(self.h, int(on)), "gptb_set_spatial")
        self.spatial = bool(on)

    def set_query_pipeline(self, on):
        """INT8-sliced path: overlap the generator of the next batch with the products of the current one (default off:
        no gain under the power cap, see include/gptb200.h)."""
        self._check(self.lib.gptb_set_query_pipeline(self.h, int(bool(on))), "gptb_set_query_pipeline")

    def executed_products(self, reset=False):
        """(plane pair, k-chunk, tile) products issued by the INT8-sliced product kernel since the last reset (include/gptb200.h)."""
        n = C.c_int64(0)
        self._check(self.lib.gptb_executed_products(self.h, C.byref(n), int(bool(reset))), "gptb_executed_products")
        return int(n.value)

    def set_debug_option(self, name, value):
        self._check(self.lib.gptb_set_debug_option(self.h, str(name).encode(), int(value)), "gptb_set_debug_option")

    def set_workspace_limit(self, nbytes):
        self._check(self.lib.gptb_set_workspace_limit(self.h, int(nbytes)), "gptb_set_workspace_limit")

    # -- state exchange --------------------------------------------------------------------------------------
    def state_alloc(self, N, d, p, with_variance=True):
        self.N, self.d, self.p = int(N), int(d), int(p)
        self._check(self.lib.gptb_state_alloc(self.h, self.N, self.d, self.p, int(with_variance)), "gptb_state_alloc")

    def state_buffer(self, which):
        p = C.c_void_p()
        n = C.c_int64(0)
        self._check(self.lib.gptb_state_buffer(self.h, which, C.byref(p), C.byref(n)), "gptb_state_buffer")
        return p.value, n.value

    def state_commit(self):
        self._check(self.lib.gptb_state_commit(self.h), "gptb_state_commit")
