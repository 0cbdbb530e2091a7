"""BASELINE config 4 end to end: synthetic 3-D GPT with N = 16384 pairs, LML hyper-parameter optimisation (L-BFGS-B on the host,
LML + gradient on the GPU), then the full M = 2^26 query stream (mean + std + Jacobian, through the host-pointer C ABI) sharded
over the ranks of a torchrun launch.  Record tool:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/run_config4_full.py [log2_M] [n_restarts]

Fit on rank 0, one NCCL broadcast of the model state, block partition of the queries, no data-path collective.  Parity at this
size is checked through size-independent properties: K alpha = y at training inputs (posterior mean identity) and the std range."""
import contextlib, io, json, os, sys, time, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
import torch.distributed as dist
import gaussian_process_transportation_b200 as g
from gaussian_process_transportation_b200 import _lib as L
from gaussian_process_transportation_b200.distributed import broadcast_model, shard_bounds
from oracle.gp_oracle import synthetic_pairs       # input generator only
from sklearn.gaussian_process.kernels import RBF, WhiteKernel, ConstantKernel as C

log2m = int(sys.argv[1]) if len(sys.argv) > 1 else 26
n_restarts = int(sys.argv[2]) if len(sys.argv) > 2 else 1
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N, M = 16384, 1 << log2m
S, T = synthetic_pairs(N, 3, seed=0)
kern = C(0.1) * RBF([0.3, 0.3, 0.3]) + WhiteKernel(1e-3)
gp = g.GaussianProcess(kernel=kern, n_restarts_optimizer=n_restarts, device=local, variance_mode="int8w5", spatial=True)
rec = {}
if rank == 0:
    t = g.GaussianProcessTransportation(kernel_transport=kern)
    t.method = g.PolicyTransportation(gp)
    t.source_distribution, t.target_distribution = S, T
    np.random.seed(0)
    l0 = gp._engine.launch_count()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        t.fit_transportation()
    rec["fit_transportation_s"] = time.perf_counter() - t0
    rec["launches_during_fit"] = gp._engine.launch_count() - l0
    rec["kernel"] = str(gp.kernel)
    rec["lml"] = float(gp.gp.log_marginal_likelihood_value_)
    gp._ensure_fitted_factor()
    a = t.method.affine_transform
    aff = np.concatenate([a.rotation_matrix.ravel(), [float(a.scale)], a.S_centroid, a.T_centroid])
    alpha = gp.gp.alpha_
    idx = np.arange(0, N, 64)
    rec["property_mean_at_train_rel_err"] = float(np.linalg.norm(gp.predict(gp.X[idx]) - (gp.Y[idx] - gp.noise_var_ * alpha[idx])) /
                                                  np.linalg.norm(gp.Y[idx]))
else:
    aff = np.zeros(16)
eng = gp._engine
if world > 1:
    warm = torch.zeros(1, device=dev); dist.all_reduce(warm); torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    broadcast_model(eng, src=0)
    at = torch.from_numpy(aff).to(dev); dist.broadcast(at, src=0); aff = at.cpu().numpy()
    torch.cuda.synchronize(dev)
    rec["bcast_s"] = time.perf_counter() - t0
eng.set_affine(aff[:9].reshape(3, 3), aff[9], aff[10:13], aff[13:16])
lo, hi = shard_bounds(M, world, rank)
m = hi - lo
rng = np.random.default_rng(100 + rank)
x = torch.from_numpy(-0.1 + 1.2 * rng.random((m, 3))).pin_memory()
mean = torch.empty(m, 3, dtype=torch.float64).pin_memory(); std = torch.empty(m, 3, dtype=torch.float64).pin_memory()
jac = torch.empty(m, 3, 3, dtype=torch.float64).pin_memory()
import ctypes as Cc
dp = Cc.POINTER(Cc.c_double); cast = lambda tt: Cc.cast(tt.data_ptr(), dp)
flags = L.MEAN | L.STD | L.JAC | L.AFFINE_IN
eng.lib.gptb_query(eng.h, cast(x), min(m, 1 << 16), flags, None, cast(mean), cast(std), cast(jac), None, None, None, None, None, None)   # warm-up
if world > 1:
    dist.barrier()
torch.cuda.synchronize(dev)
t0 = time.perf_counter()
rc = eng.lib.gptb_query(eng.h, cast(x), m, flags, None, cast(mean), cast(std), cast(jac), None, None, None, None, None, None)
assert rc == 0, eng.error()
q_s = time.perf_counter() - t0
tt = torch.tensor([q_s, float(std.min()), -float(std.max()), float(torch.isfinite(mean).all() and torch.isfinite(jac).all())], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    rec.update({"config": "c4 end to end", "N": N, "M": M, "n_gpus": world, "n_restarts": n_restarts, "variance": "int8w5 + spatial",
                "query_s_max_over_ranks": float(tt[0]), "query_points_per_s": M / float(tt[0]),
                "std_max": -float(tt[2]), "all_finite": bool(tt[3] > 0),
                "total_s_fit_bcast_query": rec["fit_transportation_s"] + rec.get("bcast_s", 0.0) + float(tt[0])})
    print(json.dumps(rec), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
