"""BASELINE config 5 end to end: synthetic 3-D GPT with N = 32768 pairs (FP64 fit on rank 0), one NCCL broadcast of the model state,
then a dense transport lattice over [-0.1, 1.1]^3 -- generated ON THE DEVICE, outputs reduced ON THE DEVICE (gptb_query_grid) -- sharded
over the ranks of a torchrun launch.  Record tool:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/run_config5_full.py [log2_points] [--oracle]

log2_points = 29 is the full 1024 x 1024 x 512 lattice of the config (2^26 points per GPU on 8 GPUs).  Parity at this size:
(1) the per-column sums of the shards add up to one checksum per output (printed; run-to-run and rank-count invariant up to the last
bits), (2) every 2^19-th lattice point is returned and -- with --oracle -- compared with a CPU Cholesky solve (one refinement step) on
rank 0's host."""
import json, os, sys, time, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
import torch.distributed as dist
from gaussian_process_transportation_b200 import _lib as L
from gaussian_process_transportation_b200.distributed import broadcast_model, shard_bounds
from bench import synthetic_pairs, kabsch, KERNEL

args = [a for a in sys.argv[1:] if not a.startswith("--")]
log2p = int(args[0]) if args else 29
N = int(args[1]) if len(args) > 1 else 32768
want_oracle = "--oracle" in sys.argv
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    w = torch.zeros(1, device=dev); dist.all_reduce(w); torch.cuda.synchronize(dev)
e = [log2p // 3 + (1 if i < log2p % 3 else 0) for i in range(3)]          # 29 -> (10, 10, 9): 1024 x 1024 x 512
dims = [1 << v for v in e]
total = int(np.prod(dims))
origin = np.array([-0.1] * 3)
step = np.array([1.2 / (n - 1) for n in dims])
eng = L.Engine(local)
eng.set_variance_mode("int8w5")
eng.set_spatial(True)
rec = {"config": "c5 end to end", "N": N, "lattice": dims, "points": total, "n_gpus": world, "variance": "int8w5 + spatial (run-time guard may add planes)"}
aff = np.zeros(16)
if rank == 0:
    S, T = synthetic_pairs(N, 3, seed=0)
    R, Sc, Tc = kabsch(S, T)
    Sr = (R @ (S - Sc).T).T + Tc
    eng.set_train(Sr, T - Sr)
    eng.factorize(KERNEL["c"], KERNEL["ell"], KERNEL["s2"], KERNEL["jitter"])
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    info, _ = eng.factorize(KERNEL["c"], KERNEL["ell"], KERNEL["s2"], KERNEL["jitter"], want_lml=False)
    rec["fit_ms"] = (time.perf_counter() - t0) * 1e3
    rec["fit_tflops"] = N ** 3 / 3.0 / (rec["fit_ms"] * 1e-3) * 1e-12
    t0 = time.perf_counter()
    eng.prepare_variance()
    rec["prepare_variance_ms"] = (time.perf_counter() - t0) * 1e3
    rec["variance_guard"] = eng.variance_guard()
    aff = np.concatenate([R.ravel(), [1.0], Sc, Tc])
if world > 1:
    torch.cuda.synchronize(dev); dist.barrier()
    t0 = time.perf_counter()
    broadcast_model(eng, src=0)
    at = torch.from_numpy(aff).to(dev); dist.broadcast(at, src=0); aff = at.cpu().numpy()
    torch.cuda.synchronize(dev)
    rec["bcast_ms"] = (time.perf_counter() - t0) * 1e3
    rec["bcast_bytes"] = int(sum(eng.state_buffer(i)[1] for i in range(4)))
eng.set_affine(aff[:9].reshape(3, 3), aff[9], aff[10:13], aff[13:16])
lo, hi = shard_bounds(total, world, rank)
fl = L.MEAN | L.STD | L.JAC | L.AFFINE_IN
stride = 1 << max(0, log2p - 10)                                          # 1024 sample points over the lattice
eng.query_grid(origin, step, dims, fl, first=lo, count=min(hi - lo, 65536))           # warm-up (guard on the receiving ranks, workspace)
torch.cuda.synchronize(dev)
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
out = eng.query_grid(origin, step, dims, fl, first=lo, count=hi - lo, sample_stride=stride)
torch.cuda.synchronize(dev)
t_q = time.perf_counter() - t0
tt = torch.tensor([t_q], dtype=torch.float64, device=dev)
sums = torch.from_numpy(out["stats"][:, :2].copy()).to(dev)
mins = torch.from_numpy(out["stats"][:, 2].copy()).to(dev)
maxs = torch.from_numpy(out["stats"][:, 3].copy()).to(dev)
samples = [None] * world
if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dist.all_reduce(sums); dist.all_reduce(mins, op=dist.ReduceOp.MIN); dist.all_reduce(maxs, op=dist.ReduceOp.MAX)
    dist.all_gather_object(samples, (out["sample_index"], out["sample"]))
else:
    samples = [(out["sample_index"], out["sample"])]
if rank == 0:
    t_max = float(tt.item())
    rec.update({"query_s_max_over_ranks": t_max, "query_points_per_s": total / t_max, "columns": out["columns"],
                "checksum_sum": sums[:, 0].cpu().tolist(), "checksum_sumsq": sums[:, 1].cpu().tolist(), "min": mins.cpu().tolist(), "max": maxs.cpu().tolist(),
                "all_finite": bool(torch.isfinite(sums).all().item()), "guard_rank0_after_queries": eng.variance_guard()})
    idx = np.concatenate([s[0] for s in samples]); smp = np.vstack([s[1] for s in samples])
    rec["sample_points"] = int(len(idx))
    try:        # kept for the CPU oracle check off the GPU box (tools/check_config5_samples.py)
        os.makedirs("gpurun_out", exist_ok=True)
        np.savez_compressed(f"gpurun_out/config5_samples_log2p{log2p}_N{N}.npz", idx=idx, sample=smp, dims=np.array(dims), origin=origin, step=step, aff=aff)
    except Exception:
        pass
    # the sampled lattice points through the ordinary host-pointer query of the same engine (explicit coordinates): the grid path adds nothing
    ii = np.stack(np.unravel_index(idx, dims), axis=1)
    xs = origin + step * ii
    o = eng.query(xs, fl)
    pack = np.hstack([o["mean"], o["std"][:, :1], o["jac"].reshape(len(xs), -1)])
    rec["sample_vs_explicit_query_max_abs"] = float(np.max(np.abs(pack - smp)))
    if want_oracle:
        import scipy.linalg as sla
        from oracle.gp_oracle import rbf_cross
        try:
            from threadpoolctl import threadpool_limits
            threadpool_limits(limits=os.cpu_count())
        except Exception:
            pass
        t0 = time.perf_counter()
        c, ell, s2 = KERNEL["c"], np.array(KERNEL["ell"]), KERNEL["s2"]
        S, T = synthetic_pairs(N, 3, seed=0)
        R, Sc, Tc = kabsch(S, T)
        Sr = (R @ (S - Sc).T).T + Tc
        D = T - Sr
        xa = (R @ (xs - Sc).T).T + Tc
        K = rbf_cross(Sr, Sr, c, ell); K[np.diag_indices(N)] += s2 + KERNEL["jitter"]
        ks = rbf_cross(xa, Sr, c, ell)
        cf = sla.cho_factor(K.copy(), lower=True, overwrite_a=True, check_finite=False)
        alpha = sla.cho_solve(cf, D, check_finite=False)
        z = sla.cho_solve(cf, ks.T, check_finite=False)
        z += sla.cho_solve(cf, ks.T - K @ z, check_finite=False)
        var = c + s2 - np.einsum("mn,nm->m", ks, z)
        std = np.sqrt(np.maximum(var, 0.0)) - np.sqrt(s2)
        mean = ks @ alpha
        rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
        rec["oracle"] = {"what": "CPU Cholesky solve + one refinement step at the sampled lattice points", "cpu_s": time.perf_counter() - t0,
                         "mean_rel": rel(smp[:, :3], mean), "std_abs_over_sqrt_prior": float(np.max(np.abs(smp[:, 3] - std)) / np.sqrt(c + s2)),
                         "tolerance": {"mean_rel": 1e-9, "std_abs_over_sqrt_prior": 1e-7}}
    print(json.dumps(rec))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
