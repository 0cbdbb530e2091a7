"""Measure the cuBLAS FP64 DGEMM rate (torch.matmul, float64) on this GPU: the "FP64 tensor peak" denominator.
MEASURED_PEAKS.json only carries HBM and bf16."""
import json, sys, time
import torch

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    burst = 2.0 * n ** 3 / best * 1e-9
    # sustained: back to back for ~3 s
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = max(3, int(3000.0 / best))
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record(); torch.cuda.synchronize()
    sustained = 2.0 * n ** 3 * reps / e0.elapsed_time(e1) * 1e-9
    print(json.dumps({"n": n, "dgemm_tflops_burst": round(burst, 2), "dgemm_tflops_sustained": round(sustained, 2),
                      "ms_best": round(best, 3), "gpu": torch.cuda.get_device_name(0)}))

if __name__ == "__main__":
    main()
