# final round-1 record: smoke, full GPU tests, the three single-GPU bench workloads, ncu launch list + capture of the product kernel
set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for w in c3 c4 c5; do
timeout 400 python bench.py --workload $w > gpurun_out/bench_${w}_spatial.json 2> gpurun_out/bench_${w}_spatial.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${w}_spatial.json").read().strip().splitlines()[-1])
print("$w spatial: value %.4g ms/step %.1f e2e %.4g share %.3f gen %.3f launch_ms %.3f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["roofline"]["share_of_step"],d["roofline"]["generator_share_of_step"],d["roofline"]["launch_ms"]), d["clocks"]["sm_mhz"], d["clocks"]["power_w"], d["parity_vs_fp64_path"]["std_abs_over_sqrt_prior"], d["fp64_dmma_variance"]["value"])
PY
done
bash tools/ncu_r1c.sh
