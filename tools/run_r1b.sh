set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in c3 c4 c5; do
timeout 400 python bench.py --workload $w > gpurun_out/bench_${w}_spatial.json 2> gpurun_out/bench_${w}_spatial.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${w}_spatial.json").read().strip().splitlines()[-1])
print("$w spatial: value %.4g ms/step %.1f e2e %.4g share %.3f gen %.3f launch_ms %.3f fit %.1f prep %.1f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["roofline"]["share_of_step"],d["roofline"]["generator_share_of_step"],d["roofline"]["launch_ms"],d["fit_ms"],d["prepare_variance_ms"]), d["clocks"]["sm_mhz"], d["clocks"]["power_w"], d["cpu_baseline"]["parity_vs_gpu"] if d["cpu_baseline"] else None)
PY
done
timeout 300 python bench.py --spatial 0 --no-cpu-baseline > gpurun_out/bench_c3_natural.json 2>/dev/null; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_c3_natural.json").read().strip().splitlines()[-1])
print("c3 natural: value %.4g ms/step %.1f launch_ms %.3f"%(d["value"],d["ms_per_step"],d["roofline"]["launch_ms"]))
PY
