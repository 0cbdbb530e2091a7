set -x
for sp in 1 0; do
python bench.py --workload c4 --steps 3 --no-cpu-baseline --spatial $sp > gpurun_out/bench_c4_sp$sp.json 2> gpurun_out/bench_c4_sp$sp.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_c4_sp$sp.json").read().strip().splitlines()[-1])
print("c4 spatial=$sp: value %.4g ms/step %.1f e2e %.4g share %.3f gen %.3f launch_ms %.3f fit %.1f prep %.1f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["roofline"]["share_of_step"],d["roofline"]["generator_share_of_step"],d["roofline"]["launch_ms"],d["fit_ms"],d["prepare_variance_ms"]), d["clocks"]["sm_mhz"], d["clocks"]["power_w"])
PY
done
python bench.py --workload c5 --steps 3 --spatial 1 > gpurun_out/bench_c5_sp1.json 2> gpurun_out/bench_c5_sp1.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_c5_sp1.json").read().strip().splitlines()[-1])
print("c5 spatial=1: value %.4g ms/step %.1f e2e %.4g share %.3f gen %.3f launch_ms %.3f fit %.1f prep %.1f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["roofline"]["share_of_step"],d["roofline"]["generator_share_of_step"],d["roofline"]["launch_ms"],d["fit_ms"],d["prepare_variance_ms"]), d["clocks"]["sm_mhz"], d["clocks"]["power_w"])
PY
