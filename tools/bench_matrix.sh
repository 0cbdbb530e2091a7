# bench.py over the non-default workloads / variance modes (does every documented invocation run and pass its parity gate?)
mkdir -p gpurun_out
i=0
while read -r args; do
  i=$((i+1))
  timeout 900 python bench.py $args > gpurun_out/matrix_$i.json 2> gpurun_out/matrix_$i.err; rc=$?
  python - "$args" $rc gpurun_out/matrix_$i.json <<'PY'
import json, sys
args, rc, f = sys.argv[1], sys.argv[2], sys.argv[3]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    g = d.get("variance_guard") or {}
    print(f"rc={rc} [{args}] value={d['value']:.4g} e2e={d['e2e']['value']:.4g} fit_ms={d['fit_ms']:.3g} guard={g.get('used_slices')}+{g.get('used_extra_diagonal')} "
          f"parity64={(d.get('parity_vs_fp64_path') or {}).get('std_abs_over_sqrt_prior')} invalid={d.get('invalid')}")
except Exception as e:
    print(f"rc={rc} [{args}] NO LINE: {e}")
PY
done <<'ARGS'
--workload c2 --steps 2 --warmup 3 --no-c4 --no-small-n
--workload c1 --steps 2 --warmup 3 --no-c4 --no-small-n --no-cpu-baseline
--workload c4 --steps 2 --warmup 3 --no-small-n --no-cpu-baseline
--workload c5 --steps 1 --warmup 3 --no-small-n
--variance fp64 --steps 2 --warmup 3 --no-c4 --no-small-n --no-cpu-baseline
--variance int8x6 --spatial 0 --steps 2 --warmup 3 --no-c4 --no-small-n --no-cpu-baseline
--variance int8w5p --steps 2 --warmup 3 --no-c4 --no-small-n --no-cpu-baseline
--variance int8w6 --spatial 0 --steps 2 --warmup 3 --no-c4 --no-small-n --no-cpu-baseline
ARGS
