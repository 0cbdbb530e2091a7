# ncu launch list of one default bench step (per-kernel share of the step), after the same command ran clean.  usage: bash tools/ncu_bench_launches.sh <tag>
TAG=${1:-r02f}
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c4 --no-small-n > gpurun_out/${TAG}_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches_bench_c3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c4 --no-small-n > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches_bench_c3.csv > gpurun_out/${TAG}_launches_bench_c3_summary.txt
head -12 gpurun_out/${TAG}_launches_bench_c3_summary.txt | cut -c1-150
ls -la gpurun_out/${TAG}_launches_bench_c3.csv
