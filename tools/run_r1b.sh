set -x
python -m pytest tests -m gpu -x -q -k "int8 or refit" 2>&1 | tail -5
python tools/ozaki_check.py 4096 16384 2>&1 | tail -2
python tools/ozaki_time.py 4096 262144 2>&1 | tail -3
python tools/ozaki_time.py 16384 65536 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_int8w5_v2.json 2> gpurun_out/bench_c3_int8w5_v2.err; tail -c 3000 gpurun_out/bench_c3_int8w5_v2.json
