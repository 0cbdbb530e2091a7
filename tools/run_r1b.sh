set -x
python tools/ozaki_check.py 32768 2>&1 | tail -1
python bench.py --workload c5 --steps 3 > gpurun_out/bench_c5_r1d.json 2> gpurun_out/bench_c5_r1d.err; tail -c 600 gpurun_out/bench_c5_r1d.json; head -c 400 gpurun_out/bench_c5_r1d.json; tail -3 gpurun_out/bench_c5_r1d.err
python -m pytest tests -m gpu -x -q -k "int8" 2>&1 | tail -2
