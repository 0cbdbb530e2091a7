set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_c3_r1b.json 2> gpurun_out/bench_c3_r1b.err; tail -c 600 gpurun_out/bench_c3_r1b.json
python bench.py --workload c4 > gpurun_out/bench_c4_r1b.json 2> gpurun_out/bench_c4_r1b.err; tail -c 600 gpurun_out/bench_c4_r1b.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_c3_r1b.json 2>/dev/null; cat gpurun_out/bench_reference_c3_r1b.json | head -c 300
