set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench_default_final.json 2> gpurun_out/bench_default_final.err; tail -c 700 gpurun_out/bench_default_final.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | head -c 250
