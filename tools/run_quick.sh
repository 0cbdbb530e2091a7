# selected GPU tests + bench.  usage: bash tools/run_quick.sh <tag> "<pytest -k expr>" [bench args]
TAG=${1:-x}; KEXPR=$2; shift; shift
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "$KEXPR" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -25 gpurun_out/${TAG}_pytest.log | cut -c1-220
if [ -n "$1" ]; then timeout 900 python bench.py "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/${TAG}_bench.err; fi
