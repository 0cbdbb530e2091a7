# full GPU test suite with durations.  usage: bash tools/run_tests.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -x -q -s --durations=15 > gpurun_out/${TAG}_pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_full.log
grep -v "^lenghtscales\|^Is the map\|^Rotation\|^Scale\|^\[\[" gpurun_out/${TAG}_pytest_full.log | tail -70
