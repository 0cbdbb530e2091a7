# product-kernel iteration: correctness of the INT8/spatial paths, overlap A/B, what-if timings.  usage: bash tools/run_oz.sh <tag>
TAG=${1:-x}
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "int8 or spatial or sliced or overlap or smoke" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/whatif.py pipeline > gpurun_out/${TAG}_pipeline.log 2>&1; tail -5 gpurun_out/${TAG}_pipeline.log
GPTB_LIB_PATH=$PWD/gaussian_process_transportation_b200/lib/libgptb200_whatif.so timeout 300 python tools/whatif.py whatif 4096 16384 > gpurun_out/${TAG}_whatif.log 2>&1; tail -40 gpurun_out/${TAG}_whatif.log
