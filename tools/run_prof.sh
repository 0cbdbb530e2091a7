# per-role cycle accounting of the skipping product kernel (what-if build).  usage: bash tools/run_prof.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
GPTB_LIB_PATH=$PWD/gaussian_process_transportation_b200/lib/libgptb200_whatif.so timeout 300 python tools/whatif.py whatif 4096 16384 > gpurun_out/${TAG}_prof.log 2>&1
grep profile gpurun_out/${TAG}_prof.log
