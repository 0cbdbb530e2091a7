# round-2 session A: state check + what-if of the product kernel + overlap A/B in spatial mode
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log; tail -3 gpurun_out/r2a_pytest.log
python tools/whatif.py pipeline > gpurun_out/r2a_pipeline.log 2>&1; tail -5 gpurun_out/r2a_pipeline.log
GPTB_LIB_PATH=$PWD/gaussian_process_transportation_b200/lib/libgptb200_whatif.so python tools/whatif.py whatif 4096 16384 > gpurun_out/r2a_whatif.log 2>&1; tail -40 gpurun_out/r2a_whatif.log
