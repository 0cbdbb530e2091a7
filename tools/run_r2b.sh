# round-2 session B: multi-warp MMA issue -- correctness (bit-identical to the dense issue order), what-if, overlap A/B
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "int8 or spatial or sliced or overlap or smoke" > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log; tail -5 gpurun_out/r2b_pytest.log
timeout 300 python tools/whatif.py pipeline > gpurun_out/r2b_pipeline.log 2>&1; tail -5 gpurun_out/r2b_pipeline.log
GPTB_LIB_PATH=$PWD/gaussian_process_transportation_b200/lib/libgptb200_whatif.so timeout 300 python tools/whatif.py whatif 4096 16384 > gpurun_out/r2b_whatif.log 2>&1; tail -40 gpurun_out/r2b_whatif.log
