python tools/ncu_fit.py 16384 > gpurun_out/plain_fit.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:potrf_trailing -s 41 -c 1 -o gpurun_out/prof_trailing_v1 python tools/ncu_fit.py 16384 > gpurun_out/ncu_trailing.log 2>&1
tail -n 2 gpurun_out/ncu_trailing.log
