python tools/ncu_fit.py 4096 > gpurun_out/plain_fit.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:potrf_diag -s 5 -c 1 -o gpurun_out/prof_diag_v2 python tools/ncu_fit.py 4096 > gpurun_out/ncu_diag.log 2>&1
tail -n 2 gpurun_out/ncu_diag.log
