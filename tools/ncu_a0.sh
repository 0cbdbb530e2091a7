python tools/ncu_a0.py 16384 262144 > gpurun_out/plain_a0.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:kstar_kernel -c 1 -o gpurun_out/prof_kstar_v1 python tools/ncu_a0.py 16384 262144 > gpurun_out/ncu_a0.log 2>&1
ncu --set full --clock-control none -k regex:gram_lower -c 1 -o gpurun_out/prof_gram_v1 python tools/ncu_a0.py 16384 1024 > gpurun_out/ncu_gram.log 2>&1
tail -2 gpurun_out/ncu_a0.log gpurun_out/ncu_gram.log
