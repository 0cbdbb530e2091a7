python tools/ozaki_time.py 4096 65536 > gpurun_out/plain_oz.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:kstar_kernel -s 3 -c 1 -o gpurun_out/prof_kstar_v3 python tools/ozaki_time.py 4096 65536 > gpurun_out/ncu_ks.log 2>&1
tail -n 2 gpurun_out/ncu_ks.log
