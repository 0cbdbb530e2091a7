python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench_c3_int8x6.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/ozaki_time.py 4096 65536 > gpurun_out/plain_oz.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ozaki_trmm -s 0 -c 1 -o gpurun_out/prof_ozaki_v4 python tools/ozaki_time.py 4096 65536 > gpurun_out/ncu_oz.log 2>&1
tail -n 2 gpurun_out/ncu_oz.log
