# round-1b ncu evidence for the default bench configuration (int8w5): launch list of one bench step, one full capture of
# the product kernel and one of the fused generator.  Each profiled command first runs clean without ncu.
set -x
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench_w5.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench_c3_int8w5.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_w5.log 2>&1
python tools/ozaki_time.py 4096 65536 > gpurun_out/plain_oz_w5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ozaki_trmm -s 3 -c 1 -o gpurun_out/prof_ozaki_v5_int8w5 python tools/ozaki_time.py 4096 65536 > gpurun_out/ncu_oz_w5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kstar_kernel -s 5 -c 1 -o gpurun_out/prof_kstar_v4_int8w5 python tools/ozaki_time.py 4096 65536 > gpurun_out/ncu_ks_w5.log 2>&1
tail -n 3 gpurun_out/plain_oz_w5.log gpurun_out/ncu_oz_w5.log gpurun_out/ncu_ks_w5.log
