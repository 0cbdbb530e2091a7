# round-1c ncu evidence for the spatial default (int8w5 + gptb_set_spatial): launch list of one bench step and one full capture
# of the skipping product kernel.  Each profiled command first runs clean without ncu.
set -x
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench_sp.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bench_c3_spatial.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_sp.log 2>&1
python tools/spatial_time.py 4096 65536 > gpurun_out/plain_oz_sp.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ozaki_trmm -s 1 -c 1 -o gpurun_out/prof_ozaki_v6_spatial python tools/spatial_time.py 4096 65536 > gpurun_out/ncu_oz_sp.log 2>&1
tail -n 2 gpurun_out/plain_oz_sp.log gpurun_out/ncu_oz_sp.log
