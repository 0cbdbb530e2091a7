# round-2 ncu evidence: launch list of one bench step and full captures of the skipping product kernel at N = 4096 and N = 16384.
# Each profiled command first runs clean without ncu.  usage: bash tools/ncu_r2.sh <tag>
TAG=${1:-r02}
set -x
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c4 --no-small-n > gpurun_out/${TAG}_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches_bench_c3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c4 --no-small-n > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/spatial_time.py 4096 65536 > gpurun_out/${TAG}_plain_oz4096.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ozaki_trmm -s 1 -c 1 -o gpurun_out/${TAG}_prof_ozaki_N4096 python tools/spatial_time.py 4096 65536 > gpurun_out/${TAG}_ncu_oz4096.log 2>&1
python tools/spatial_time.py 16384 65536 > gpurun_out/${TAG}_plain_oz16384.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ozaki_trmm -s 1 -c 1 -o gpurun_out/${TAG}_prof_ozaki_N16384 python tools/spatial_time.py 16384 65536 > gpurun_out/${TAG}_ncu_oz16384.log 2>&1
tail -n 2 gpurun_out/${TAG}_plain_oz4096.log gpurun_out/${TAG}_plain_oz16384.log
ls -la gpurun_out/${TAG}_prof*
