# ncu source-level capture of the generator (kstar_kernel, fused digit emission) at N = 4096.  usage: bash tools/ncu_kstar_r2.sh <tag>
TAG=${1:-r02k}
mkdir -p gpurun_out
python tools/spatial_time.py 4096 65536 > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:kstar_kernel -s 2 -c 1 -o gpurun_out/${TAG}_prof_kstar_N4096 python tools/spatial_time.py 4096 65536 > gpurun_out/${TAG}_ncu.log 2>&1
ls -la gpurun_out/${TAG}_prof*; tail -2 gpurun_out/${TAG}_plain.log
