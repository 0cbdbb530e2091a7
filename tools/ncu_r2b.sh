# round-2 final ncu evidence for the skipping product kernel (eight epilogue warps): N = 4096 and N = 16384 with five planes, N = 16384 with the
# extra diagonal (the mode the guard selects there).  Each profiled command first runs clean without ncu.  usage: bash tools/ncu_r2b.sh <tag>
TAG=${1:-r02b}
set -x
mkdir -p gpurun_out
for cfg in "4096 int8w5" "16384 int8w5" "16384 int8w5p"; do
  set -- $cfg
  python tools/spatial_time.py $1 65536 $2 0 > gpurun_out/${TAG}_plain_$1_$2.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:ozaki_trmm -s 1 -c 1 -o gpurun_out/${TAG}_prof_ozaki_N$1_$2 python tools/spatial_time.py $1 65536 $2 0 > gpurun_out/${TAG}_ncu_$1_$2.log 2>&1
  tail -n 1 gpurun_out/${TAG}_plain_$1_$2.log
  python tools/ncu_keymetrics.py gpurun_out/${TAG}_prof_ozaki_N$1_$2.ncu-rep > gpurun_out/${TAG}_ozaki_N$1_$2_ncu_key_metrics.txt
  ncu -i gpurun_out/${TAG}_prof_ozaki_N$1_$2.ncu-rep --page details 2>/dev/null | sed -n 1,140p | cut -c1-170 > gpurun_out/${TAG}_ozaki_N$1_$2_ncu_details.txt
  rm -f gpurun_out/${TAG}_prof_ozaki_N$1_$2.ncu-rep      # 29 MB each: the merged gpurun_out/ is limited to 64 MiB
done
