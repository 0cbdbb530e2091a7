# ncu capture of the matrix-vector variance kernel (one point at N = 16384: L^-1 streamed once).  usage: bash tools/ncu_trmv.sh <tag>
TAG=${1:-r02v}
mkdir -p gpurun_out
python tools/small_batch_time.py 16384 > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:trmv_partial_kernel -s 30 -c 1 -o gpurun_out/${TAG}_prof_trmv python tools/small_batch_time.py 16384 > gpurun_out/${TAG}_ncu.log 2>&1
python tools/ncu_keymetrics.py gpurun_out/${TAG}_prof_trmv.ncu-rep > gpurun_out/${TAG}_trmv_ncu_key_metrics.txt
rm -f gpurun_out/${TAG}_prof_trmv.ncu-rep
cat gpurun_out/${TAG}_trmv_ncu_key_metrics.txt | head -30
