# ncu capture of the factorisation's chain kernels (potrf_diag_kernel, potrf_spine_kernel, trsv_back_chain_kernel) in the N = 4096 fit.
# One ncu use per session, after the same command ran clean.  usage: bash tools/ncu_spine.sh <tag>
TAG=${1:-r02s}
mkdir -p gpurun_out
python tools/fit_launches.py 4096 3 > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'potrf_diag_kernel|potrf_spine_kernel|trsv_back_chain_kernel' -s 10 -c 3 \
    -o gpurun_out/${TAG}_prof_spine python tools/fit_launches.py 4096 1 > gpurun_out/${TAG}_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_plain.log
python tools/ncu_keymetrics.py gpurun_out/${TAG}_prof_spine.ncu-rep > gpurun_out/${TAG}_spine_ncu_key_metrics.txt
ncu -i gpurun_out/${TAG}_prof_spine.ncu-rep --page details 2>/dev/null | cut -c1-170 > gpurun_out/${TAG}_spine_ncu_details.txt
rm -f gpurun_out/${TAG}_prof_spine.ncu-rep
head -c 3000 gpurun_out/${TAG}_spine_ncu_key_metrics.txt
