# multi-GPU session: bench.py under torchrun + the config-5 lattice tool.  usage: bash tools/run_multi.sh <tag> <ngpu> <log2 lattice points> [--oracle]
TAG=$1; NG=$2; LP=$3; shift; shift; shift
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $NG --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_${NG}gpu.json 2> gpurun_out/${TAG}_bench_${NG}gpu.err; echo "bench rc=$?"
tail -c 600 gpurun_out/${TAG}_bench_${NG}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 tools/run_config5_full.py $LP "$@" > gpurun_out/${TAG}_config5_${NG}gpu.json 2> gpurun_out/${TAG}_config5_${NG}gpu.err; echo "c5 rc=$?"
tail -c 600 gpurun_out/${TAG}_config5_${NG}gpu.err; tail -c 1500 gpurun_out/${TAG}_config5_${NG}gpu.json
