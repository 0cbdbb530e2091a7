set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_c3_8gpu_spatial.json 2> gpurun_out/bench_c3_8gpu.err; tail -c 400 gpurun_out/bench_c3_8gpu_int8w5.json; tail -2 gpurun_out/bench_c3_8gpu.err
timeout 300 $TR bench.py --gpus 8 --steps 5 --warmup 3 --workload c4 --no-cpu-baseline > gpurun_out/bench_c4_8gpu_spatial.json 2> gpurun_out/bench_c4_8gpu.err; tail -c 400 gpurun_out/bench_c4_8gpu_int8w5.json; tail -2 gpurun_out/bench_c4_8gpu.err
timeout 400 $TR bench.py --gpus 8 --steps 3 --warmup 3 --workload c5 > gpurun_out/bench_c5_8gpu_spatial.json 2> gpurun_out/bench_c5_8gpu.err; tail -c 400 gpurun_out/bench_c5_8gpu_spatial.json; tail -2 gpurun_out/bench_c5_8gpu.err
