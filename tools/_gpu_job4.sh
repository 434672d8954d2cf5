timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 8 --warmup 3 > gpurun_out/r37_bench_4gpu.json 2> gpurun_out/r37_bench_4gpu.err
cut -c1-260 gpurun_out/r37_bench_4gpu.json; tail -3 gpurun_out/r37_bench_4gpu.err
