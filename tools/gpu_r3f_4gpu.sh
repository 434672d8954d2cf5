set -x
O=gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 50 --warmup 3 > $O/r3f_bench_4gpu.json 2> $O/r3f_bench_4gpu.err
echo "rc=$?"
cut -c1-260 $O/r3f_bench_4gpu.json
tail -3 $O/r3f_bench_4gpu.err
