set -x
O=gpurun_out
date +%s
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 > $O/r2t_bench_2gpu.json 2> $O/r2t_bench_2gpu.err
echo "rc=$?"; date +%s
tail -3 $O/r2t_bench_2gpu.err
cut -c1-200 $O/r2t_bench_2gpu.json
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 1 --impl reference > $O/r2t_ref_2gpu.json 2> $O/r2t_ref_2gpu.err
echo "rc=$?"; date +%s
cut -c1-300 $O/r2t_ref_2gpu.json
