timeout 300 python -m pytest tests/test_nets_gpu.py tests/test_graph_gpu.py -q 2>&1 | tail -5 > gpurun_out/r14_tests.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r14_bench2.json 2> gpurun_out/r14_bench2.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r14_bench1.json 2> gpurun_out/r14_bench1.err
tail -3 gpurun_out/r14_tests.log; cut -c1-300 gpurun_out/r14_bench2.json; tail -5 gpurun_out/r14_bench2.err; cut -c1-200 gpurun_out/r14_bench1.json
