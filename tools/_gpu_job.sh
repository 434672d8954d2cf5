timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_gemm_gpu.py tests/test_nets_gpu.py -q 2>&1 | tail -25 > gpurun_out/r8_tests.log
timeout 200 python tools/bench_conv.py > gpurun_out/r8_bench_conv.txt 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r8_bench.json 2> gpurun_out/r8_bench.err
timeout 200 python tools/step_profile.py --top 25 > gpurun_out/r8_step_profile.txt 2>&1
tail -3 gpurun_out/r8_tests.log; cut -c1-200 gpurun_out/r8_bench.json
