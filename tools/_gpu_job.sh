timeout 300 python -m pytest tests/test_conv_gpu.py -q 2>&1 | grep -vE "^\s*$" | tail -60 > gpurun_out/r4_conv_tests.log
timeout 300 python -m pytest tests/test_nets_gpu.py tests/test_gemm_gpu.py -q 2>&1 | tail -15 > gpurun_out/r4_nets_tests.log
timeout 200 python tools/bench_conv.py > gpurun_out/r4_bench_conv.txt 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r4_bench.json 2> gpurun_out/r4_bench.err
timeout 200 python tools/step_profile.py > gpurun_out/r4_step_profile.txt 2>&1
tail -5 gpurun_out/r4_conv_tests.log
