timeout 600 python -m pytest tests/test_aspect_gpu.py -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r32_tests.log; cat gpurun_out/r32_tests.log
timeout 100 python tools/bench_aspect.py 2>&1 | tail -4 | tee gpurun_out/r32_bench_aspect.txt
