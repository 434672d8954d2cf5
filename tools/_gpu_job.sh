timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_nets_gpu.py tests/test_step_gpu.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r36_tests.log; cat gpurun_out/r36_tests.log
timeout 200 python tools/bench_gemm.py 2>&1 | cut -c1-75 | grep "gelu" | tee gpurun_out/r36_gemm.txt
