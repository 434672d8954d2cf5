timeout 600 python -m pytest tests/test_fullsize_gpu.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r44_tests.log; cat gpurun_out/r44_tests.log | cut -c1-300
