timeout 600 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r38_tests.log; cat gpurun_out/r38_tests.log | cut -c1-300
