timeout 600 python -m pytest tests/test_nets_gpu.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r41_tests.log; cat gpurun_out/r41_tests.log | cut -c1-250
