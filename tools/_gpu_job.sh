timeout 600 python -m pytest tests/test_aspect_gpu.py tests/test_graph_gpu.py -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r33_tests.log; cat gpurun_out/r33_tests.log
