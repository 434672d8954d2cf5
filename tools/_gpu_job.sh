timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_nets_gpu.py tests/test_step_gpu.py tests/test_graph_gpu.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r42_tests.log; cat gpurun_out/r42_tests.log | cut -c1-250
timeout 200 python tools/step_profile.py --top 70 > gpurun_out/r42_step_profile.txt 2>&1; grep "distinct\|head" gpurun_out/r42_step_profile.txt
