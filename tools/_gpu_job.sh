timeout 600 python -m pytest tests/test_loss_gpu.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r26_tests.log
cat gpurun_out/r26_tests.log
timeout 200 python tools/profile_loss.py > gpurun_out/r26_profile_loss.txt 2>&1; cat gpurun_out/r26_profile_loss.txt
