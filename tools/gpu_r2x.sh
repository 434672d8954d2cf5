set -x
O=gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -q -rf -x > $O/r2x_tests_conv.log 2>&1
tail -15 $O/r2x_tests_conv.log
timeout 300 python tools/bench_conv.py > $O/r2x_conv.txt 2>&1
STV_CONV_ROWSEG=0 timeout 300 python tools/bench_conv.py > $O/r2x_conv_off.txt 2>&1
STV_CONV_ROWSEG=2 timeout 300 python tools/bench_conv.py > $O/r2x_conv_all.txt 2>&1
