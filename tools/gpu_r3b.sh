set -x
O=gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -q -rf -x > $O/r3b_tests_conv.log 2>&1
tail -12 $O/r3b_tests_conv.log
timeout 300 python tools/bench_conv.py > $O/r3b_conv.txt 2>&1
export STV_LIB=$PWD/slowtv_monodepth_b200/csrc/build/libstv_trace.so
STV_CONV_ROWSEG=2 timeout 120 python tools/conv3_trace.py 16,96,160,64,64 8,386,642,32,16 8,194,322,32,32 > $O/r3b_conv3_trace.txt 2>&1
cat $O/r3b_conv3_trace.txt | cut -c1-330
