set -x
O=gpurun_out
export STV_LIB=$PWD/slowtv_monodepth_b200/csrc/build/libstv_trace.so
STV_CONV_ROWSEG=2 python tools/conv3_trace.py 16,96,160,64,64 8,386,642,32,16 8,194,322,32,32 > $O/r2y_conv3_trace.txt 2>&1
cat $O/r2y_conv3_trace.txt | cut -c1-400
