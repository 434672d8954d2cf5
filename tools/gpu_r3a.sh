set -x
O=gpurun_out
export STV_LIB=$PWD/slowtv_monodepth_b200/csrc/build/libstv_trace.so
STV_C3_NOSHIFT=1 STV_CONV_ROWSEG=2 python tools/conv3_trace.py 16,96,160,64,64 8,386,642,32,16 > $O/r3a_noshift.txt 2>&1
cat $O/r3a_noshift.txt | cut -c1-330
