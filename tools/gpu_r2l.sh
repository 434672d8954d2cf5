set -x
O=gpurun_out
export STV_LIB=$PWD/slowtv_monodepth_b200/csrc/build/libstv_trace.so
S="7680,384,1536 1920,768,3072 7680,1536,384 30720,192,768 1536,384,7680,1,1,8 7680,384,1536,0,1"
STV_GEMM_PAIR=0 python tools/gemm_trace.py $S > $O/r2l_trace_single.txt 2>&1
STV_GEMM_PAIR=0 STV_GEMM_RESIDENT=1 python tools/gemm_trace.py $S > $O/r2l_trace_single_r1.txt 2>&1
STV_GEMM_PAIR=2 python tools/gemm_trace.py $S > $O/r2l_trace_pair.txt 2>&1
unset STV_LIB
python tools/bench_gemm.py > $O/r2l_gemm_heur.txt 2>&1
