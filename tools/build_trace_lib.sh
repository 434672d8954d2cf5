# Developer build of libstv with the GEMM timeline instrumentation (never shipped; lives under the git-ignored csrc/build/).
set -e
cd "$(dirname "$0")/.."
B=slowtv_monodepth_b200/csrc/build
python -c "from slowtv_monodepth_b200 import _build; _build.build()"
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DSTV_GEMM_TRACE"
nvcc $F -c slowtv_monodepth_b200/csrc/stv_gemm.cu -o $B/stv_gemm_trace.o
nvcc $F -c slowtv_monodepth_b200/csrc/stv_conv3.cu -o $B/stv_conv3_trace.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $B/libstv_trace.so $(ls $B/*.o | grep -v "stv_gemm.o\|stv_conv3.o\|_trace.o") $B/stv_gemm_trace.o $B/stv_conv3_trace.o -lcudart
ls -la $B/libstv_trace.so
