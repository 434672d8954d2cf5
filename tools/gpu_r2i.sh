set -x
O=gpurun_out
python -m pytest tests/test_fullsize_gpu.py tests/test_loss_gpu.py -m gpu -q -rf > $O/r2i_tests.log 2>&1
tail -3 $O/r2i_tests.log
STV_GEMM_PAIR=0 ncu --set full --import-source on --clock-control none -k regex:gemm_tf32 -o $O/r2i_gemm_pair0 -f python tools/bench_gemm.py --once --only 2 > $O/r2i_ncu0.log 2>&1
STV_GEMM_PAIR=1 ncu --set full --import-source on --clock-control none -k regex:gemm_tf32 -o $O/r2i_gemm_pair1 -f python tools/bench_gemm.py --once --only 2 > $O/r2i_ncu1.log 2>&1
tail -2 $O/r2i_ncu1.log
