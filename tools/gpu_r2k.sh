set -x
O=gpurun_out
STV_GEMM_PAIR=0 ncu --set full --import-source on --clock-control none -k regex:gemm_tf32 --launch-skip 1 --launch-count 1 -o $O/r2k_p0r2 -f python tools/bench_gemm.py --once --only 3 > $O/r2k_ncu0.log 2>&1
STV_GEMM_PAIR=0 STV_GEMM_RESIDENT=1 ncu --set full --import-source on --clock-control none -k regex:gemm_tf32 --launch-skip 1 --launch-count 1 -o $O/r2k_p0r1 -f python tools/bench_gemm.py --once --only 3 > $O/r2k_ncu1.log 2>&1
STV_GEMM_PAIR=2 ncu --set full --import-source on --clock-control none -k regex:gemm_tf32 --launch-count 2 -o $O/r2k_p2 -f python tools/bench_gemm.py --once --only 2 > $O/r2k_ncu2.log 2>&1
tail -2 $O/r2k_ncu2.log
ls -la $O
