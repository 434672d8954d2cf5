set -x
O=gpurun_out
python -m pytest tests/test_gemm_gpu.py tests/test_gemm_pair_gpu.py tests/test_conv_gpu.py -m gpu -q -rf -x > $O/r2j_tests.log 2>&1
tail -3 $O/r2j_tests.log
STV_GEMM_PAIR=0 python tools/bench_gemm.py > $O/r2j_gemm_p0r2.txt 2>&1
STV_GEMM_PAIR=0 STV_GEMM_RESIDENT=1 python tools/bench_gemm.py > $O/r2j_gemm_p0r1.txt 2>&1
STV_GEMM_PAIR=1 python tools/bench_gemm.py > $O/r2j_gemm_p1.txt 2>&1
STV_GEMM_PAIR=2 python tools/bench_gemm.py > $O/r2j_gemm_p2.txt 2>&1
python tools/bench_conv.py > $O/r2j_conv.txt 2>&1
python tools/profile_loss.py --mode disp > $O/r2j_profile_disp.txt 2>&1
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-torch-baseline > $O/r2j_bench.json 2> $O/r2j_bench.err
