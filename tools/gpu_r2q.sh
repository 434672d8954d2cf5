set -x
O=gpurun_out
STV_GEMM_PAIR=2 python -m pytest tests/test_conv_gpu.py tests/test_gemm_gpu.py tests/test_gemm_pair_gpu.py tests/test_nets_gpu.py -m gpu -q -rf > $O/r2q_tests_p2.log 2>&1
tail -4 $O/r2q_tests_p2.log
python -m pytest tests -m gpu -q -rf > $O/r2q_tests.log 2>&1
tail -4 $O/r2q_tests.log
python tools/bench_conv.py > $O/r2q_conv.txt 2>&1
STV_GEMM_PAIR=0 python tools/bench_conv.py > $O/r2q_conv_p0.txt 2>&1
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-torch-baseline > $O/r2q_bench.json 2> $O/r2q_bench.err
tail -2 $O/r2q_bench.err
python tools/profile_loss.py --mode disp --kernels > $O/r2q_profile_disp.txt 2>&1
