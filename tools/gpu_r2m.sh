set -x
O=gpurun_out
python -m pytest tests/test_gemm_gpu.py tests/test_gemm_pair_gpu.py tests/test_conv_gpu.py -m gpu -q -rf -x > $O/r2m_tests.log 2>&1
tail -3 $O/r2m_tests.log
STV_GEMM_RESIDENT=1 python -m pytest tests/test_gemm_gpu.py tests/test_conv_gpu.py -m gpu -q -rf -x > $O/r2m_tests_r1.log 2>&1
tail -3 $O/r2m_tests_r1.log
STV_GEMM_PAIR=0 python tools/bench_gemm.py > $O/r2m_gemm_p0r2.txt 2>&1
STV_GEMM_PAIR=0 STV_GEMM_RESIDENT=1 python tools/bench_gemm.py > $O/r2m_gemm_p0r1.txt 2>&1
STV_GEMM_PAIR=2 python tools/bench_gemm.py > $O/r2m_gemm_p2.txt 2>&1
STV_GEMM_PAIR=2 STV_GEMM_PAIR_EPI=8 python tools/bench_gemm.py > $O/r2m_gemm_p2e8.txt 2>&1
python tools/bench_conv.py > $O/r2m_conv_r2.txt 2>&1
STV_GEMM_RESIDENT=1 python tools/bench_conv.py > $O/r2m_conv_r1.txt 2>&1
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-torch-baseline > $O/r2m_bench.json 2> $O/r2m_bench.err
STV_GEMM_RESIDENT=1 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-torch-baseline > $O/r2m_bench_r1.json 2> $O/r2m_bench_r1.err
STV_GEMM_RESIDENT=1 STV_GEMM_PAIR=2 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-torch-baseline > $O/r2m_bench_r1p2.json 2> $O/r2m_bench_r1p2.err
