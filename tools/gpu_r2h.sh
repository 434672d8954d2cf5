set -x
O=gpurun_out
python -m pytest tests -m gpu -q -rf > $O/r2h_tests.log 2>&1
tail -5 $O/r2h_tests.log
python tools/debug_fullsize.py > $O/r2h_debug_fullsize.txt 2>&1
STV_GEMM_PAIR=0 python tools/bench_gemm.py > $O/r2h_gemm_pair0.txt 2>&1
STV_GEMM_PAIR=1 python tools/bench_gemm.py > $O/r2h_gemm_pair1.txt 2>&1
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-torch-baseline > $O/r2h_bench.json 2> $O/r2h_bench.err
tail -3 $O/r2h_bench.err
STV_GEMM_PAIR=0 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-torch-baseline > $O/r2h_bench_pair0.json 2> $O/r2h_bench_pair0.err
