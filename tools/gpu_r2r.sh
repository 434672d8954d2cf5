set -x
O=gpurun_out
python -m pytest tests -m gpu -q -rf > $O/r2r_tests.log 2>&1
tail -4 $O/r2r_tests.log
python tools/bench_conv.py > $O/r2r_conv.txt 2>&1
STV_GEMM_PAIR=0 python tools/bench_conv.py > $O/r2r_conv_p0.txt 2>&1
python tools/step_profile.py > $O/r2r_step_profile.txt 2>&1
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-torch-baseline > $O/r2r_bench.json 2> $O/r2r_bench.err
