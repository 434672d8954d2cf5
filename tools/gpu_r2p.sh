set -x
O=gpurun_out
python -m pytest tests -m gpu -q -rf > $O/r2p_tests.log 2>&1
tail -4 $O/r2p_tests.log
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-torch-baseline > $O/r2p_bench.json 2> $O/r2p_bench.err
tail -2 $O/r2p_bench.err
STV_BRANCH_STREAM_OFF=1 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-torch-baseline > $O/r2p_bench_nobranch.json 2> $O/r2p_bench_nobranch.err
python tools/profile_loss.py --mode disp --kernels > $O/r2p_profile_disp.txt 2>&1
