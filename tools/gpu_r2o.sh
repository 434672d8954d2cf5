set -x
O=gpurun_out
python -m pytest tests -m gpu -q -rf > $O/r2o_tests.log 2>&1
tail -4 $O/r2o_tests.log
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-torch-baseline > $O/r2o_bench.json 2> $O/r2o_bench.err
tail -2 $O/r2o_bench.err
STV_WGRAD_STREAM_OFF=1 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-torch-baseline > $O/r2o_bench_off.json 2> $O/r2o_bench_off.err
python tools/profile_loss.py --mode disp > $O/r2o_profile_disp.txt 2>&1
