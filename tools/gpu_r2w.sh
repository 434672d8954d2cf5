set -x
O=gpurun_out
python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-torch-baseline > $O/r2w_bench.json 2> $O/r2w_bench.err
STV_DW_L8=1 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-torch-baseline > $O/r2w_bench_l8.json 2> $O/r2w_bench_l8.err
python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-torch-baseline > $O/r2w_bench_b.json 2> $O/r2w_bench_b.err
python tools/step_profile.py > $O/r2w_step_profile.txt 2>&1
STV_DW_L8=1 python tools/step_profile.py > $O/r2w_step_profile_l8.txt 2>&1
