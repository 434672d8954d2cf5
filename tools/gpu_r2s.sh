set -x
O=gpurun_out
python tools/bench_conv.py > $O/r2s_conv.txt 2>&1
STV_GEMM_PAIR_CONV=1 python tools/bench_conv.py > $O/r2s_conv_pc.txt 2>&1
python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-torch-baseline > $O/r2s_bench.json 2> $O/r2s_bench.err
python tools/step_profile.py > $O/r2s_step_profile.txt 2>&1
