set -x
O=gpurun_out
python -m pytest tests/test_gemm_gpu.py tests/test_gemm_pair_gpu.py tests/test_conv_gpu.py tests/test_nets_gpu.py -m gpu -q -rf > $O/r2v_tests.log 2>&1
tail -3 $O/r2v_tests.log
for pf in 0 2 4 8; do
  STV_GEMM_PF=$pf STV_GEMM_PAIR=0 python tools/bench_gemm.py > $O/r2v_gemm_pf$pf.txt 2>&1
done
STV_GEMM_PF=0 python tools/bench_conv.py > $O/r2v_conv_pf0.txt 2>&1
STV_GEMM_PF=4 python tools/bench_conv.py > $O/r2v_conv_pf4.txt 2>&1
STV_GEMM_PF=0 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-torch-baseline > $O/r2v_bench_pf0.json 2> $O/r2v_bench_pf0.err
STV_GEMM_PF=4 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-torch-baseline > $O/r2v_bench_pf4.json 2> $O/r2v_bench_pf4.err
STV_GEMM_PF=8 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-torch-baseline > $O/r2v_bench_pf8.json 2> $O/r2v_bench_pf8.err
python tools/bench_dw.py > $O/r2v_dw.txt 2>&1
