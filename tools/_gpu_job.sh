timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r17_tests.log
timeout 200 python tools/bench_gemm.py > gpurun_out/r17_bench_gemm.txt 2>&1
timeout 200 python tools/bench_conv.py > gpurun_out/r17_bench_conv.txt 2>&1
timeout 100 python tools/bench_dw.py > gpurun_out/r17_bench_dw.txt 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r17_bench.json 2> gpurun_out/r17_bench.err
STV_GEMM_PERSISTENT=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r17_bench_np.json 2> gpurun_out/r17_bench_np.err
timeout 200 python tools/step_profile.py --top 30 > gpurun_out/r17_step_profile.txt 2>&1
cat gpurun_out/r17_tests.log; cut -c1-200 gpurun_out/r17_bench.json;  cut -c1-200 gpurun_out/r17_bench_np.json; cat gpurun_out/r17_bench_dw.txt
