timeout 600 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|Error|learn_K=|assert" | tail -25 > gpurun_out/r16_tests.log
timeout 100 python tools/bench_dw.py > gpurun_out/r16_bench_dw.txt 2>&1
timeout 200 python tools/bench_gemm.py > gpurun_out/r16_bench_gemm.txt 2>&1
timeout 200 python tools/bench_conv.py > gpurun_out/r16_bench_conv.txt 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r16_bench.json 2> gpurun_out/r16_bench.err
timeout 200 python tools/step_profile.py --top 40 > gpurun_out/r16_step_profile.txt 2>&1
cat gpurun_out/r16_tests.log; cut -c1-200 gpurun_out/r16_bench.json; cat gpurun_out/r16_bench_dw.txt
