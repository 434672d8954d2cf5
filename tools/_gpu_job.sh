timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r11_tests.log
timeout 200 python tools/bench_gemm.py > gpurun_out/r11_bench_gemm.txt 2>&1
timeout 200 python tools/bench_conv.py > gpurun_out/r11_bench_conv.txt 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r11_bench.json 2> gpurun_out/r11_bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r11_bench_eager.json 2> gpurun_out/r11_bench_eager.err
timeout 200 python tools/step_profile.py --top 40 > gpurun_out/r11_step_profile.txt 2>&1
tail -3 gpurun_out/r11_tests.log; cut -c1-260 gpurun_out/r11_bench.json; cut -c1-260 gpurun_out/r11_bench_eager.json
