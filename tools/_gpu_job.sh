timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r10_tests.log
timeout 200 python tools/bench_gemm.py > gpurun_out/r10_bench_gemm.txt 2>&1
timeout 100 python tools/profile_loss.py > gpurun_out/r10_profile_loss.txt 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r10_bench.json 2> gpurun_out/r10_bench.err
timeout 200 python tools/step_profile.py --top 30 > gpurun_out/r10_step_profile.txt 2>&1
tail -3 gpurun_out/r10_tests.log; cut -c1-260 gpurun_out/r10_bench.json
