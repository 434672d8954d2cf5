timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r12_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r12_bench.json 2> gpurun_out/r12_bench.err
timeout 200 python tools/step_profile.py --top 40 > gpurun_out/r12_step_profile.txt 2>&1
tail -3 gpurun_out/r12_tests.log; cut -c1-260 gpurun_out/r12_bench.json
