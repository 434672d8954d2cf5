timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r28_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r28_bench.json 2> gpurun_out/r28_bench.err
timeout 200 python tools/step_profile.py --top 60 > gpurun_out/r28_step_profile.txt 2>&1
cat gpurun_out/r28_tests.log; cut -c1-400 gpurun_out/r28_bench.json; tail -3 gpurun_out/r28_bench.err
