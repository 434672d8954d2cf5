timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r31_tests.log; cat gpurun_out/r31_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r31_bench.json 2> gpurun_out/r31_bench.err
cut -c1-330 gpurun_out/r31_bench.json; tail -3 gpurun_out/r31_bench.err
