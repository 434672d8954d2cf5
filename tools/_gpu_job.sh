timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r46_tests.log; cat gpurun_out/r46_tests.log | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-160
