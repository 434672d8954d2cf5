timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r39_tests.log; cat gpurun_out/r39_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r39_smoke.log
timeout 600 python bench.py > gpurun_out/r39_bench.json 2> gpurun_out/r39_bench.err; cut -c1-220 gpurun_out/r39_bench.json; tail -2 gpurun_out/r39_bench.err
