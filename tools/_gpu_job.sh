timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r43_tests.log; cat gpurun_out/r43_tests.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r43_smoke.log
timeout 200 python tools/step_profile.py --top 70 > gpurun_out/r43_step_profile.txt 2>&1; grep "distinct\|colred4\|bn_" gpurun_out/r43_step_profile.txt | cut -c1-120
timeout 600 python bench.py > gpurun_out/r43_bench.json 2> gpurun_out/r43_bench.err; cut -c1-200 gpurun_out/r43_bench.json; tail -2 gpurun_out/r43_bench.err
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r43_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r43_ncu_bench.log 2>&1
