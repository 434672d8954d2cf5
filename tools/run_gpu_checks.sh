# Everything the round-end driver runs, on one B200 (gpurun -- bash tools/run_gpu_checks.sh): GPU test-suite, default bench,
# the other BASELINE.json configurations, the reference arm, smoke(). Outputs land in gpurun_out/.
set -x
O=gpurun_out
python -m pytest tests -m gpu -q -rf > $O/checks_tests.log 2>&1
tail -4 $O/checks_tests.log
python bench.py > $O/checks_bench.json 2> $O/checks_bench.err
tail -2 $O/checks_bench.err
for c in c2 c4 c5; do
  python bench.py --config $c --steps 30 --warmup 3 --no-cpu-baseline --no-torch-baseline > $O/checks_bench_$c.json 2> $O/checks_bench_$c.err
  tail -1 $O/checks_bench_$c.err
done
python bench.py --impl reference --steps 5 --warmup 2 > $O/checks_bench_ref.json 2> $O/checks_bench_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/checks_smoke.log 2>&1; tail -2 $O/checks_smoke.log
