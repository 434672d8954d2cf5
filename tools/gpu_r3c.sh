set -x
O=gpurun_out
python -m pytest tests -m gpu -q -rf > $O/r3c_tests.log 2>&1
tail -4 $O/r3c_tests.log
python bench.py > $O/r3c_bench.json 2> $O/r3c_bench.err
tail -2 $O/r3c_bench.err
for c in c2 c4 c5; do
  python bench.py --config $c --steps 30 --warmup 3 --no-cpu-baseline --no-torch-baseline > $O/r3c_bench_$c.json 2> $O/r3c_bench_$c.err
  tail -1 $O/r3c_bench_$c.err
done
python bench.py --impl reference --steps 5 --warmup 2 > $O/r3c_bench_ref.json 2> $O/r3c_bench_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/r3c_smoke.log 2>&1; tail -2 $O/r3c_smoke.log
