set -x
python -m pytest tests -m gpu -q -rf > gpurun_out/r2f_all.log 2>&1
python tools/profile_loss.py --mode disp > gpurun_out/r2f_profile_disp.txt 2>&1
STV_NO_TEX=1 python tools/profile_loss.py --mode disp > gpurun_out/r2f_profile_notex.txt 2>&1
python bench.py --steps 50 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
python bench.py --steps 30 --warmup 3 --api native --no-cpu-baseline --no-torch-baseline > gpurun_out/r2f_bench_native.json 2> gpurun_out/r2f_bench_native.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1
