set -x
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
python tools/profile_loss.py > gpurun_out/r1_profile_loss.txt 2>&1
python tools/step_profile.py > gpurun_out/r1_step_profile.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'photo|smooth' -c 12 -o gpurun_out/r1_loss python tools/profile_loss.py --iters 2 > gpurun_out/r1_ncu_loss.log 2>&1
ls -la gpurun_out
