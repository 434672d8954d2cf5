set -x
ls -la oracle/_ref > gpurun_out/r2d_ls.txt 2>&1
python tools/profile_loss.py --mode disp > gpurun_out/r2d_profile_disp.txt 2>&1
python -m pytest tests -m gpu -q -rf > gpurun_out/r2d_all.log 2>&1
python -m pytest tests/test_step_gpu.py tests/test_plugin_gpu.py -m gpu -q -s 2>&1 | grep -E "whole-gradient|of the bound|drop-in|passed|failed" > gpurun_out/r2d_diag.log
python bench.py --steps 30 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
python tools/step_profile.py > gpurun_out/r2d_step_profile.txt 2>&1
