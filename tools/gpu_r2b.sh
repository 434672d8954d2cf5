set -x
python -m pytest tests/test_loss_gpu.py tests/test_fullsize_gpu.py -m gpu -q -rf 2>&1 | tail -80 > gpurun_out/r2b_tests.log
python tools/profile_loss.py --mode disp > gpurun_out/r2b_profile_disp.txt 2>&1
python tools/profile_loss.py --mode depth > gpurun_out/r2b_profile_depth.txt 2>&1
python tools/profile_loss.py --mode two-pass > gpurun_out/r2b_profile_twopass.txt 2>&1
for r in 8 16 32 64; do STV_FUSED_ROWS=$r python tools/profile_loss.py --mode disp 2>&1 | tail -2 > gpurun_out/r2b_profile_rows$r.txt; done
python tools/profile_loss.py --mode disp --n 4 --b 4 > gpurun_out/r2b_profile_n4.txt 2>&1
python -m pytest tests/test_plugin_gpu.py -m gpu -q -rf -s 2>&1 | grep -v Warning | tail -60 > gpurun_out/r2b_plugin.log
python -m pytest tests -m gpu -q -rf 2>&1 | tail -40 > gpurun_out/r2b_all.log
