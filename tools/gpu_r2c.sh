set -x
ls -la oracle/_ref
python -m pytest tests/test_loss_gpu.py tests/test_fullsize_gpu.py -m gpu -q -rf 2>&1 | tail -80 > gpurun_out/r2c_tests.log
python tools/profile_loss.py --mode disp > gpurun_out/r2c_profile_disp.txt 2>&1
STV_LIB=$PWD/slowtv_monodepth_b200/libstv_mb3.so python tools/profile_loss.py --mode disp > gpurun_out/r2c_profile_mb3.txt 2>&1
STV_LIB=$PWD/slowtv_monodepth_b200/libstv_mb4.so python tools/profile_loss.py --mode disp > gpurun_out/r2c_profile_mb4.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:photo_fused -c 1 -o gpurun_out/r2c_fused python tools/profile_loss.py --iters 2 > gpurun_out/r2c_ncu.log 2>&1
python -m pytest tests/test_plugin_gpu.py -m gpu -q -rf -s 2>&1 | grep -v Warning | tail -80 > gpurun_out/r2c_plugin.log
python -m pytest tests -m gpu -q -rf 2>&1 | tail -40 > gpurun_out/r2c_all.log
