set -x
python -m pytest tests -m gpu -q -rf > gpurun_out/r2e_all.log 2>&1
python tools/bench_dw.py > gpurun_out/r2e_dw_rows.txt 2>&1
STV_DW_RING=1 python tools/bench_dw.py > gpurun_out/r2e_dw_ring.txt 2>&1
python tools/profile_loss.py --mode disp > gpurun_out/r2e_profile_disp.txt 2>&1
python bench.py --steps 50 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
python tools/step_profile.py > gpurun_out/r2e_step_profile.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:'photo_fused|pull_|photo_error|fused_' -c 8 -o gpurun_out/r2e_loss python tools/profile_loss.py --iters 2 > gpurun_out/r2e_ncu.log 2>&1
