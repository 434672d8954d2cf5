set -x
O=gpurun_out
ncu --set full --import-source on --clock-control none -k "regex:photo_fused|photo_error|pull_|fused_finalize|fused_loss_reduce" --launch-count 7 -o $O/r2u_loss -f python tools/profile_loss.py --iters 2 > $O/r2u_ncu_loss.log 2>&1
tail -2 $O/r2u_ncu_loss.log
ncu --set full --import-source on --clock-control none -k regex:gemm_tf32 --launch-count 6 -o $O/r2u_gemm -f python tools/bench_gemm.py --once --only 2 > $O/r2u_ncu_gemm.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $O/r2u_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-torch-baseline > $O/r2u_bench_under_ncu.log 2>&1
gzip -f $O/r2u_bench_launches.csv
export STV_LIB=$PWD/slowtv_monodepth_b200/csrc/build/libstv_trace.so
S="7680,384,1536 1920,768,3072 7680,1536,384 30720,192,768 1536,384,7680,1,1,8"
STV_GEMM_PAIR=0 python tools/gemm_trace.py $S > $O/r2u_trace_single.txt 2>&1
STV_GEMM_PAIR=2 python tools/gemm_trace.py $S > $O/r2u_trace_pair.txt 2>&1
unset STV_LIB
python tools/bench_gemm.py > $O/r2u_gemm_heur.txt 2>&1
STV_GEMM_PAIR=0 python tools/bench_gemm.py > $O/r2u_gemm_p0.txt 2>&1
ls -la $O
