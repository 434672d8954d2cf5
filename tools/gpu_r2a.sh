set -x
nvidia-smi -L
python -m pytest tests -m gpu -q -rf 2>&1 | tail -60 > gpurun_out/r2a_tests.log
python -m pytest tests/test_step_gpu.py -m gpu -q -s 2>&1 | grep -v Warning | tail -40 > gpurun_out/r2a_step.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1
timeout 60 tools/probe/cta2_gemm_probe > gpurun_out/r2a_probe_cta2.txt 2>&1; echo "rc=$?" >> gpurun_out/r2a_probe_cta2.txt
timeout 60 tools/probe/umma_rowshift_probe > gpurun_out/r2a_probe_rowshift.txt 2>&1; echo "rc=$?" >> gpurun_out/r2a_probe_rowshift.txt
nvidia-smi --query-gpu=name,memory.used --format=csv
