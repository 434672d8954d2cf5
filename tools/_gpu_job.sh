timeout 300 python -m pytest tests/test_nets_gpu.py -m gpu -q -k "layernorm or networks" 2>&1 | tail -4 | cut -c1-200
timeout 100 python tools/bench_dw.py 2>&1 | tail -5
