timeout 600 python -m pytest tests -m gpu -x -q -k "gemm_tcgen05" 2>&1 | tail -3
timeout 120 python tools/gemm_bench.py 2>&1 | tail -1
ONLY_VISION=1 timeout 200 python tools/prefill_profile.py 2>&1 | tail -1
