timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
SWEEP_B=32 SWEEP_T=388 timeout 300 python tools/kernel_sweep.py 2>&1 | tail -1
timeout 300 python tools/prefill_profile.py 2>&1 | tail -1
