timeout 600 python -m pytest tests -m gpu -x -q -k "attention_tc" 2>&1 | tail -3
ONLY_PREFILL=1 timeout 300 python tools/prefill_profile.py 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -3
