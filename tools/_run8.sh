timeout 900 python -m pytest tests/test_gpu_ops.py -x -q 2>&1 | tail -3 > gpurun_out/t_ops8.log
cat gpurun_out/t_ops8.log
timeout 300 python tools/prefill_profile.py > gpurun_out/pp_8.log 2>&1; tail -1 gpurun_out/pp_8.log
PG_OP_TIMING=1 ONLY_VISION=1 timeout 300 python tools/prefill_profile.py > gpurun_out/pp_8_ops.log 2>&1; tail -1 gpurun_out/pp_8_ops.log
