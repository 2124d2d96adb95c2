timeout 900 python -m pytest tests/test_gpu_ops.py -k "wide_swap_ab" -x -q 2>&1 | tail -6 > gpurun_out/t_wide.log
cat gpurun_out/t_wide.log
timeout 300 python tools/prefill_profile.py > gpurun_out/pp_9.log 2>&1; tail -1 gpurun_out/pp_9.log
PG_WIDE=0 timeout 300 python tools/prefill_profile.py > gpurun_out/pp_9_nowide.log 2>&1; tail -1 gpurun_out/pp_9_nowide.log
