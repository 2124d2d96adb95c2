timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/t_all3.log
cat gpurun_out/t_all3.log
timeout 300 python tools/prefill_profile.py > gpurun_out/pp_7.log 2>&1; tail -1 gpurun_out/pp_7.log
PG_TC_PDL=0 timeout 300 python tools/prefill_profile.py > gpurun_out/pp_7_nopdl.log 2>&1; tail -1 gpurun_out/pp_7_nopdl.log
PG_OP_TIMING=1 ONLY_VISION=1 timeout 300 python tools/prefill_profile.py > gpurun_out/pp_7_ops.log 2>&1; tail -1 gpurun_out/pp_7_ops.log
