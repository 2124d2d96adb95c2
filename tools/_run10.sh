timeout 900 python -m pytest tests/test_gpu_ops.py -k "wide_swap_ab" -x -q 2>&1 | tail -2
VISION_BATCH=1 timeout 300 python tools/prefill_profile.py > gpurun_out/pp_10.log 2>&1; tail -1 gpurun_out/pp_10.log
PG_WIDE_ROT=0 VISION_BATCH=1 timeout 300 python tools/prefill_profile.py > gpurun_out/pp_10_norot.log 2>&1; tail -1 gpurun_out/pp_10_norot.log
VISION_BATCH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm|attention|rope|norm' -c 400 --csv --log-file gpurun_out/launches_wide.csv python tools/prefill_profile.py > gpurun_out/ncu_wide.log 2>&1
