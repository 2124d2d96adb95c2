ONLY_PREFILL=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_wide -s 40 -c 5 -o gpurun_out/prof_wide -f python tools/prefill_profile.py > gpurun_out/ncu_wide2.log 2>&1
tail -2 gpurun_out/ncu_wide2.log
