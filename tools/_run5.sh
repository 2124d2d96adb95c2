timeout 600 python -m pytest tests/test_gpu_ops.py -k "gemm_tcgen05 or attention_tc" -x -q 2>&1 | tail -3 > gpurun_out/t_5.log
cat gpurun_out/t_5.log
timeout 120 python tools/gemm_bench.py > gpurun_out/gb_5.log 2>&1; tail -1 gpurun_out/gb_5.log
ONLY_VISION=1 timeout 200 python tools/prefill_profile.py > gpurun_out/pp_5.log 2>&1; tail -1 gpurun_out/pp_5.log
