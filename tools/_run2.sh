timeout 300 python -m pytest tests/test_gpu_ops.py -k "gemm_tcgen05" -x -q 2>&1 | tail -5 > gpurun_out/t_gemm2.log
cat gpurun_out/t_gemm2.log
timeout 120 python tools/gemm_bench.py > gpurun_out/gb_2cta_b.log 2>&1; tail -1 gpurun_out/gb_2cta_b.log
ONLY_VISION=1 timeout 200 python tools/prefill_profile.py > gpurun_out/pp_2cta_b.log 2>&1; tail -1 gpurun_out/pp_2cta_b.log
