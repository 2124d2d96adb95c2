timeout 900 python bench.py > gpurun_out/bench6.json 2> gpurun_out/bench6.err; tail -c 600 gpurun_out/bench6.json
ONLY_VISION=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_vit -s 60 -c 1 -o gpurun_out/prof_r1_attn_vit -f python tools/prefill_profile.py > gpurun_out/ncu_attn_vit.log 2>&1
ONLY_VISION=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 200 -c 4 -o gpurun_out/prof_r1_gemm_tc2 -f python tools/prefill_profile.py > gpurun_out/ncu_gemm_tc2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
