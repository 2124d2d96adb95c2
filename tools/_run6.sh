timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/t_all2.log
cat gpurun_out/t_all2.log
timeout 600 python bench.py --steps 64 --warmup 8 --cpu-steps 0 --kv-off-steps 0 --vision-batch 0 > gpurun_out/bench7.json 2> gpurun_out/bench7.err
python -c "
import json; d=json.load(open('gpurun_out/bench7.json')); print(d['value'], d['ms_per_step'], d['e2e'])"
SWEEP_B=32 SWEEP_T=388 timeout 300 python tools/kernel_sweep.py > gpurun_out/b32_plain2.log 2>&1; tail -1 gpurun_out/b32_plain2.log
SWEEP_B=32 SWEEP_T=388 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm|attention|rope|rmsnorm|argmax|step_advance|embed' -c 1500 --csv --log-file gpurun_out/launches_b32_r1.csv python tools/kernel_sweep.py > gpurun_out/ncu_b32_r1.log 2>&1
tail -1 gpurun_out/ncu_b32_r1.log
