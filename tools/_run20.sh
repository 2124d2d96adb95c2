PG_SKINNY=2 SWEEP_B=32 SWEEP_T=388 timeout 300 python tools/kernel_sweep.py 2>&1 | tail -1
PG_SKINNY=2 SWEEP_B=8 SWEEP_T=388 timeout 300 python tools/kernel_sweep.py 2>&1 | tail -1
SWEEP_B=32 SWEEP_T=388 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm|attention|rope|norm|argmax|step_advance|embed' -c 900 --csv --log-file gpurun_out/launches_b32_r1b.csv python tools/kernel_sweep.py > gpurun_out/ncu_b32_r1b.log 2>&1
