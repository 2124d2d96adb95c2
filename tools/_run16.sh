timeout 900 python -m pytest tests -m gpu -x -q -k "skinny or splitk or batch or decode or gemm_tcgen05" 2>&1 | tail -3
for b in 8 32; do SWEEP_B=$b SWEEP_T=388 timeout 300 python tools/kernel_sweep.py 2>&1 | tail -1; done
PG_TC_PDL=0 SWEEP_B=32 SWEEP_T=388 timeout 300 python tools/kernel_sweep.py 2>&1 | tail -1
