timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/t_final.log; cat gpurun_out/t_final.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; tail -c 300 gpurun_out/bench_final2.json; tail -2 gpurun_out/bench_final2.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_final2.csv python bench.py --steps 8 --warmup 3 --cpu-steps 0 --kv-off-steps 1 --vision-batch 8 --batched 0 > gpurun_out/ncu_final2.log 2>&1
tail -c 200 gpurun_out/ncu_final2.log
