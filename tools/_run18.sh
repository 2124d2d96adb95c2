timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 32 --warmup 4 > gpurun_out/bench_n2_r1b.json 2> gpurun_out/bench_n2_r1b.err; tail -c 900 gpurun_out/bench_n2_r1b.json; tail -3 gpurun_out/bench_n2_r1b.err
timeout 600 python -m pytest tests -m gpu -x -q -k "tensor_parallel or tp" 2>&1 | tail -3
